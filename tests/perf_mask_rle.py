#!/usr/bin/env python
"""Instance masks of 300 detections at 1024^2 from blob-like 128^2 masks: dense booleans + pinned copy to the host against
COCO run-length codes computed on the device (development aid; wall clock with synchronisation, 5 repeats)."""
import os, sys, time
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ape_b200  # noqa: E402
from ape_b200 import ops  # noqa: E402
from ape_b200.structures import Instances, Boxes  # noqa: E402

N, S, H, W = 300, 128, 1024, 1024
g = torch.Generator().manual_seed(0)
x = torch.randn(N, 1, S // 8, S // 8, generator=g)
masks = (F.interpolate(x, size=(S, S), mode="bicubic", align_corners=False)[:, 0] > 0).cuda()
c = torch.rand(N, 2, generator=g) * torch.tensor([W, H])
wh = torch.rand(N, 2, generator=g) * torch.tensor([W, H]) * 0.5 + 8.0
boxes = torch.cat([(c - wh / 2), (c + wh / 2)], 1).cuda()

def T(fn, n=5):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): r = fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3, r

def dense():
    m = ops.paste_masks_in_image(masks, boxes, (H, W))
    return Instances((H, W), pred_boxes=Boxes(boxes), pred_masks=m).to("cpu")

ms_d, inst = T(dense)
ms_r, rles = T(lambda: ops.paste_masks_rle(masks, boxes, (H, W)))
runs = sum(len(r["counts"]) for r in rles)
print(f"dense paste + pinned D2H of {inst.pred_masks.numel() / 1e6:.0f} MB: {ms_d:.2f} ms;  run-length codes on the device: {ms_r:.2f} ms "
      f"({runs / 1e3:.0f} k characters for {N} masks)")
