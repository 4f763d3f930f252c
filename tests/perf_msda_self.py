#!/usr/bin/env python
"""Encoder-shape MSDA (APE-L_D 1024^2, fp16, fused entry): region/window kernel vs generic fused kernel, CUDA events,
rotating inputs (4 sets > L2).  Development aid; prints one JSON line per variant."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ape_b200  # noqa: E402
from ape_b200 import ops  # noqa: E402

DEV = "cuda:0"
shapes = [(256, 256), (128, 128), (64, 64), (32, 32), (16, 16)]
H, D, P, L = 8, 32, 4, 5
S = sum(h * w for h, w in shapes)
ss = torch.tensor(shapes)
st = torch.cat([ss.new_zeros(1), (ss[:, 0] * ss[:, 1]).cumsum(0)[:-1]])
pts = []
for (h, w) in shapes:
    ys, xs = torch.meshgrid((torch.arange(h) + 0.5) / h, (torch.arange(w) + 0.5) / w, indexing="ij")
    pts.append(torch.stack([xs.reshape(-1), ys.reshape(-1)], -1))
ref = torch.cat(pts, 0)[None, :, None, :].expand(1, S, L, 2).contiguous().to(DEV)
n_off = H * L * P * 2
for dtype in (torch.float16, torch.bfloat16):
    for off_scale in (1.0, 2.5, 4.0):
        sets = []
        g = torch.Generator().manual_seed(0)
        # sampling_offsets bias grid (+-1..4 px) + noise, as the module initialises it
        th = torch.arange(H, dtype=torch.float32) * (2 * 3.141592653589793 / H)
        grid = torch.stack([th.cos(), th.sin()], -1)
        grid = (grid / grid.abs().max(-1, keepdim=True)[0]).view(H, 1, 1, 2).repeat(1, L, P, 1)
        for i in range(P):
            grid[:, :, i] *= i + 1
        for _ in range(4):
            value = torch.randn(1, S, H, D, generator=g).to(DEV, dtype)
            offs = grid.view(1, 1, -1) + torch.randn(1, S, n_off, generator=g) * (off_scale / 4)
            qo = torch.cat([offs, torch.randn(1, S, H * L * P, generator=g)], -1).to(DEV, dtype)
            sets.append((value, qo))
        for name, hs in (("generic", None), ("region", shapes)):
            def run(i):
                v, qo = sets[i % 4]
                return ops.ms_deform_attn_fused_forward(v, ss.to(DEV), st.to(DEV), qo[..., :n_off], qo[..., n_off:], ref, P, host_shapes=hs)
            ssd, std = ss.to(DEV), st.to(DEV)
            for i in range(4):
                run(i)
            torch.cuda.synchronize()
            evs = []
            for i in range(20):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); run(i); b.record()
                evs.append((a, b))
            torch.cuda.synchronize()
            ts = sorted(a.elapsed_time(b) for a, b in evs)
            nbytes = 2 * S * H * D + 2 * S * H * L * P * 3 + 4 * S * L * 2 + 2 * S * H * D
            print(json.dumps({"dtype": str(dtype), "noise_px": off_scale / 4, "variant": name, "ms_median": round(ts[10], 4),
                              "ms_min": round(ts[0], 4), "alg_GBps": round(nbytes / ts[10] / 1e6, 1)}), flush=True)
