"""Multi-GPU plumbing of the detection path: images shard one (or B/G) per GPU with no data-path
collective; the only exchange is ONE gather of fixed-shape packed detections per batch
(SURVEY.md §8e).  It replaces the reference's pickled `comm.gather(self._predictions, dst=0)` over a
gloo side group (ape/evaluation/lvis_evaluation.py:103-104) with a single tensor collective on the
job's own process group (NCCL over NVLink/NVSwitch on GPUs, gloo in the CPU tests)."""
from typing import List, Optional

import torch
import torch.distributed as dist

from .structures import Boxes, Instances

PACK_WIDTH = 8  # x1, y1, x2, y2, score, class, image_height, image_width


def shard(items: List, rank: int, world: int) -> List:
    """Contiguous, balanced split of a batch across ranks (rank r gets items [lo_r, hi_r))."""
    n = len(items)
    lo = (n * rank) // world
    hi = (n * (rank + 1)) // world
    return items[lo:hi]


def pack_detections(instances: List[Instances], max_det: int, device) -> torch.Tensor:
    """[len(instances), max_det + 1, PACK_WIDTH] fp32: row 0 holds the count, rows 1.. the detections."""
    out = torch.zeros((len(instances), max_det + 1, PACK_WIDTH), dtype=torch.float32, device=device)
    for i, inst in enumerate(instances):
        k = min(len(inst), max_det)
        out[i, 0, 0] = k
        if k:
            out[i, 1:k + 1, 0:4] = inst.pred_boxes.tensor[:k].to(device)
            out[i, 1:k + 1, 4] = inst.scores[:k].to(device)
            out[i, 1:k + 1, 5] = inst.pred_classes[:k].to(device=device, dtype=torch.float32)
        out[i, 0, 6], out[i, 0, 7] = float(inst.image_size[0]), float(inst.image_size[1])
    return out


def unpack_detections(packed: torch.Tensor) -> List[Instances]:
    res = []
    packed = packed.cpu()
    for p in packed:
        k = int(p[0, 0])
        res.append(Instances((int(p[0, 6]), int(p[0, 7])), pred_boxes=Boxes(p[1:k + 1, 0:4].clone()),
                             scores=p[1:k + 1, 4].clone(), pred_classes=p[1:k + 1, 5].to(torch.int64)))
    return res


def gather_detections(instances: List[Instances], max_det: int, device, dst: int = 0,
                      group: Optional[dist.ProcessGroup] = None) -> Optional[List[Instances]]:
    """One collective: every rank contributes the same number of images (pad with empty Instances if needed).
    Returns the concatenated per-image results on `dst` (rank order = image order), None elsewhere."""
    packed = pack_detections(instances, max_det, device)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return unpack_detections(packed)
    world = dist.get_world_size(group)
    if dist.get_rank(group) == dst:
        bufs = [torch.empty_like(packed) for _ in range(world)]
        dist.gather(packed, bufs, dst=dst, group=group)
        return unpack_detections(torch.cat(bufs, 0))
    dist.gather(packed, None, dst=dst, group=group)
    return None


# ---- device-side path: no host round trip before the collective ---------------------------------------------------
def unpack_packed(packed: torch.Tensor) -> List[dict]:
    """Rows of `DeformableDETRSegmVL.forward_packed` ([images, topk, 13], any device) -> the reference's output list
    [{"instances": Instances}] on the host: the kept detections rescaled to the requested output size, clipped, empty boxes
    dropped (detectron2 detector_postprocess).  One device->host copy for everything."""
    host = packed.to("cpu")
    out = []
    for p in host:
        nk = int(p[0, 8])
        h, w, oh, ow = (float(v) for v in p[0, 9:13])
        r = p[:nk]
        b = r[:, :4].clone()
        if h > 0 and w > 0:
            b[:, 0::2] *= ow / w
            b[:, 1::2] *= oh / h
        b = torch.stack((b[:, 0].clamp(min=0, max=ow), b[:, 1].clamp(min=0, max=oh), b[:, 2].clamp(min=0, max=ow),
                         b[:, 3].clamp(min=0, max=oh)), dim=-1)
        keep = ((b[:, 2] - b[:, 0]) > 0) & ((b[:, 3] - b[:, 1]) > 0)
        out.append({"instances": Instances((int(oh), int(ow)), pred_boxes=Boxes(b[keep]), scores=r[keep, 4].clone(),
                                           pred_classes=r[keep, 5].to(torch.int64), query_index=r[keep, 6].to(torch.int64)),
                    "num_candidates": int(p[0, 7])})
    return out


def gather_packed(packed: torch.Tensor, dst: int = 0, group: Optional[dist.ProcessGroup] = None) -> Optional[List[dict]]:
    """ONE collective on the packed DEVICE tensor (NCCL gather on the compute stream; gloo in the CPU tests): non-destination
    ranks neither copy to the host nor synchronise.  Returns the per-image results of all ranks on `dst`, None elsewhere."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return unpack_packed(packed)
    world = dist.get_world_size(group)
    if dist.get_rank(group) == dst:
        buf = torch.empty((world,) + tuple(packed.shape), dtype=packed.dtype, device=packed.device)
        dist.gather(packed, list(buf.unbind(0)), dst=dst, group=group)
        return unpack_packed(buf.flatten(0, 1))
    dist.gather(packed, None, dst=dst, group=group)
    return None
