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
