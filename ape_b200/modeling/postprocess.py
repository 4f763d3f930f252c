"""Panoptic merging of mask predictions (`DeformableDETRSegmVL._postprocess_panoptic`,
ape/modeling/ape_deta/deformable_detr_segm_vl.py:919-998), device-agnostic and without per-segment host round trips.

The reference walks the kept queries in a Python loop and calls `.item()` three times per query (areas of three masks).
Here the three areas of every query come from two `bincount`s over the per-pixel argmax, ONE small device->host copy
brings them over, the (inherently sequential, tiny) segment-id bookkeeping runs on those K-element arrays, and a lookup
table paints the segment ids.  Same decisions, same ids, same `segments_info`."""
from typing import Dict, Iterable, List, Tuple

import torch
import torch.nn.functional as F


def postprocess_panoptic(mask_cls: torch.Tensor, mask_pred: torch.Tensor, image_size: Tuple[int, int], height: int,
                         width: int, thing_ids: Iterable[int], num_thing_classes: int, stuff_first_is_things: bool,
                         cfg: Dict) -> Tuple[torch.Tensor, List[Dict]]:
    """mask_cls [K, N_t] class logits and mask_pred [K, H, W] mask logits (padded-image resolution) of the queries kept
    for one image -> (panoptic_seg int32 [height, width], segments_info).  cfg: prob, pano_temp, transform_eval,
    object_mask_threshold, overlap_threshold (the reference's `panoptic_configs`)."""
    prob, T = cfg["prob"], cfg["pano_temp"]
    # sem_seg_postprocess: crop to the unpadded size, bilinear resize to the output size
    m = mask_pred[:, : image_size[0], : image_size[1]].expand(1, -1, -1, -1)
    m = F.interpolate(m, size=(height, width), mode="bilinear", align_corners=False)[0]
    scores, labels = mask_cls.sigmoid().max(-1)
    m = m.sigmoid()
    keep = scores > cfg["object_mask_threshold"]
    if cfg["transform_eval"]:
        scores, labels = F.softmax(mask_cls.sigmoid() / T, dim=-1).max(-1)
    cur_scores, cur_classes, cur_masks = scores[keep], labels[keep], m[keep]
    K = int(cur_classes.shape[0])
    panoptic_seg = torch.zeros((height, width), dtype=torch.int32, device=m.device)
    if K == 0:
        return panoptic_seg, []
    cur_mask_ids = (cur_scores.view(-1, 1, 1) * cur_masks).argmax(0)                      # [h, w] winner per pixel
    winner_prob = torch.gather(cur_masks, 0, cur_mask_ids[None])[0]                        # mask prob of the winner
    solid = winner_prob >= prob                                                            # winner is also >= prob there
    mask_area = torch.bincount(cur_mask_ids.flatten(), minlength=K)                        # (cur_mask_ids == k).sum()
    inter_area = torch.bincount(cur_mask_ids[solid], minlength=K)                          # ((ids == k) & (m_k >= prob)).sum()
    original_area = (cur_masks >= prob).flatten(1).sum(1)                                  # (m_k >= prob).sum()
    stats = torch.stack([mask_area, original_area, inter_area, cur_classes.to(mask_area.dtype)]).cpu()  # the one D2H
    thing_ids = set(int(t) for t in thing_ids)
    lut = torch.zeros(K, dtype=torch.int32)
    segments_info, stuff_memory, current = [], {}, 0
    for k in range(K):
        area, orig, inter, pred_class = (int(v) for v in stats[:, k])
        isthing = pred_class in thing_ids
        if area > 0 and orig > 0 and inter > 0:
            if area / orig < cfg["overlap_threshold"]:
                continue
            if not isthing:
                if pred_class in stuff_memory:
                    lut[k] = stuff_memory[pred_class]
                    continue
                stuff_memory[pred_class] = current + 1
            current += 1
            lut[k] = current
            if not isthing and stuff_first_is_things:
                pred_class = pred_class - num_thing_classes + 1
            segments_info.append({"id": current, "isthing": bool(isthing), "category_id": int(pred_class)})
    painted = lut.to(m.device)[cur_mask_ids]
    panoptic_seg = torch.where(solid, painted, panoptic_seg)
    return panoptic_seg, segments_info
