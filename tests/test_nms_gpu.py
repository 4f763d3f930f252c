"""GPU: the NMS kernels against torchvision.ops.nms / batched_nms (same kept indices, same order)."""
import pytest
import torch
import torchvision

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def ops():
    import ape_b200

    return ape_b200.ops


def boxes_like_proposals(n, seed, spread=1.0):
    g = torch.Generator().manual_seed(seed)
    c = torch.rand(n, 2, generator=g) * spread
    wh = torch.rand(n, 2, generator=g) * 0.2 + 0.02
    return torch.cat([c - wh / 2, c + wh / 2], 1).to(DEV), torch.rand(n, generator=g).to(DEV)


@pytest.mark.parametrize("n", [1, 63, 64, 65, 127, 500, 5000, 12600, 12700, 20000])
@pytest.mark.parametrize("thr", [0.7, 0.9])
def test_nms_equals_torchvision(ops, n, thr):
    b, s = boxes_like_proposals(n, n)
    want = torchvision.ops.nms(b, s, thr)
    got = ops.nms(b, s, thr)
    assert torch.equal(got, want)


def test_batched_nms_equals_torchvision(ops):
    b, s = boxes_like_proposals(5000, 7)
    idx = torch.randint(0, 5, (5000,), device=DEV)
    assert torch.equal(ops.batched_nms(b, s, idx, 0.9), torchvision.ops.boxes.batched_nms(b, s, idx, 0.9))
    # many heavily overlapping boxes of one class (everything suppressed by the first)
    b2 = b[:1].repeat(300, 1)
    assert torch.equal(ops.nms(b2, s[:300], 0.5), torchvision.ops.nms(b2, s[:300], 0.5))
    assert ops.nms(b[:0], s[:0], 0.5).numel() == 0


@pytest.mark.parametrize("seed", [0, 1])
def test_static_selection_equals_reference_loop(seed):
    """transformer.select_proposals (static shapes, graph capturable) against a literal transcription of the
    reference's per-level loop (deformable_transformer_vl.py:569-625) on torchvision's batched_nms."""
    from ape_b200 import configs
    from ape_b200.layers.common import box_cxcywh_to_xyxy
    from ape_b200.modeling import build_model

    spec = dict(configs.MINI)
    spec.update(num_queries=300)
    tr = build_model(spec).transformer.to(DEV)
    shapes = [(64, 64), (32, 32), (16, 16), (8, 8), (4, 4)]
    S = sum(h * w for h, w in shapes)
    g = torch.Generator().manual_seed(seed)
    logit = torch.randn(S, generator=g).to(DEV)
    coord = torch.randn(S, 4, generator=g).to(DEV) * (0.3 if seed else 1.5)
    level_ids = torch.cat([torch.full((h * w,), i, dtype=torch.long) for i, (h, w) in enumerate(shapes)]).to(DEV)
    got = tr.select_proposals(logit, coord, level_ids, 5)

    topk, pre_k = tr.two_stage_num_proposals, tr.pre_nms_topk
    boxes = box_cxcywh_to_xyxy(coord.sigmoid()).clamp(0, 1)
    pre = []
    for lvl in range(5):
        order = torch.sort(logit.sigmoid() * (level_ids == lvl), descending=True, stable=True)[1]
        pre.append(order[: min(pre_k, S)])
    pre = torch.cat(pre)
    post = torchvision.ops.boxes.batched_nms(boxes[pre].float(), logit[pre].float(), level_ids[pre], tr.nms_thresh_enc)
    keep = pre[post]
    if len(keep) < topk:
        keep = torch.sort(logit, descending=True, stable=True)[1][:topk]
    q_per_l = topk // 5
    is_lvl = level_ids[keep][None] == torch.arange(5, device=DEV)[:, None]
    km = (is_lvl & (is_lvl.cumsum(1) <= q_per_l)).any(0)
    if km.sum() < topk:
        pad = (~km).nonzero()[: topk - km.sum()]
        km[pad] = True
    assert torch.equal(got, keep[km])
