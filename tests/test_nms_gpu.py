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


@pytest.mark.parametrize("Q,N,thresh", [(900, 1203, 0.0), (300, 80, 0.3), (33, 7, 0.0), (1024, 64, 0.5), (5, 3, 2.0)])
def test_classwise_nms_equals_per_class_torchvision(ops, Q, N, thresh):
    """ops.nms_classwise (one shared Q x Q IoU matrix, one warp per class) against torchvision's per-class loop
    (`_batched_nms_vanilla`, what batched_nms runs above 25 000 boxes on CUDA): identical surviving (query, class) sets."""
    g = torch.Generator().manual_seed(Q * 7 + N)
    c = torch.rand(Q, 2, generator=g) * 600
    wh = torch.rand(Q, 2, generator=g) * 200 + 10
    boxes = torch.cat([c - wh / 2, c + wh / 2], 1).to(DEV).contiguous()
    scores = torch.rand(Q, N, generator=g).to(DEV)
    valid = (torch.rand(Q, generator=g) > 0.05).to(DEV)
    got = ops.nms_classwise(boxes, scores, thresh, 0.7, row_valid=valid.to(torch.uint8))   # [N, Q]
    assert got.shape == (N, Q)
    want = torch.full((N, Q), float("-inf"), device=DEV)
    for cls in range(N):
        cand = ((scores[:, cls] > thresh) & valid).nonzero()[:, 0]
        if cand.numel():
            keep = cand[torchvision.ops.nms(boxes[cand], scores[cand, cls], 0.7)]
            want[cls, keep] = scores[keep, cls]
    assert torch.equal(got, want)


def test_inference_threshold_zero_full_vocabulary_is_bounded():
    """test_score_thresh 0.0 with 900 queries x 1203 names = 1.08 M candidates must not allocate an n x n
    IoU matrix; result = top-k of the per-class NMS survivors, as the reference's batched_nms (vanilla branch)."""
    from ape_b200 import configs
    from ape_b200.layers.common import box_cxcywh_to_xyxy
    from ape_b200.modeling import build_model

    spec = dict(configs.MINI)
    m = build_model(spec).to(DEV)
    m.test_score_thresh, m.test_topk_per_image = 0.0, 300
    g = torch.Generator().manual_seed(5)
    Q, N = 900, 1203
    box_cls = (torch.randn(1, Q, N, generator=g) * 2).to(DEV)
    cxcy = torch.rand(1, Q, 2, generator=g) * 0.8 + 0.1
    wh = torch.rand(1, Q, 2, generator=g) * 0.3 + 0.02
    box_pred = torch.cat([cxcy, wh], -1).to(DEV)
    torch.cuda.reset_peak_memory_stats()
    base = torch.cuda.memory_allocated()
    res = m._inference_static(box_cls, box_pred, [(800, 1000)])[0]
    assert torch.cuda.max_memory_allocated() - base < 200 * 2 ** 20
    assert len(res) == 300
    # literal per-class reference on the same tensors
    scores = box_cls[0].sigmoid()
    boxes = box_cxcywh_to_xyxy(box_pred[0]) * torch.tensor([1000, 800, 1000, 800], device=DEV)
    boxes = torch.stack((boxes[:, 0].clamp(0, 1000), boxes[:, 1].clamp(0, 800), boxes[:, 2].clamp(0, 1000), boxes[:, 3].clamp(0, 800)), -1)
    surv = torch.full((Q, N), float("-inf"), device=DEV)
    for cls in range(N):
        keep = torchvision.ops.nms(boxes, scores[:, cls], m.test_nms_thresh)
        surv[keep, cls] = scores[keep, cls]
    topv, topi = torch.topk(surv.flatten(), 300)
    assert torch.equal(res.scores, topv.cpu())
    # float32 sigmoid values tie now and then, and top-k orders ties arbitrarily: every returned (query, class) must be a
    # survivor carrying exactly the score reported for it, no pair twice, and equal to top-k wherever the score is unique
    got = res.query_index * N + res.pred_classes
    assert got.unique().numel() == 300
    assert torch.equal(surv.flatten().cpu()[got], res.scores)
    uniq = torch.ones(300, dtype=torch.bool)
    uniq[1:] &= topv.cpu()[1:] != topv.cpu()[:-1]
    uniq[:-1] &= topv.cpu()[1:] != topv.cpu()[:-1]
    assert torch.equal(got[uniq], topi.cpu()[uniq]) and int(uniq.sum()) > 200
