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


@pytest.mark.parametrize("n", [1, 63, 64, 65, 500, 5000, 20000])
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
