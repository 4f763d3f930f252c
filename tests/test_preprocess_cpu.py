"""CPU: the predictor's resize.  The oracle's restatement of Pillow's bilinear resample is pinned
against PIL itself — the implementation the reference's predictor runs (ape/engine/defaults.py:221 -> detectron2
ResizeTransform.apply_image -> PIL Image.resize) — and the host-side tap tables of libape_b200.so against the oracle's."""
import numpy as np
import pytest
import torch

from oracle import resize as R

PIL_Image = pytest.importorskip("PIL.Image")

CASES = [  # (H, W, C, new_h, new_w)
    (37, 53, 3, 74, 106),      # exact 2x up
    (64, 48, 3, 100, 75),      # fractional up
    (301, 200, 3, 77, 51),     # ~3.9x down (wide filter)
    (120, 160, 3, 120, 160),   # identity
    (50, 70, 3, 50, 33),       # one axis only
    (33, 41, 1, 90, 17),       # single channel, mixed
    (480, 640, 3, 768, 1024),  # the golden image geometry
    (7, 5, 3, 1, 1),           # degenerate
]


def _img(H, W, C, seed):
    g = np.random.default_rng(seed)
    x = g.integers(0, 256, (H, W, C), dtype=np.uint8)
    x[: H // 3, : W // 3] = 255  # saturated and flat regions: rounding at the clip boundaries
    x[-(H // 4 + 1):, -(W // 4 + 1):] = 0
    return x


@pytest.mark.parametrize("H,W,C,nh,nw", CASES)
def test_oracle_resize_is_pil_bit_for_bit(H, W, C, nh, nw):
    img = _img(H, W, C, seed=H * 1000 + W)
    src = img[:, :, 0] if C == 1 else img
    want = np.asarray(PIL_Image.fromarray(src).resize((nw, nh), PIL_Image.BILINEAR))
    got = R.resize_u8(src, nh, nw)
    assert got.shape == want.shape and got.dtype == np.uint8
    assert np.array_equal(got, want), f"max |diff| = {np.abs(got.astype(int) - want.astype(int)).max()}"


@pytest.mark.parametrize("in_size,out_size", [(53, 106), (48, 75), (301, 77), (160, 160), (640, 1024), (5, 1), (1, 9), (1333, 800)])
def test_library_tap_tables_equal_the_oracle(built, in_size, out_size):
    import ape_b200

    lib = ape_b200._lib.lib
    bounds, kk = R.coeffs(in_size, out_size)
    ksize = lib.ape_resample_ksize(in_size, out_size)
    assert ksize == kk.shape[1]
    b = torch.empty((out_size, 2), dtype=torch.int32)
    k = torch.empty((out_size, ksize), dtype=torch.int32)
    assert lib.ape_resample_coeffs_u8(in_size, out_size, b.data_ptr(), k.data_ptr()) == 0
    assert np.array_equal(b.numpy(), bounds)
    assert np.array_equal(k.numpy(), kk)
    assert (k.sum(1) - (1 << R.PRECISION_BITS)).abs().max() <= kk.shape[1]  # taps sum to one up to rounding


def test_bad_arguments_are_rejected(built):
    import ape_b200

    lib = ape_b200._lib.lib
    assert lib.ape_resample_ksize(0, 5) < 0
    assert lib.ape_resample_coeffs_u8(4, 4, None, None) < 0
    assert b"null" in lib.ape_last_error()
    with pytest.raises(RuntimeError, match="CUDA uint8"):
        ape_b200.ops.resize_u8_bilinear(torch.zeros(4, 4, 3, dtype=torch.uint8), 8, 8)  # no CPU path


def test_get_output_shape_matches_detectron2_known_answers():
    from ape_b200.engine import ResizeShortestEdge

    # detectron2's documented behaviour: shorter side -> size, longer side capped at max_size, int(x + 0.5)
    known = {(480, 640, 800, 1333): (800, 1067), (640, 480, 800, 1333): (1067, 800), (500, 1500, 800, 1333): (444, 1333),
             (512, 512, 1024, 1024): (1024, 1024), (768, 1024, 1024, 1024): (768, 1024), (427, 640, 1024, 1024): (683, 1024)}
    for (h, w, s, m), want in known.items():
        assert ResizeShortestEdge.get_output_shape(h, w, s, m) == want
        assert R.get_output_shape(h, w, s, m) == want
    t = ResizeShortestEdge(1024, 1024).get_transform(np.zeros((427, 640, 3), np.uint8))
    assert (t.h, t.w, t.new_h, t.new_w) == (427, 640, 683, 1024)
    assert ResizeShortestEdge((0, 0), 1024, "choice").get_transform(np.zeros((8, 9, 3), np.uint8)).new_w == 9  # size 0: no-op
