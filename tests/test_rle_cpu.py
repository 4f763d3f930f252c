"""CPU: the oracle's restatement of cocoapi's mask run-length code (oracle/rle.py; parity unpinned: pycocotools is absent) against
hand-derived vectors and round trips, and the host-side string encoder of libape_b200.so against the oracle's."""
import numpy as np
import pytest

from oracle import rle as R


def test_hand_derived_vectors():
    # 2 x 2, column-major pixels 0,1,1,1 -> runs [1, 3]
    m = np.array([[0, 1], [1, 1]], np.uint8)
    assert R.encode_counts(m).tolist() == [1, 3]
    assert R.counts_to_string([1, 3]) == b"13"                       # one character per small count: chr(48 + c)
    # a mask that starts with a 1: the first (zero) run has length 0
    assert R.encode_counts(np.array([[1, 0], [0, 0]], np.uint8)).tolist() == [0, 1, 3]
    # all zeros / all ones
    assert R.encode_counts(np.zeros((3, 2), np.uint8)).tolist() == [6]
    assert R.encode_counts(np.ones((3, 2), np.uint8)).tolist() == [0, 6]
    # counts >= 16 need the continuation bit; from the fourth count on the difference to the count two places earlier is stored
    assert R.counts_to_string([37]) == bytes([48 + (5 | 0x20), 48 + 1])      # 37 = 1*32 + 5
    assert R.counts_to_string([5, 3, 7, 3]) == bytes([53, 51, 55, 48])       # 4th: 3 - 3 = 0
    assert R.counts_to_string([5, 3, 7, 1]) == bytes([53, 51, 55, 48 + 0x1E])  # 4th: 1 - 3 = -2 -> 0b11110, sign bit set, no continuation
    assert R.string_to_counts(b"13").tolist() == [1, 3]


@pytest.mark.parametrize("H,W,seed", [(1, 1, 0), (7, 5, 1), (64, 48, 2), (300, 200, 3), (1024, 768, 4)])
def test_round_trip(H, W, seed):
    g = np.random.default_rng(seed)
    yy, xx = np.mgrid[:H, :W]
    m = (((yy - H * g.random()) ** 2 + (xx - W * g.random()) ** 2) < (0.3 * max(H, W)) ** 2).astype(np.uint8)
    m ^= (g.random((H, W)) < 0.01).astype(np.uint8)  # speckle: many short runs
    rle = R.encode(m)
    assert rle["size"] == [H, W] and isinstance(rle["counts"], bytes)
    assert np.array_equal(R.decode(rle), m)
    counts = R.encode_counts(m)
    assert counts.sum() == H * W
    assert np.array_equal(R.string_to_counts(rle["counts"]), counts)


def test_library_string_encoder_equals_the_oracle(built):
    import ape_b200

    g = np.random.default_rng(9)
    for n in (0, 1, 3, 4, 1000):
        counts = g.integers(0, 5000, n).astype(np.uint32)
        if n > 5:
            counts[5] = 1 << 20  # a long run (four characters)
        assert ape_b200.ops.rle_counts_to_string(counts) == R.counts_to_string(counts)
