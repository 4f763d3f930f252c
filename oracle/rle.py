"""oracle/rle.py — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement of the run-length code the reference's evaluators hand to COCO-style tools:

    mask_util.encode(np.array(mask[:, :, None], order="F", dtype="uint8"))[0]
        ape/evaluation/d3_evaluation.py:466-468, refcoco_evaluation.py:450-452; detectron2 instances_to_coco_json
        (ape/evaluation/lvis_evaluation.py hands its predictions to it)
      -> pycocotools (cocoapi, a detectron2 dependency absent from /root/reference and from this image)
           common/maskApi.c: rleEncode (runs of equal pixels in column-major order, first run = zeros),
           rleToString (each count from the fourth on as the difference to the count two places earlier; 5 bits per character +
           continuation bit, offset 48), rleFrString / rleDecode.

**Parity unpinned**: pycocotools is not installed here, so the restatement is checked against hand-derived vectors and by
round trips (encode -> string -> parse -> decode = the mask) only; see tests/test_rle_cpu.py."""
import numpy as np


def encode_counts(mask):
    """rleEncode for one [H, W] mask: uint32 run lengths in column-major order, the first run counting zeros (may be 0)."""
    v = np.asarray(mask, dtype=np.uint8).T.reshape(-1)  # column-major pixel order
    if v.size == 0:
        return np.zeros((0,), np.uint32)
    change = np.flatnonzero(v[1:] != v[:-1]) + 1
    bounds = np.concatenate(([0], change, [v.size]))
    counts = np.diff(bounds)
    if v[0] != 0:
        counts = np.concatenate(([0], counts))
    return counts.astype(np.uint32)


def counts_to_string(counts):
    """rleToString."""
    out = bytearray()
    c = [int(x) for x in counts]
    for i, x in enumerate(c):
        if i > 2:
            x -= c[i - 2]
        more = True
        while more:
            ch = x & 0x1F
            x >>= 5
            more = (x != -1) if (ch & 0x10) else (x != 0)
            if more:
                ch |= 0x20
            out.append(ch + 48)
    return bytes(out)


def string_to_counts(s):
    """rleFrString."""
    counts, p = [], 0
    while p < len(s):
        x, k, more = 0, 0, True
        while more:
            c = s[p] - 48
            x |= (c & 0x1F) << (5 * k)
            more = bool(c & 0x20)
            p += 1
            k += 1
            if not more and (c & 0x10):
                x |= -1 << (5 * k)
        if len(counts) > 2:
            x += counts[-2]
        counts.append(x)
    return np.asarray(counts, dtype=np.int64)


def decode(rle):
    """rleDecode: {"size": [H, W], "counts": bytes} -> uint8 [H, W]."""
    H, W = rle["size"]
    counts = string_to_counts(rle["counts"])
    v = np.zeros((H * W,), np.uint8)
    pos, val = 0, 0
    for c in counts:
        v[pos:pos + int(c)] = val
        pos += int(c)
        val ^= 1
    return v.reshape(W, H).T


def encode(mask):
    """mask_util.encode(np.asfortranarray(mask)) for one [H, W] mask."""
    return {"size": [int(mask.shape[0]), int(mask.shape[1])], "counts": counts_to_string(encode_counts(mask))}
