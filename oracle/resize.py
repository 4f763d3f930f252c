"""oracle/resize.py — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement (numpy) of the image resize of the reference's predictor:

    DefaultPredictor.__call__                 ape/engine/defaults.py:216-222
      -> ResizeShortestEdge.get_transform     detectron2 @ 017abbf (requirements.txt:10), data/transforms/augmentation_impl.py
      -> ResizeTransform.apply_image          detectron2 data/transforms/transform.py: uint8 -> PIL Image.resize(BILINEAR)
      -> Pillow (a detectron2 dependency, absent from /root/reference; Pillow 12.2.0 is installed in this image)
           libImaging/Resample.c: precompute_coeffs, normalize_coeffs_8bpc, ImagingResampleHorizontal_8bpc,
           ImagingResampleVertical_8bpc, clip8 — triangle filter of support max(scale, 1), taps normalised in double and
           rounded to 22-bit fixed point, accumulator started at 1 << 21, `>> 22` and clamp to [0, 255], a uint8 image
           between the horizontal and the vertical pass.

Pinned by tests/test_preprocess_cpu.py against PIL itself (the implementation the reference runs) on up-/down-scaling, identity,
odd sizes, 1 and 3 channels: bit-exact."""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def coeffs(in_size, out_size):
    """precompute_coeffs + normalize_coeffs_8bpc for the box (0, in_size): bounds int32 [out,2], taps int32 [out,ksize]."""
    in0, in1 = np.float32(0.0), np.float32(in_size)
    filterscale = scale = float(in1 - in0) / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = float(in0) + (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        x = np.arange(xmax, dtype=np.float64)
        w = np.abs((x + xmin - center + 0.5) * ss)
        w = np.where(w < 1.0, 1.0 - w, 0.0)
        ww = 0.0
        for v in w:  # Pillow sums the taps in order
            ww += v
        if ww != 0.0:
            w = w / ww
        fixed = np.where(w < 0, -0.5 + w * (1 << PRECISION_BITS), 0.5 + w * (1 << PRECISION_BITS))
        kk[xx, :xmax] = np.trunc(fixed).astype(np.int32)
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def _pass(img, bounds, kk, axis):
    """One 8bpc pass along `axis` (0 = vertical, 1 = horizontal) of a uint8 [H,W,C] image."""
    img = np.moveaxis(img, axis, 0).astype(np.int64)
    out = np.empty((bounds.shape[0],) + img.shape[1:], np.uint8)
    for o in range(bounds.shape[0]):
        lo, n = int(bounds[o, 0]), int(bounds[o, 1])
        acc = (1 << (PRECISION_BITS - 1)) + np.tensordot(kk[o, :n].astype(np.int64), img[lo: lo + n], axes=(0, 0))
        out[o] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def resize_u8(img, new_h, new_w):
    """`np.asarray(Image.fromarray(img).resize((new_w, new_h), Image.BILINEAR))` for uint8 [H,W] / [H,W,C]."""
    two_d = img.ndim == 2
    x = img[:, :, None] if two_d else img
    H, W = x.shape[:2]
    if new_w != W:  # ImagingResample: horizontal pass first, skipped when the width is unchanged
        x = _pass(x, *coeffs(W, new_w), axis=1)
    if new_h != H:
        x = _pass(x, *coeffs(H, new_h), axis=0)
    x = np.ascontiguousarray(x)
    return x[:, :, 0] if two_d else x


def get_output_shape(oldh, oldw, short_edge_length, max_size):
    """ResizeShortestEdge.get_output_shape (detectron2 augmentation_impl.py)."""
    h, w = oldh, oldw
    size = short_edge_length * 1.0
    scale = size / min(h, w)
    if h < w:
        newh, neww = size, scale * w
    else:
        newh, neww = scale * h, size
    if max(newh, neww) > max_size:
        scale = max_size * 1.0 / max(newh, neww)
        newh, neww = newh * scale, neww * scale
    return int(newh + 0.5), int(neww + 0.5)


def predictor_image(original_image, short_edge_length, max_size, input_format="RGB"):
    """defaults.py:216-222: the float32 CHW array the reference's predictor hands to the model for a BGR uint8 image."""
    if input_format == "RGB":
        original_image = original_image[:, :, ::-1]
    h, w = original_image.shape[:2]
    nh, nw = get_output_shape(h, w, short_edge_length, max_size)
    image = resize_u8(np.ascontiguousarray(original_image), nh, nw)
    return image.astype("float32").transpose(2, 0, 1)
