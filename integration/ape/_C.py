"""ape/_C.py — drop this file into the reference's `ape/` package (shenyunhang/APE) in place of the CUDAExtension that
`setup.py:41-108` builds from `ape/layers/csrc/`.

The reference loads its only native library with `from ape import _C`
(`ape/layers/multi_scale_deform_attn.py:415-423`; without it `MultiScaleDeformableAttention` becomes a dummy class that
raises ImportError even for `pytorch_attn=True`) and then calls
`torch.ops.ape.ms_deform_attn_forward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, im2col_step)`
(`:44-51`, `:322-332`), an operator registered by `TORCH_LIBRARY(ape, m)` in `ape/layers/csrc/vision.cpp:76-79`.

Importing `ape_b200` loads `libape_b200.so` (raising if it has not been built: there is no CPU / PyTorch fallback) and
registers the same two operator names with the same schemas through `torch.library`, forwarding to the C-ABI of
`include/ape_b200.h` (`ape_msda_fwd`).  Nothing else in the reference changes.
"""
import ape_b200  # noqa: F401  (side effect: torch.ops.ape.ms_deform_attn_forward / ms_deform_attn_backward)
from ape_b200 import ops as _ops

ms_deform_attn_forward = _ops.ms_deform_attn_forward
__all__ = ["ms_deform_attn_forward"]
