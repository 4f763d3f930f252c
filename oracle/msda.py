"""oracle/msda.py — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU oracles for multi-scale deformable attention forward:

* `msda_c`      ctypes front-end of oracle/msda_ref.c (restates the reference CUDA kernel
                ape/layers/csrc/MsDeformAttn/ms_deform_im2col_cuda.cuh:237-299, :33-84).
* `msda_torch`  restatement of the reference's portable path
                ape/layers/multi_scale_deform_attn.py:84-124 (per-level F.grid_sample, bilinear,
                zeros padding, align_corners=False, then the attention-weighted sum).
* `ref_cuda`    the reference's own CUDA kernel (oracle/_ref/libref_msda.so, built from
                /root/reference by oracle/Makefile) — GPU box only, second oracle + kernel to beat.

Pinned by tests/test_oracle_golden.py against vectors generated from the reference's own
`multi_scale_deformable_attn_pytorch` (tests/golden/gen_msda_golden.py)."""
import ctypes
import os
import subprocess

import torch
import torch.nn.functional as F

_HERE = os.path.dirname(os.path.abspath(__file__))
_C_LIB = os.path.join(_HERE, "_build", "libmsda_oracle.so")
_REF_LIB = os.path.join(_HERE, "_ref", "libref_msda.so")
_c = None
_ref = None


def build():
    """Compile the C restatement (and, when /root/reference exists, oracle/_ref)."""
    subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)


def _load_c():
    global _c
    if _c is None:
        if not os.path.exists(_C_LIB):
            build()
        _c = ctypes.CDLL(_C_LIB)
        _c.msda_ref_forward.restype = ctypes.c_int
        _c.msda_ref_forward.argtypes = [ctypes.c_void_p] * 6 + [ctypes.c_int] * 9
        _c.msda_ref_max_threads.restype = ctypes.c_int
    return _c


def max_threads() -> int:
    return int(_load_c().msda_ref_max_threads())


def level_start_index(spatial_shapes: torch.Tensor) -> torch.Tensor:
    """deformable_transformer_vl.py:459-461: cat(0, cumsum(H_l*W_l)[:-1])."""
    areas = spatial_shapes[:, 0] * spatial_shapes[:, 1]
    return torch.cat([areas.new_zeros(1), areas.cumsum(0)[:-1]])


def msda_c(value, spatial_shapes, level_start, loc, attn, acc_double=True, nthreads=0):
    """All tensors on CPU; computes in fp32 inputs (casts), returns fp32 [B,Q,H*D]."""
    lib = _load_c()
    v = value.detach().to("cpu", torch.float32).contiguous()
    lo = loc.detach().to("cpu", torch.float32).contiguous()
    at = attn.detach().to("cpu", torch.float32).contiguous()
    sh = spatial_shapes.detach().to("cpu", torch.int64).contiguous()
    st = level_start.detach().to("cpu", torch.int64).contiguous()
    B, S, H, D = v.shape
    _, Q, _, L, P, _ = lo.shape
    out = torch.empty((B, Q, H * D), dtype=torch.float32)
    rc = lib.msda_ref_forward(v.data_ptr(), sh.data_ptr(), st.data_ptr(), lo.data_ptr(), at.data_ptr(),
                              out.data_ptr(), B, S, H, D, L, Q, P, 1 if acc_double else 0, int(nthreads))
    assert rc == 0
    return out


def msda_torch(value, spatial_shapes, loc, attn):
    """multi_scale_deform_attn.py:84-124 restated: works on any device/dtype torch supports."""
    B, _, H, D = value.shape
    _, Q, _, L, P, _ = loc.shape
    sizes = [int(h) * int(w) for h, w in spatial_shapes.tolist()]
    per_level = value.split(sizes, dim=1)
    grids = 2 * loc - 1  # :94  [0,1] -> [-1,1]
    sampled = []
    for lvl, (h, w) in enumerate(spatial_shapes.tolist()):
        # (B, h*w, H, D) -> (B*H, D, h, w)                                   :101-103
        v = per_level[lvl].flatten(2).transpose(1, 2).reshape(B * H, D, int(h), int(w))
        # (B, Q, H, P, 2) -> (B*H, Q, P, 2)                                   :107
        g = grids[:, :, :, lvl].transpose(1, 2).flatten(0, 1)
        sampled.append(F.grid_sample(v, g, mode="bilinear", padding_mode="zeros", align_corners=False))
    a = attn.transpose(1, 2).reshape(B * H, 1, Q, L * P)  # :116-118
    out = (torch.stack(sampled, dim=-2).flatten(-2) * a).sum(-1).view(B, H * D, Q)
    return out.transpose(1, 2).contiguous()


def have_ref_cuda() -> bool:
    return os.path.exists(_REF_LIB)


def ref_cuda(value, spatial_shapes, level_start, loc, attn, im2col_step=64):
    """Run the reference's own CUDA kernel (fp32 / fp16) on CUDA tensors."""
    global _ref
    if _ref is None:
        _ref = ctypes.CDLL(_REF_LIB)
        _ref.ref_msda_forward.restype = ctypes.c_int
        _ref.ref_msda_forward.argtypes = [ctypes.c_void_p] * 6 + [ctypes.c_int] * 9 + [ctypes.c_void_p]
    code = {torch.float32: 0, torch.float16: 1}[value.dtype]
    B, S, H, D = value.shape
    _, Q, _, L, P, _ = loc.shape
    out = torch.zeros((B, Q, H * D), dtype=value.dtype, device=value.device)  # at::zeros, ms_deform_attn_cuda.cu:55
    rc = _ref.ref_msda_forward(value.data_ptr(), spatial_shapes.data_ptr(), level_start.data_ptr(),
                               loc.data_ptr(), attn.data_ptr(), out.data_ptr(), B, S, H, D, L, Q, P, code,
                               int(im2col_step), torch.cuda.current_stream().cuda_stream)
    if rc != 0:
        raise RuntimeError(f"reference kernel launch failed: {rc}")
    return out


def make_inputs(B, Q, H, D, shapes, P, seed=3, border=False, dtype=torch.float32, device="cpu"):
    """SURVEY.md §8(d) synthetic tensors: value~N(0,1), loc~U(0,1) (border: U(-0.1,1.1)),
    attn = softmax(N(0,1)) over L*P."""
    g = torch.Generator().manual_seed(seed)
    ss = torch.tensor(shapes, dtype=torch.int64)
    L = ss.shape[0]
    S = int((ss[:, 0] * ss[:, 1]).sum())
    value = torch.randn(B, S, H, D, generator=g)
    loc = torch.rand(B, Q, H, L, P, 2, generator=g)
    if border:
        loc = loc * 1.2 - 0.1
    attn = torch.randn(B, Q, H, L * P, generator=g).softmax(-1).view(B, Q, H, L, P)
    st = level_start_index(ss)
    return (value.to(device, dtype), ss.to(device), st.to(device), loc.to(device, dtype), attn.to(device, dtype))
