"""GPU: ms_deform_attn_backward through the C-ABI (ape_msda_bwd) and through the reference's operator name, against
PyTorch autograd in float64 through the oracle's restatement of `multi_scale_deformable_attn_pytorch`
(oracle/msda.py:msda_torch; the reference's own CUDA-vs-PyTorch check, ape/layers/csrc tests, compares exactly these)."""
import pytest
import torch

from oracle import msda as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def autograd_reference(value, ss, loc, attn, grad_out):
    v = value.double().cpu().requires_grad_(True)
    lo = loc.double().cpu().requires_grad_(True)
    at = attn.double().cpu().requires_grad_(True)
    out = O.msda_torch(v, ss.cpu(), lo, at)
    out.backward(grad_out.double().cpu())
    return v.grad, lo.grad, at.grad


@pytest.mark.parametrize("shapes,B,Q,H,D,P", [
    ([(12, 17), (6, 9), (3, 5)], 2, 77, 8, 32, 4),
    ([(8, 8)], 1, 5, 4, 16, 2),
    ([(20, 14), (10, 7)], 2, 300, 8, 32, 4),
    ([(9, 9), (5, 5), (3, 3), (2, 2), (1, 1)], 1, 33, 2, 64, 3),
])
@pytest.mark.parametrize("border", [False, True])
def test_backward_matches_float64_autograd(shapes, B, Q, H, D, P, border):
    import ape_b200

    value, ss, st, loc, attn = O.make_inputs(B, Q, H, D, shapes, P, seed=31, border=border)
    g = torch.Generator().manual_seed(1)
    grad_out = torch.randn(B, Q, H * D, generator=g)
    want = autograd_reference(value, ss, loc, attn, grad_out)
    got = ape_b200.ops.ms_deform_attn_backward(value.to(DEV), ss.to(DEV), st.to(DEV), loc.to(DEV), attn.to(DEV), grad_out.to(DEV))
    for name, a, b in zip(("grad_value", "grad_sampling_loc", "grad_attn_weight"), got, want):
        scale = b.abs().max().item() + 1e-6
        err = (a.double().cpu() - b).abs().max().item()
        assert err < 2e-5 * scale + 1e-6, f"{name}: max|err| {err:.3e} on scale {scale:.3e}"


def test_backward_through_the_reference_operator_and_16bit():
    """torch.ops.ape.ms_deform_attn_backward (the schema the reference registers, vision.cpp:78) returns the three
    gradients; fp16 inputs are computed in fp32 (the reference computes and atomically accumulates in half)."""
    import ape_b200  # noqa: F401

    shapes = [(16, 16), (8, 8), (4, 4)]
    value, ss, st, loc, attn = O.make_inputs(2, 200, 8, 32, shapes, 4, seed=7)
    grad_out = torch.randn(2, 200, 256, generator=torch.Generator().manual_seed(2))
    want = autograd_reference(value.half().float(), ss, loc.half().float(), attn.half().float(), grad_out.half().float())
    outs = torch.ops.ape.ms_deform_attn_backward(value.to(DEV).half(), ss.to(DEV), st.to(DEV), loc.to(DEV).half(),
                                                 attn.to(DEV).half(), grad_out.to(DEV).half(), 64)
    assert len(outs) == 3 and all(o.dtype == torch.float16 for o in outs)
    for a, b in zip(outs, want):
        scale = b.abs().max().item()
        assert (a.double().cpu() - b).abs().max().item() < 4e-3 * scale


def test_backward_full_encoder_shape_properties():
    """APE-L_D 1024^2 encoder shape: linearity in grad_out and sum(grad_value) = sum over samples of weights * grad_out."""
    import ape_b200

    shapes = [(256, 256), (128, 128), (64, 64), (32, 32), (16, 16)]
    S = sum(h * w for h, w in shapes)
    value, ss, st, loc, attn = (t.to(DEV) for t in O.make_inputs(1, S, 8, 32, shapes, 4, seed=3))
    g1 = torch.randn(1, S, 256, device=DEV)
    a = ape_b200.ops.ms_deform_attn_backward(value, ss, st, loc, attn, g1)
    b = ape_b200.ops.ms_deform_attn_backward(value, ss, st, loc, attn, 2 * g1)
    for x, y in zip(a, b):
        torch.testing.assert_close(y, 2 * x, rtol=2e-4, atol=1e-4)
    # d/dvalue of <out, g>: perturbing value by a constant c changes <out, g> by c * sum(grad_value)
    out0 = ape_b200.ops.ms_deform_attn_forward(value, ss, st, loc, attn)
    out1 = ape_b200.ops.ms_deform_attn_forward(value + 0.5, ss, st, loc, attn)
    lhs = ((out1 - out0).double() * g1.double()).sum().item()
    rhs = 0.5 * a[0].double().sum().item()
    assert abs(lhs - rhs) < 1e-3 * max(1.0, abs(rhs))
