"""GPU: LayerNorm and 2-D RoPE kernels against plain PyTorch fp32 references of the same ops."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def ops():
    import ape_b200

    return ape_b200.ops


@pytest.mark.parametrize("C", [256, 1024, 2730, 64, 4096])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
def test_layernorm(ops, C, dtype):
    g = torch.Generator().manual_seed(C)
    rows = 777
    pitch = (C + 7) // 8 * 8 + 16
    buf = torch.zeros(rows, pitch, dtype=dtype, device=DEV)
    buf[:, :C] = (torch.randn(rows, C, generator=g) * 2 + 0.5).to(dtype).to(DEV)
    buf[:, C:] = 7.0  # garbage in the padding must not influence the statistics
    x = buf[:, :C]
    w = (1 + 0.1 * torch.randn(C, generator=g)).to(DEV)
    b = (0.1 * torch.randn(C, generator=g)).to(DEV)
    out = torch.full((rows, pitch), 3.0, dtype=dtype, device=DEV)
    y = ops.layernorm(x, w, b, eps=1e-6, out=out[:, :C])
    want = F.layer_norm(x.float(), (C,), w, b, 1e-6)
    tol = {torch.float32: 1e-5, torch.float16: 4e-3, torch.bfloat16: 3e-2}[dtype]
    torch.testing.assert_close(y.float(), want, rtol=tol, atol=tol)
    pad_end = (C + 7) // 8 * 8
    assert (out[:, C:pad_end] == 0).all()  # alignment padding zeroed
    assert (out[:, pad_end:] == 3.0).all()  # nothing written beyond


def test_layernorm_row_map_and_dtype_change(ops):
    rows, C = 64, 256
    x = torch.randn(rows, C, device=DEV)
    w, b = torch.ones(C, device=DEV), torch.zeros(C, device=DEV)
    perm = torch.randperm(rows, device=DEV).to(torch.int32)
    y = ops.layernorm(x, w, b, out_dtype=torch.float16, row_map=perm)
    want = torch.empty_like(x)
    want[perm.long()] = F.layer_norm(x, (C,))
    torch.testing.assert_close(y.float(), want, rtol=2e-3, atol=2e-3)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
def test_rope_matches_reference_formula(ops, dtype):
    heads, hd, npos, M = 4, 64, 16, 48
    C = heads * hd
    g = torch.Generator().manual_seed(0)
    qkv = torch.randn(M, 3 * C, generator=g).to(dtype).to(DEV)
    ang = torch.randn(npos, hd, generator=g)
    cos, sin = ang.cos().to(DEV), ang.sin().to(DEV)
    pos = torch.randint(0, npos, (M,), generator=g).to(torch.int32).to(DEV)

    def ref(t, p):  # utils_eva02.py:248-252,346
        t = t.float().view(M, heads, hd)
        x = t.reshape(M, heads, hd // 2, 2)
        rot = torch.stack((-x[..., 1], x[..., 0]), -1).flatten(-2)
        return (t * cos[p][:, None] + rot * sin[p][:, None]).reshape(M, C)

    for pm in (pos, None):
        buf = qkv.clone()
        ops.rope_qk_(buf, cos, sin, C, hd, pos_map=pm)
        p = pos.long() if pm is not None else torch.arange(M, device=DEV) % npos
        tol = 1e-5 if dtype == torch.float32 else 4e-3
        torch.testing.assert_close(buf[:, :C].float(), ref(qkv[:, :C], p), rtol=tol, atol=tol)
        torch.testing.assert_close(buf[:, C:2 * C].float(), ref(qkv[:, C:2 * C], p), rtol=tol, atol=tol)
        assert torch.equal(buf[:, 2 * C:], qkv[:, 2 * C:])  # v untouched


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
@pytest.mark.parametrize("rows", [1, 100, 4096, 65536])
def test_groupnorm_nhwc(ops, dtype, rows):
    B, C, G = 2, 256, 32
    g = torch.Generator().manual_seed(rows)
    x = (torch.randn(B, rows, C, generator=g) * 1.5 + 0.3).to(dtype).to(DEV)
    w = (1 + 0.1 * torch.randn(C, generator=g)).to(DEV)
    b = (0.1 * torch.randn(C, generator=g)).to(DEV)
    y = ops.groupnorm_nhwc(x, w, b, G, eps=1e-5)
    want = F.group_norm(x.float().transpose(1, 2), G, w, b, 1e-5).transpose(1, 2)
    tol = 2e-5 if dtype == torch.float32 else 4e-3
    torch.testing.assert_close(y.float(), want, rtol=tol, atol=tol)
    assert torch.equal(y, ops.groupnorm_nhwc(x, w, b, G, eps=1e-5))  # deterministic


def test_groupnorm_writes_into_strided_slice(ops):
    """The neck writes each level straight into its slice of the flattened [B, S, C] tensor."""
    B, rows, C, G = 2, 300, 256, 32
    x = torch.randn(B, rows, C, device=DEV).half()
    w, b = torch.ones(C, device=DEV), torch.zeros(C, device=DEV)
    flat = torch.full((B, 1000, C), 9.0, dtype=torch.float16, device=DEV)
    ops.groupnorm_nhwc(x, w, b, G, out=flat[:, 200:500])
    assert torch.equal(flat[:, 200:500], ops.groupnorm_nhwc(x, w, b, G))
    assert (flat[:, :200] == 9).all() and (flat[:, 500:] == 9).all()


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
@pytest.mark.parametrize("second,col,row", [(False, False, False), (True, True, True), (False, True, True), (True, False, False)])
@pytest.mark.parametrize("C", [256, 1024, 64])
def test_layernorm_ex(ops, dtype, second, col, row, C):
    """ape_layernorm_ex = LN -> [LN] -> [+ per-image vector] -> (y, [y + row_add]) against the op sequence in fp32."""
    B, rows = 2, 333
    g = torch.Generator().manual_seed(C + second)
    x = (torch.randn(B, rows, C, generator=g) * 2 + 0.3).to(dtype).to(DEV)
    w1, b1, w2, b2 = [(1 + 0.1 * torch.randn(C, generator=g)).to(DEV) if i % 2 == 0 else (0.1 * torch.randn(C, generator=g)).to(DEV)
                      for i in range(4)]
    ca = torch.randn(B, C, generator=g).to(DEV) if col else None
    ra = torch.randn(B, rows, C, generator=g).to(dtype).to(DEV) if row else None
    y, y2 = ops.layernorm_ex(x, w1, b1, 1e-5, weight2=w2 if second else None, bias2=b2 if second else None, eps2=1e-6,
                             col_add=ca, row_add=ra)
    t = F.layer_norm(x.float(), (C,), w1, b1, 1e-5)
    if second:
        t = F.layer_norm(t, (C,), w2, b2, 1e-6)
    if col:
        t = t + ca[:, None]
    tol = {torch.float32: 2e-5, torch.float16: 4e-3, torch.bfloat16: 3e-2}[dtype]
    torch.testing.assert_close(y.float(), t, rtol=tol, atol=tol)
    if row:
        assert torch.equal(y2, (y.float() + ra.float()).to(dtype))  # exactly `y + row_add` on the stored tensors
    else:
        assert y2 is None


@pytest.mark.parametrize("B,N,K", [(1, 2312, 1024), (2, 1024, 2048), (4, 37, 256), (7, 300, 64)])
def test_gemv_f32_matches_linear(B, N, K):
    """ape_gemv_f32 (the folded language-side maps of VisionLanguageFusion) against F.linear in fp64."""
    import ape_b200

    g = torch.Generator().manual_seed(B * 1000 + N)
    x = torch.randn(B, 1, K, generator=g).cuda()
    w = (torch.randn(N, K, generator=g) * K ** -0.5).cuda()
    b = torch.randn(N, generator=g).cuda()
    n0 = ape_b200._lib.launch_count()
    y = ape_b200.ops.gemv_f32(x, w, b)
    assert ape_b200._lib.launch_count() - n0 == (B + 3) // 4
    want = torch.nn.functional.linear(x.double(), w.double(), b.double()).float()
    assert y.shape == (B, 1, N)
    torch.testing.assert_close(y, want, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(ape_b200.ops.gemv_f32(x, w), want - b, rtol=1e-5, atol=1e-5)


def test_ref_update_equals_the_pytorch_sequence_bit_for_bit():
    """ape_ref_update against the decoder's elementwise sequence (deformable_transformer_vl.py:268-300; detrex inverse_sigmoid)."""
    import ape_b200
    from ape_b200.layers.common import inverse_sigmoid

    g = torch.Generator().manual_seed(5)
    B, Q, L = 2, 900, 5
    ref = torch.rand(B, Q, 4, generator=g).cuda()
    ref[0, :5] = torch.tensor([0.0, 1.0, 1e-4, 0.9995])  # the clamps of inverse_sigmoid bite
    delta = (torch.randn(B, Q, 4, generator=g) * 2).cuda()
    vr = (0.5 + 0.5 * torch.rand(B, L, 2, generator=g)).cuda()
    new_ref, ref_in = ape_b200.ops.ref_update(delta, ref, vr)
    want = (delta + inverse_sigmoid(ref)).sigmoid()
    want_in = want[:, :, None] * torch.cat([vr, vr], -1)[:, None]
    assert torch.equal(new_ref, want), (new_ref - want).abs().max().item()
    assert torch.equal(ref_in, want_in)
