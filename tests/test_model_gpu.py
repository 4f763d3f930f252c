"""GPU parity of the engine's full detection forward (ape_b200.modeling) against
(1) golden tensors recorded from the REFERENCE model on the MINI spec, and
(2) the oracle's CPU port (oracle/ape_forward.py) at MINI and at the full APE-L_D 1024^2 size.
Weights: name-derived synthetic (oracle/synth.py); fp32, TF32 off."""
import pytest
import torch

from conftest import load_golden
from ape_b200 import configs
from oracle import ape_forward as AF
from oracle import synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

CASES = {
    "single": ([(48, 64, 96, 128)], None),
    "batch2": ([(64, 64, 64, 64), (40, 56, 80, 112)], None),
    "phrase": ([(64, 48, 64, 48)], "red apple,a dog on grass,tall tree"),
    "expression": ([(64, 56, 128, 112)], ["the red apple on the left", "a dog"]),
    "maskprompt": ([(64, 64, 64, 64)], None, (8, 40, 16, 56)),
}


def _build(spec):
    from ape_b200.modeling import build_model

    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    m = build_model(spec)
    synth.fill_state_dict(m)
    return m


@pytest.fixture(scope="module")
def mini():
    m = _build(configs.MINI)
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    return m.to(DEV), sd


@pytest.mark.parametrize("case", sorted(CASES))
def test_mini_matches_reference_golden(case, mini):
    model, _ = mini
    sizes, text = CASES[case][:2]
    rect = CASES[case][2] if len(CASES[case]) > 2 else None
    g = load_golden(f"model_mini_{case}.npz")
    inputs = []
    for i, (h, w, oh, ow) in enumerate(sizes):
        d = {"image": synth.image(h, w, seed=i), "height": oh, "width": ow}
        if rect is not None:
            mp = torch.zeros(h, w)
            mp[rect[0]:rect[1], rect[2]:rect[3]] = 1.0
            d["mask_prompt"] = mp
        if isinstance(text, list):
            d.update(prompt="expression", expressions=list(text))
        elif text:
            d.update(prompt="text", text_prompt=text)
        inputs.append(d)
    out = model(inputs)
    lo = model.last_outputs
    tol = dict(rtol=2e-3, atol=2e-3)
    for k in ("p2", "p3", "p4", "p5", "p6"):
        torch.testing.assert_close(lo["features"][k][:, ::4].cpu(), g[f"backbone.{k}"], **tol)
    for i in range(5):
        torch.testing.assert_close(lo["neck"][i][:, ::8].cpu(), g[f"neck.{i}"], **tol)
    # (fp32 GPU vs fp32 CPU reduction orders; one element in 22 k reaches 2.1e-3 on the phrase case)
    torch.testing.assert_close(lo["memory"][:, ::4].cpu(), g["memory"], rtol=5e-3, atol=5e-3)
    # bit-exact: proposal indices selected by top-k + NMS + per-level quota.  On the toy spec every level has
    # fewer tokens than pre_nms_topk, so the reference's per-level torch.topk(sigmoid*level_mask) also returns
    # zero-score entries whose order is implementation-defined (CPU vs CUDA top-k); those padded / invalid
    # proposals (infinite coordinates, identical features) are compared as a multiset of coordinates instead.
    sel, want = model.transformer.last_topk_proposals.cpu(), g["topk_proposals"]
    valid = torch.isfinite(g["init_reference"]).all(-1) & (g["init_reference"] < 1).all(-1)
    assert torch.equal(sel[valid], want[valid])
    assert sel.shape == want.shape
    torch.testing.assert_close(lo["init_reference"].cpu(), g["init_reference"], **tol)
    torch.testing.assert_close(lo["inter_states"].cpu(), g["inter_states"], rtol=5e-3, atol=5e-3)
    torch.testing.assert_close(lo["inter_references"].cpu(), g["inter_references"], **tol)
    for b, o in enumerate(out):
        inst = o["instances"]
        assert inst.pred_boxes.tensor.device.type == "cpu"  # results are handed back on the host like the reference
        assert torch.equal(inst.pred_classes, g[f"det{b}.classes"])  # bit-exact kept (query, class) pairs
        torch.testing.assert_close(inst.scores, g[f"det{b}.scores"], rtol=1e-3, atol=1e-5)
        torch.testing.assert_close(inst.pred_boxes.tensor, g[f"det{b}.boxes"], rtol=1e-3, atol=2e-2)


def test_mini_masks_and_semantic_match_reference_golden(mini):
    """mask head, instance-mask post-processing and the semantic branch against
    tensors recorded from the reference model (tests/golden/gen_model_golden.py masks)."""
    import numpy as np

    model, _ = mini
    g = load_golden("model_mini_masks.npz")
    model.test_mask_on, model.semantic_on = True, True
    try:
        out = model([{"image": synth.image(48, 64, seed=0), "height": 96, "width": 128}])
        lo = model.last_outputs
    finally:
        model.test_mask_on, model.semantic_on = False, False
    torch.testing.assert_close(lo["pred_masks"].cpu(), g["pred_masks"], rtol=5e-3, atol=5e-3)
    inst = out[0]["instances"]
    assert torch.equal(inst.pred_classes, g["det0.classes"])
    torch.testing.assert_close(inst.scores, g["det0.scores"], rtol=1e-3, atol=1e-5)
    want = torch.from_numpy(np.unpackbits(g["det0.masks_packed"].numpy(), axis=-1)).bool()
    want = want[..., : int(g["det0.masks_shape"][2])]
    assert inst.pred_masks.dtype == torch.bool and tuple(inst.pred_masks.shape) == tuple(want.shape)
    # thresholded masks: a logit within fp32 noise of 0 may flip a pixel on the boundary
    assert (inst.pred_masks != want).float().mean().item() < 2e-3
    sem = out[0]["sem_seg"].cpu()
    assert sem.shape == g["sem_seg"].shape
    torch.testing.assert_close(sem, g["sem_seg"], rtol=2e-3, atol=2e-3)


def test_masks_engine_fp16_mode_close_to_fp32(mini):
    """fp16 engine mode of the mask / semantic rows (a17 / a19 / a20) against the fp32 mode of the same model (which is pinned
    to the reference golden above): mask logits of the proposals both modes selected, the semantic map and the instance masks."""
    model, _ = mini
    inp = [{"image": synth.image(56, 64, seed=2), "height": 112, "width": 128}]
    model.test_mask_on, model.semantic_on = True, True
    try:
        ref = model(inp)
        ref_logits = model.last_outputs["pred_masks"].float().clone()
        ref_sel = model.transformer.last_topk_proposals[0].tolist()
        model.engine_dtype = torch.float16
        got = model(inp)
        got_logits = model.last_outputs["pred_masks"].float()
        got_sel = model.transformer.last_topk_proposals[0].tolist()
    finally:
        model.engine_dtype = torch.float32
        model.test_mask_on, model.semantic_on = False, False
    assert got_logits.shape == ref_logits.shape and got[0]["sem_seg"].shape == ref[0]["sem_seg"].shape
    common = sorted(set(ref_sel) & set(got_sel))
    assert len(common) >= 0.8 * len(ref_sel), "the two precisions selected mostly different proposals"
    ia = torch.tensor([got_sel.index(i) for i in common], device=got_logits.device)
    ib = torch.tensor([ref_sel.index(i) for i in common], device=got_logits.device)
    a, b = got_logits[0, ia], ref_logits[0, ib]
    rms = b.pow(2).mean().sqrt().item()
    err = (a - b).abs().max().item()
    print(f"mask logits fp16 engine vs fp32: max|err| {err:.3e} on rms {rms:.3e} ({len(common)}/{len(ref_sel)} common proposals)")
    assert err < 3e-2 * max(rms, 1.0)
    sem_err = (got[0]["sem_seg"].float() - ref[0]["sem_seg"].float()).abs().max().item()
    print(f"sem_seg fp16 engine vs fp32: max|err| {sem_err:.3e}")
    assert sem_err < 5e-2
    gi, ri = got[0]["instances"], ref[0]["instances"]
    assert gi.pred_masks.dtype == torch.bool and abs(len(gi) - len(ri)) <= max(2, len(ri) // 5)


def test_mini_matches_oracle_port(mini):
    model, sd = mini
    spec = configs.MINI
    images = [synth.image(56, 64, seed=7), synth.image(64, 33, seed=8)]
    outs = [(112, 128), (64, 33)]
    res, taps = AF.forward(images, outs, synth.text_features(8192, spec["lang_dim"])[: spec["num_classes"]], sd, spec)
    out = model([{"image": im, "height": o[0], "width": o[1]} for im, o in zip(images, outs)])
    valid = (taps["init_reference"] < 1).all(-1)  # see test_mini_matches_reference_golden: ties among padded proposals
    assert torch.equal(model.transformer.last_topk_proposals.cpu()[valid], taps["topk_proposals"][valid])
    torch.testing.assert_close(model.last_outputs["pred_logits"].cpu(), taps["pred_logits"], rtol=1e-3, atol=1e-3)
    torch.testing.assert_close(model.last_outputs["pred_boxes"].cpu(), taps["pred_boxes"], rtol=1e-3, atol=1e-4)
    for r, o in zip(res, out):
        assert torch.equal(o["instances"].pred_classes, r["classes"])
        assert torch.equal(o["instances"].query_index, r["query_index"])
        torch.testing.assert_close(o["instances"].scores, r["scores"], rtol=1e-3, atol=1e-5)


@pytest.mark.parametrize("dtype,tol", [(torch.float16, 2e-2), (torch.bfloat16, 1e-1)])
def test_vit_engine_path_matches_fp32_library_path(mini, dtype, tol):
    """The tensor-core ViT path (tcgen05 GEMMs + LayerNorm/RoPE kernels, window-major token order)
    against the fp32 path of the same module (which is pinned to the reference goldens above)."""
    model, _ = mini
    net = model.backbone.net
    img = torch.randn(2, 3, 64, 64, generator=torch.Generator().manual_seed(5)).to(DEV)  # independent of test order
    want = net(img)["last_feat"]
    got = net(img.to(dtype))["last_feat"]
    assert got.dtype == dtype
    torch.testing.assert_close(got.float(), want, rtol=tol, atol=tol)


def test_pyramid_and_neck_engine_paths_match_fp32_library_path(mini):
    """SimpleFeaturePyramid + ChannelMapper on the token-major engine path (GEMM transposed convs with the pixel
    shuffle folded into LayerNorm row maps, GroupNorm kernel) vs the fp32 library path of the same modules."""
    model, _ = mini
    img = torch.randn(2, 3, 64, 64, generator=torch.Generator().manual_seed(6)).to(DEV)
    want = model.backbone(img)
    want_neck = model.neck(want)
    got = model.backbone(img.half())
    got_neck = model.neck(got)
    for k in want:
        assert got[k].shape == want[k].shape
        torch.testing.assert_close(got[k].float(), want[k], rtol=3e-2, atol=3e-2)
    for a, b in zip(got_neck, want_neck):
        torch.testing.assert_close(a.float(), b, rtol=3e-2, atol=3e-2)


def test_engine_fp16_mode_end_to_end(mini):
    model, sd = mini
    spec = configs.MINI
    images = [synth.image(56, 64, seed=7)]
    res, taps = AF.forward(images, [(112, 128)], synth.text_features(8192, spec["lang_dim"])[: spec["num_classes"]], sd, spec)
    model.engine_dtype = torch.float16
    try:
        out = model([{"image": images[0], "height": 112, "width": 128}])
        logits = model.last_outputs["pred_logits"].float().cpu()
    finally:
        model.engine_dtype = torch.float32
    # fp16 tensor-core mode vs the fp32 oracle: logits within 3e-2 abs on O(5) values (~1e-2 rel)
    torch.testing.assert_close(logits, taps["pred_logits"], rtol=3e-2, atol=3e-2)
    assert len(out[0]["instances"]) == len(res[0]["scores"])


def test_cuda_graph_mode_equals_eager(mini):
    model, _ = mini
    inputs = [{"image": synth.image(64, 64, seed=3), "height": 64, "width": 64}]
    model.engine_dtype = torch.float16
    try:
        eager = model(inputs)
        le = model.last_outputs["pred_logits"].clone()
        model.use_cuda_graphs = True
        for seed in (3, 4, 3):  # first call captures, later calls replay (with a different image in between)
            out = model([{"image": synth.image(64, 64, seed=seed), "height": 64, "width": 64}])
        lg = model.last_outputs["pred_logits"].clone()
    finally:
        model.engine_dtype, model.use_cuda_graphs = torch.float32, False
    torch.testing.assert_close(lg, le, rtol=0, atol=0)
    assert torch.equal(out[0]["instances"].pred_classes, eager[0]["instances"].pred_classes)


def test_no_cpu_fallback(mini):
    model, _ = mini
    with pytest.raises(RuntimeError):
        model.transformer.encoder.layers[0].attentions[0](
            torch.zeros(1, 4, 256), reference_points=torch.zeros(1, 4, 5, 2),
            spatial_shapes=torch.tensor([[2, 2]]), level_start_index=torch.tensor([0]))


@pytest.mark.slow
def test_ape_l_d_1024_matches_oracle_port_stagewise():
    """BASELINE.json config 2 (APE-L_D, 1024^2, 1203 names, boxes only, B=1), fp32: the oracle port runs on
    the host cores (about a minute), the engine on the GPU; same name-derived weights."""
    spec = configs.APE_L_D
    model = _build(spec)
    sd = {k: v for k, v in model.state_dict().items()}
    img = synth.image(1024, 768, seed=0)
    text = synth.text_features(8192, spec["lang_dim"])[:1203]
    torch.set_num_threads(max(1, torch.get_num_threads()))
    res, taps = AF.forward([img], [(1024, 768)], text, sd, spec)
    model = model.to(DEV)
    model.vocabulary = {model.dataset_names[0]: [f"c{i}" for i in range(1203)]}
    out = model([{"image": img, "height": 1024, "width": 768}])
    lo = model.last_outputs
    for k in ("p2", "p4", "p6"):
        torch.testing.assert_close(lo["features"][k].cpu(), taps[f"backbone.{k}"], rtol=2e-3, atol=2e-3)
    # 6 encoder layers after a 24-block ViT, fp32 on both sides with different reduction orders: 241 of 22 M
    # elements exceeded 2e-3 (max 4.4e-3) on B200, so the bound here is 1e-2 absolute on O(1) values
    torch.testing.assert_close(lo["memory"].cpu(), taps["memory"], rtol=1e-2, atol=1e-2)
    sel, want = model.transformer.last_topk_proposals.cpu(), taps["topk_proposals"]
    # Stage-wise parity of the selection itself (bit-exact requirement): the engine's top-k / NMS / quota code on
    # the ORACLE's encoder outputs must return exactly the oracle's indices.
    geo = model._geometry((1, 3, 1024, 1024), [(1024, 768)], None)
    stage = model.transformer.stage_select(taps["enc_outputs_class"].to(DEV), taps["enc_outputs_coord_unact"].to(DEV), geo)
    assert torch.equal(stage.cpu(), want), "selection differs on identical inputs"
    # Selected proposals: identical except where upstream fp32 noise (see above) flips a near-tie of the top-k /
    # NMS ordering; report the agreement and require it to be near-total, then compare the heads on the queries
    # that both sides selected at the same slot.
    a, b = sel[0].tolist(), want[0].tolist()
    common = set(a) & set(b)
    frac = len(common) / len(b)
    print(f"proposal set agreement at full size: {frac:.4f} ({len(common)}/{len(b)})")
    # end to end the two sides see encoder outputs that differ by fp32 reduction-order noise (above); with untrained
    # weights the proposal scores are nearly flat, so the discontinuous top-k / IoU>0.9 NMS flips a few percent
    assert frac > 0.85, f"selected proposal sets differ ({frac:.3f} in common)"
    ia = torch.tensor([a.index(i) for i in sorted(common)])
    ib = torch.tensor([b.index(i) for i in sorted(common)])
    torch.testing.assert_close(lo["pred_logits"].cpu()[0][ia], taps["pred_logits"][0][ib], rtol=1e-2, atol=3e-2)
    torch.testing.assert_close(lo["pred_boxes"].cpu()[0][ia], taps["pred_boxes"][0][ib], rtol=1e-2, atol=5e-3)
    inst = out[0]["instances"]
    # final detections: same number, and the top-scoring ones agree in class and box
    assert abs(len(inst) - len(res[0]["scores"])) <= 3
    k = min(20, len(inst), len(res[0]["scores"]))
    torch.testing.assert_close(inst.scores[:k], res[0]["scores"][:k], rtol=2e-2, atol=1e-3)


@pytest.mark.slow
def test_ape_ti_1024_matches_reference_golden():
    """BASELINE.json configs[0] on the engine: APE-Ti (vit_eva02.py backbone on the fp32 library path, engine kernels
    for deformable attention / NMS), one 768 x 1024 image padded to 1024^2, 80 names, against tensors recorded from the
    reference model on the CPU (tests/golden/model_ti_1024.npz)."""
    spec = configs.APE_TI
    g = load_golden("model_ti_1024.npz")
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    from ape_b200.modeling import build_model

    model = build_model(spec, num_text=80)
    synth.fill_state_dict(model)
    model = model.to(DEV)
    out = model([{"image": synth.image(768, 1024, seed=11), "height": 384, "width": 512}])
    lo = model.last_outputs
    for k in ("p2", "p4", "p6"):
        torch.testing.assert_close(lo["features"][k][:, ::16, ::4, ::4].cpu(), g[f"backbone.{k}"], rtol=2e-3, atol=2e-3)
    # fp32 on both sides, different reduction orders over 6 encoder layers (cf. the APE-L_D full-size test)
    torch.testing.assert_close(lo["memory"][:, ::128].cpu(), g["memory"], rtol=1e-2, atol=1e-2)
    sel, want = model.transformer.last_topk_proposals.cpu()[0].tolist(), g["topk_proposals"][0].tolist()
    frac = len(set(sel) & set(want)) / len(want)
    print(f"APE-Ti proposal set agreement with the reference: {frac:.4f}")
    assert frac > 0.85
    inst = out[0]["instances"]
    assert abs(len(inst) - len(g["det0.scores"])) <= 3
    k = min(20, len(inst))
    torch.testing.assert_close(inst.scores[:k], g["det0.scores"][:k], rtol=2e-2, atol=1e-3)
