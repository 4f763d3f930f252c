"""GPU: the predictor's input pipeline (csrc/preprocess.cu, ape_b200/engine/defaults.py) against PIL — the implementation the
reference's predictor runs (ape/engine/defaults.py:216-222 -> detectron2 ResizeTransform.apply_image -> PIL Image.resize) — and
against the oracle's restatement: bit-exact (uint8 arithmetic)."""
import numpy as np
import pytest
import torch

from oracle import resize as R
from test_preprocess_cpu import CASES, _img

pytestmark = pytest.mark.gpu
PIL_Image = pytest.importorskip("PIL.Image")


@pytest.mark.parametrize("H,W,C,nh,nw", CASES + [(1365, 2048, 3, 683, 1024), (1024, 1024, 3, 1024, 1024), (512, 512, 3, 1024, 1024)])
def test_resize_is_pil_bit_for_bit(built, H, W, C, nh, nw):
    import ape_b200

    img = _img(H, W, C, seed=H + W)
    src = img[:, :, 0] if C == 1 else img
    want = np.asarray(PIL_Image.fromarray(src).resize((nw, nh), PIL_Image.BILINEAR))
    n0 = ape_b200._lib.launch_count()
    got = ape_b200.ops.resize_u8_bilinear(torch.from_numpy(src).cuda(), nh, nw)
    assert ape_b200._lib.launch_count() - n0 == 2
    want_chw = torch.from_numpy(want.astype(np.float32)).view(nh, nw, C).permute(2, 0, 1)
    assert got.dtype == torch.float32 and tuple(got.shape) == (C, nh, nw)
    assert torch.equal(got.cpu(), want_chw), f"max |diff| = {(got.cpu() - want_chw).abs().max().item()}"
    if H * W <= 200 * 301:
        assert np.array_equal(R.resize_u8(src, nh, nw), want)


def test_channel_flip_pitched_rows_and_padded_destination(built):
    import ape_b200

    img = _img(90, 130, 3, seed=5)
    wide = torch.zeros((90, 160, 3), dtype=torch.uint8, device="cuda")
    wide[:, :130] = torch.from_numpy(img).cuda()
    view = wide[:, :130]  # rows pitched at 480 bytes
    batch = torch.full((3, 256, 256), -7.0, device="cuda")
    out = ape_b200.ops.resize_u8_bilinear(view, 177, 256, flip_channels=True, out=batch)
    want = R.predictor_image(img, 177, 256, "RGB")  # BGR -> RGB view, then the resize (defaults.py:218-222)
    assert want.shape == (3, 177, 256)
    assert np.array_equal(out.cpu().numpy(), want)
    assert np.array_equal(batch[:, :177].cpu().numpy(), want) and bool((batch[:, 177:] == -7.0).all())


def test_default_predictor_equals_the_reference_pipeline_on_the_host(built):
    """DefaultPredictor(bgr uint8 image) == model([{image: float CHW tensor built as defaults.py:216-222 builds it (PIL)}])."""
    from ape_b200 import configs
    from ape_b200.engine import DefaultPredictor, ResizeShortestEdge
    from ape_b200.modeling import build_model
    from oracle import synth

    spec = configs.MINI
    model = build_model(spec)
    synth.fill_state_dict(model)
    model = model.cuda().eval()
    g = np.random.default_rng(11)
    bgr = g.integers(0, 256, (45, 60, 3), dtype=np.uint8)
    aug = ResizeShortestEdge(48, 64)
    pred = DefaultPredictor(model, aug, input_format="RGB")

    rgb = bgr[:, :, ::-1]
    nh, nw = R.get_output_shape(45, 60, 48, 64)
    pil = np.asarray(PIL_Image.fromarray(np.ascontiguousarray(rgb)).resize((nw, nh), PIL_Image.BILINEAR))
    host_image = torch.as_tensor(pil.astype("float32").transpose(2, 0, 1))
    inputs = pred.preprocess(bgr)
    assert inputs["height"] == 45 and inputs["width"] == 60 and inputs["image"].is_cuda
    assert torch.equal(inputs["image"].cpu(), host_image)

    want = model([{"image": host_image, "height": 45, "width": 60}])[0]["instances"]
    got = pred(bgr)["instances"]
    assert len(got) == len(want) and len(got) > 0
    assert torch.equal(got.pred_boxes.tensor, want.pred_boxes.tensor)
    assert torch.equal(got.scores, want.scores) and torch.equal(got.pred_classes, want.pred_classes)

    # a stream of images: upload + resize of image i+1 overlap the forward of image i; same results as one by one
    imgs = [g.integers(0, 256, (45, 60, 3), dtype=np.uint8) for _ in range(4)] + [g.integers(0, 256, (60, 45, 3), dtype=np.uint8)]
    one_by_one = [pred(im)["instances"] for im in imgs]
    piped = [o["instances"] for o in pred.predict_batch(imgs)]
    for a, b in zip(one_by_one, piped):
        assert torch.equal(a.pred_boxes.tensor, b.pred_boxes.tensor) and torch.equal(a.scores, b.scores)

    # the shipped mode: 16-bit engine + CUDA graph; predict_batch rides the same graph (selection inside) through forward_packed
    model.engine_dtype, model.use_cuda_graphs = torch.float16, True
    same = [g.integers(0, 256, (45, 60, 3), dtype=np.uint8) for _ in range(5)]
    one_by_one = [pred(im)["instances"] for im in same]
    piped = [o["instances"] for o in pred.predict_batch(same)]
    for a, b in zip(one_by_one, piped):
        assert len(a) == len(b) > 0
        assert torch.equal(a.pred_boxes.tensor, b.pred_boxes.tensor) and torch.equal(a.scores, b.scores)
        assert torch.equal(a.pred_classes, b.pred_classes)
    model.engine_dtype, model.use_cuda_graphs = torch.float32, False

    # a mask prompt rides the same transform (defaults.py:227-229); uint8 masks follow PIL's single-channel path
    mask = (g.random((45, 60)) > 0.6).astype(np.uint8) * 255
    mp = pred.preprocess(bgr, mask_prompt=mask)["mask_prompt"]
    want_mp = np.asarray(PIL_Image.fromarray(mask).resize((nw, nh), PIL_Image.BILINEAR)).astype("float32")
    assert np.array_equal(mp.cpu().numpy(), want_mp)
