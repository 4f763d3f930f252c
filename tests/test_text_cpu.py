"""CPU: the text tower mirror (ape_b200/modeling/text.py) against tensors recorded from the reference's own TextTransformer
(tests/golden/gen_text_golden.py), parameter names, and the `forward_text` dict contract (clip_wrapper_eva02.py:94-158)."""
import numpy as np
import torch

from conftest import load_golden
from oracle import synth


def _small():
    from ape_b200.modeling.text import TextTransformer

    m = TextTransformer(context_length=77, vocab_size=1000, width=128, heads=2, layers=3, output_dim=64).eval()
    synth.fill_state_dict(m)
    return m


def test_literal_path_equals_reference_golden(built):
    g = load_golden("text_tower_small.npz")
    m = _small()
    with torch.no_grad():
        eot, xx = m.encode(g["tokens"])
    torch.testing.assert_close(eot, g["eot"], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(xx[:, ::7], g["all"], rtol=1e-5, atol=1e-6)


def test_parameter_names_equal_the_reference(built):
    import os

    from conftest import GOLDEN

    want = set(bytes(np.load(os.path.join(GOLDEN, "text_tower_small.npz"))["keys"]).decode().split("\n"))
    assert set(_small().state_dict().keys()) == want
    from ape_b200.modeling.text import EVA02CLIP

    with torch.device("meta"):
        big = EVA02CLIP("EVA02-CLIP-bigE-14-plus")
    sd = big.state_dict()
    assert sd["net.text.transformer.resblocks.31.attn.in_proj_weight"].shape == (3 * 1280, 1280)
    assert sd["net.text.text_projection"].shape == (1280, 1024) and sd["net.text.token_embedding.weight"].shape == (49408, 1280)
    assert "net.logit_scale" in sd and sum(v.numel() for v in sd.values()) > 6.5e8


def test_forward_text_contract(built):
    from ape_b200.modeling.text import EVA02CLIP

    clip = EVA02CLIP(text_cfg=dict(context_length=77, vocab_size=1000, width=128, heads=2, layers=2), embed_dim=64, dtype="float32",
                     tokenizer=lambda texts: torch.stack([torch.cat([torch.arange(1, 1 + len(t.split())), torch.tensor([999]),
                                                                     torch.zeros(76 - len(t.split()), dtype=torch.long)]) for t in texts]))
    out = clip.forward_text(["a dog", "the red apple on the left"], cache=True)
    assert set(out) == {"end_token_idx", "attention_mask", "last_hidden_state", "last_hidden_state_eot"}
    assert out["last_hidden_state_eot"].shape == (2, 64) and out["last_hidden_state"].shape == (2, 77, 64)
    assert out["end_token_idx"].tolist() == [2, 6] and out["attention_mask"].sum(1).tolist() == [3, 7]
    assert clip.forward_text(["a dog", "the red apple on the left"], cache=True) is out  # cached (clip_wrapper_eva02.py:82-85)
