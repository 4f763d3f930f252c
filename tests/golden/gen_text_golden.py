"""Generates tests/golden/text_tower_small.npz by running the REFERENCE's own TextTransformer
(/root/reference/ape/modeling/text/eva02_clip/transformer.py:642-737, executed unmodified; its package __init__ is bypassed
because it imports timm / the vision tower) with name-derived synthetic weights on prompts of different lengths.
Build container only:   python tests/golden/gen_text_golden.py"""
import importlib
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import refshim, synth  # noqa: E402

CFG = dict(context_length=77, vocab_size=1000, width=128, heads=2, layers=3, output_dim=64)


def reference_text_transformer():
    refshim.install()
    import timm.models.layers as tl

    if not hasattr(tl, "trunc_normal_"):
        tl.trunc_normal_ = torch.nn.init.trunc_normal_
    for name, path in (("ape.modeling.text", "/root/reference/ape/modeling/text"),
                       ("ape.modeling.text.eva02_clip", "/root/reference/ape/modeling/text/eva02_clip")):
        m = types.ModuleType(name)
        m.__path__ = [path]
        sys.modules[name] = m
    return importlib.import_module("ape.modeling.text.eva02_clip.transformer")


def tokens():
    g = torch.Generator().manual_seed(4)
    lens = [3, 9, 17, 40, 76, 77]
    t = torch.zeros(len(lens), CFG["context_length"], dtype=torch.long)
    for i, n in enumerate(lens):
        t[i, : n - 1] = torch.randint(1, CFG["vocab_size"] - 1, (n - 1,), generator=g)
        t[i, n - 1] = CFG["vocab_size"] - 1  # end-of-text = the highest id (argmax picks it, transformer.py:736)
    return t


def main():
    mod = reference_text_transformer()
    torch.manual_seed(0)
    ref = mod.TextTransformer(**CFG).eval()
    synth.fill_state_dict(ref)
    t = tokens()
    with torch.no_grad():
        eot = ref(t)
        x = ref.token_embedding(t) + ref.positional_embedding
        x = ref.transformer(x.permute(1, 0, 2), attn_mask=ref.attn_mask).permute(1, 0, 2)
        xx = ref.ln_final(x) @ ref.text_projection
    np.savez_compressed(os.path.join(HERE, "text_tower_small.npz"), tokens=t.numpy(), eot=eot.numpy(), all=xx[:, ::7].numpy(),
                        keys=np.frombuffer("\n".join(sorted(ref.state_dict().keys())).encode(), dtype=np.uint8))
    print("text golden", tuple(eot.shape), tuple(xx.shape), float(eot.abs().mean()))


if __name__ == "__main__":
    main()
