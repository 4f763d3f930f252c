"""Deterministic synthetic weights / inputs (there are no checkpoints or datasets offline).

Every tensor is a pure function of (parameter name, shape), so the reference-side golden generator
(tests/golden/gen_model_golden.py), the oracle port and the CUDA engine obtain identical weights
by calling `fill_state_dict` on their own state_dict, and bench.py gets "random-init weights of the
real architecture" that are not degenerate (nothing committed, nothing downloaded).  Values are chosen so activations stay O(1) through the network (1/sqrt(fan_in)
matrices, unit norms with 2 % noise) and so that nothing the reference zero-initialises stays
degenerate (multi_scale_deform_attn.py:194,208-209; deformable_detr.py:119-120)."""
import math
import zlib

import torch

_KEEP = ("sampling_offsets.bias",     # grid init (multi_scale_deform_attn.py:195-207): realistic sampling locality
         "log_scale", "name_prompt_fusion_feature", "freqs_cos", "freqs_sin", "pixel_mean", "pixel_std",
         "features_phrase_bank")


def tensor_for(name, shape, dtype=torch.float32):
    g = torch.Generator().manual_seed(zlib.crc32(name.encode()) & 0x7FFFFFFF)
    shape = tuple(shape)
    r = torch.randn(shape, generator=g, dtype=torch.float32)
    leaf = name.rsplit(".", 1)[-1]
    if len(shape) >= 2:
        fan_in = 1
        for s in shape[1:]:
            fan_in *= s
        if "pos_embed" in name or "level_embeds" in name:
            t = r * 0.5
        elif leaf == "in_proj_weight" or "weight" in leaf:
            t = r / math.sqrt(fan_in)
        else:
            t = r * 0.02
    elif leaf in ("gamma_v", "gamma_l"):
        t = 1.0 / 6 + 0.02 * r
    elif leaf == "bias0":
        t = torch.full(shape, -math.log(99.0)) + 0.02 * r
    elif leaf == "weight":  # norm scales
        t = 1.0 + 0.02 * r
    elif "attention_weights.bias" in name:
        t = 0.5 * r
    else:  # biases, bias_lang, q_bias, v_bias
        t = 0.05 * r
    return t.to(dtype)


def fill_state_dict(module):
    """Overwrite every floating-point entry of module.state_dict() in place (except _KEEP)."""
    sd = module.state_dict()
    # shared tensors appear under several names (class_embed / bbox_embed are aliased under
    # transformer.decoder, deformable_detr.py:158-166): the lexicographically smallest name decides
    canon = {}
    for name, p in sd.items():
        key = (p.data_ptr(), tuple(p.shape))
        canon[key] = min(canon.get(key, name), name)
    with torch.no_grad():
        for name, p in sd.items():
            if not p.is_floating_point() or any(k in name for k in _KEEP):
                continue
            cname = canon[(p.data_ptr(), tuple(p.shape))]
            if cname == name:
                p.copy_(tensor_for(name, p.shape, p.dtype))
    return sd


def image(h, w, seed=0):
    """randint(0,256) float32 CHW."""
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, 256, (3, h, w), generator=g).to(torch.float32)


def text_features(n, dim, seed=2):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(n, dim, generator=g)


def suppress_invalid_anchor_logits(module, target=-math.log(99.0)):
    """Shift the biases of the two-stage proposal class heads so that INVALID anchors (padding / border tokens: their
    memory is zeroed before `enc_output`, deformable_transformer_vl.py:354-366, so they all share one constant logit)
    score at the 0.01 prior, as with trained weights.  With purely random heads that constant lands among the top
    scores by chance (1.32 for the APE-L_D names), the per-level top-k fills up with identical degenerate proposals and
    the reference falls into its `nms proposals < topk` branch.  A pure function of the state_dict: the reference-side
    golden generator, the oracle port and the engine all apply it after `fill_state_dict`."""
    sd = module.state_dict()
    w, b = sd["transformer.enc_output.weight"], sd["transformer.enc_output.bias"]
    nw, nb = sd["transformer.enc_output_norm.weight"], sd["transformer.enc_output_norm.bias"]
    const = torch.nn.functional.layer_norm(b.float()[None], (b.numel(),), nw.float(), nb.float(), 1e-5)[0]
    heads = sorted(k[: -len(".weight")] for k in sd if k.endswith(".weight") and sd[k].dim() == 2 and sd[k].shape[0] == 1
                   and k.startswith("transformer.decoder.class_embed"))
    seen = set()
    with torch.no_grad():
        for h in heads:
            hb = sd[h + ".bias"]
            if hb.data_ptr() in seen:
                continue
            seen.add(hb.data_ptr())
            logit = (sd[h + ".weight"].float() @ const.to(sd[h + ".weight"].device) + hb.float())[0]
            hb.add_((target - logit).to(hb.dtype))
    return sd
