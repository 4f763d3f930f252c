"""oracle/ape_forward.py — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement ("port") of the reference's detection forward pass — APE-L_D style models: EVA-02
ViT + SimpleFeaturePyramid -> ChannelMapper -> deformable VL encoder / two-stage proposals /
decoder -> VisionLanguageAlign + box heads -> thresholding + class-aware NMS.  Plain functional
PyTorch over a state_dict that uses the REFERENCE's parameter names, fp32, written to follow the
reference statement by statement (file:line cited at each function; paths relative to the APE
repository).  Third-party pieces the reference calls (detrex / detectron2) are restated from
knowledge of those libraries at the pinned commits ("parity unpinned" for those).

Pinned by tests/test_oracle_model_golden.py against tests/golden/model_mini_*.npz, which were
produced by the reference's own files executed under oracle/refshim.py
(tests/golden/gen_model_golden.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  It is also what the GPU box uses as CPU baseline (/root/reference does not
travel)."""
import math

import torch
import torch.nn.functional as F
import torchvision

from oracle.msda import msda_torch


def _ln(x, sd, p, eps=1e-5):
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], eps)


def _lin(x, sd, p):
    return F.linear(x, sd[p + ".weight"], sd.get(p + ".bias"))


# ------------------------------------------------------------------------------------------------
# input (deformable_detr_segm_vl.py:846-855, :363-368; detectron2 ImageList.from_tensors)
# ------------------------------------------------------------------------------------------------
def preprocess(images, spec):
    mean = torch.tensor(spec["pixel_mean"]).view(3, 1, 1)
    std = torch.tensor(spec["pixel_std"]).view(3, 1, 1)
    sq = spec["backbone"]["square_pad"]
    sizes = [(int(im.shape[-2]), int(im.shape[-1])) for im in images]
    batch = torch.zeros(len(images), 3, sq, sq)
    masks = torch.ones(len(images), sq, sq)
    for i, im in enumerate(images):
        h, w = sizes[i]
        batch[i, :, :h, :w] = (im.to(torch.float32) - mean) / std  # pad value 0 AFTER normalisation
        masks[i, :h, :w] = 0
    return batch, masks, sizes


# ------------------------------------------------------------------------------------------------
# ViT backbone (vit_eva_clip.py:743-754, blocks :505-523, attention :218-319, SwiGLU :125-132)
# ------------------------------------------------------------------------------------------------
def rope_tables(half_head_dim, pt_seq_len, ft_seq_len, theta=10000.0):
    """VisionRotaryEmbeddingFast.__init__ (utils_eva02.py:307-343)."""
    freqs = 1.0 / (theta ** (torch.arange(0, half_head_dim, 2)[: (half_head_dim // 2)].float() / half_head_dim))
    t = torch.arange(ft_seq_len) / ft_seq_len * pt_seq_len
    freqs = torch.einsum("i,f->if", t, freqs).repeat_interleave(2, dim=-1)  # '... n -> ... (n r)', r=2
    fh = freqs[:, None, :].expand(ft_seq_len, ft_seq_len, -1)
    fw = freqs[None, :, :].expand(ft_seq_len, ft_seq_len, -1)
    freqs = torch.cat([fh, fw], dim=-1)
    return freqs.cos().reshape(-1, freqs.shape[-1]), freqs.sin().reshape(-1, freqs.shape[-1])


def rotate_half(x):
    """utils_eva02.py:248-252."""
    x = x.reshape(*x.shape[:-1], -1, 2)
    x1, x2 = x.unbind(dim=-1)
    return torch.stack((-x2, x1), dim=-1).flatten(-2)


def abs_pos(pos_embed, hw):
    """get_abs_pos (utils_eva02.py:158-187), cls token dropped, bicubic resize."""
    h, w = hw
    ap = pos_embed[:, 1:]
    size = int(math.sqrt(ap.shape[1]))
    assert size * size == ap.shape[1]
    if size != h or size != w:
        new = F.interpolate(ap.reshape(1, size, size, -1).permute(0, 3, 1, 2), size=(h, w), mode="bicubic",
                            align_corners=False)
        return new.permute(0, 2, 3, 1)
    return ap.reshape(1, h, w, -1)


def window_partition(x, ws):
    """utils_eva02.py:19-41."""
    B, H, W, C = x.shape
    pad_h, pad_w = (ws - H % ws) % ws, (ws - W % ws) % ws
    if pad_h > 0 or pad_w > 0:
        x = F.pad(x, (0, 0, 0, pad_w, 0, pad_h))
    Hp, Wp = H + pad_h, W + pad_w
    x = x.view(B, Hp // ws, ws, Wp // ws, ws, C)
    return x.permute(0, 1, 3, 2, 4, 5).contiguous().view(-1, ws, ws, C), (Hp, Wp)


def window_unpartition(win, ws, pad_hw, hw):
    """utils_eva02.py:44-63."""
    Hp, Wp = pad_hw
    H, W = hw
    B = win.shape[0] // (Hp * Wp // ws // ws)
    x = win.view(B, Hp // ws, Wp // ws, ws, ws, -1).permute(0, 1, 3, 2, 4, 5).contiguous().view(B, Hp, Wp, -1)
    if Hp > H or Wp > W:
        x = x[:, :H, :W, :].contiguous()
    return x


def vit_attention(x, sd, p, num_heads, rope):
    """Attention.forward, subln branch + SDPA (vit_eva_clip.py:218-268, :316)."""
    B, H, W, C = x.shape
    N = H * W
    x = x.reshape(B, N, C)
    fused = p + ".qkv.weight" in sd  # vit_eva02.py Attention (APE-Ti): one qkv projection, no inner LayerNorm
    if fused:
        bias = torch.cat((sd[p + ".q_bias"], torch.zeros_like(sd[p + ".v_bias"]), sd[p + ".v_bias"]))
        qkv = F.linear(x, sd[p + ".qkv.weight"], bias).reshape(B, N, 3, num_heads, -1).permute(2, 0, 3, 1, 4)
        q, k, v = qkv[0], qkv[1], qkv[2]
    else:
        q = F.linear(x, sd[p + ".q_proj.weight"], sd[p + ".q_bias"])
        k = F.linear(x, sd[p + ".k_proj.weight"], None)
        v = F.linear(x, sd[p + ".v_proj.weight"], sd[p + ".v_bias"])
        q = q.reshape(B, N, num_heads, -1).permute(0, 2, 1, 3)
        k = k.reshape(B, N, num_heads, -1).permute(0, 2, 1, 3)
        v = v.reshape(B, N, num_heads, -1).permute(0, 2, 1, 3)
    cos, sin = rope
    q = q * cos + rotate_half(q) * sin
    k = k * cos + rotate_half(k) * sin
    scale = q.shape[-1] ** -0.5
    o = F.scaled_dot_product_attention(q, k, v, dropout_p=0.0, scale=scale)
    o = o.permute(0, 2, 1, 3).reshape(B, N, -1)
    if not fused:
        o = _ln(o, sd, p + ".inner_attn_ln", 1e-6)
    o = _lin(o, sd, p + ".proj")
    return o.view(B, H, W, C)


def vit_block(x, sd, p, num_heads, window_size, rope):
    """Block.forward, gamma_1 is None and not postnorm (vit_eva_clip.py:505-523)."""
    shortcut = x
    x = _ln(x, sd, p + ".norm1", 1e-6)
    if window_size > 0:
        H, W = x.shape[1], x.shape[2]
        x, pad_hw = window_partition(x, window_size)
    x = vit_attention(x, sd, p + ".attn", num_heads, rope)
    if window_size > 0:
        x = window_unpartition(x, window_size, pad_hw, (H, W))
    x = shortcut + x
    y = _ln(x, sd, p + ".norm2", 1e-6)
    if p + ".mlp.w12.weight" in sd:  # packed SwiGLU without inner norm (vit_eva02.py xops_SwiGLU, APE-Ti)
        w12, b12 = sd[p + ".mlp.w12.weight"], sd[p + ".mlp.w12.bias"]
        hid = w12.shape[0] // 2
        hidden = F.silu(F.linear(y, w12[:hid], b12[:hid])) * F.linear(y, w12[hid:], b12[hid:])
        return x + _lin(hidden, sd, p + ".mlp.w3")
    # SwiGLU (vit_eva_clip.py:125-132)
    hidden = F.silu(_lin(y, sd, p + ".mlp.w1")) * _lin(y, sd, p + ".mlp.w2")
    hidden = _ln(hidden, sd, p + ".mlp.ffn_ln", 1e-6)
    return x + _lin(hidden, sd, p + ".mlp.w3")


def vit_forward(img, sd, spec, prefix="backbone.net"):
    b = spec["backbone"]
    ps = b["patch_size"]
    x = F.conv2d(img, sd[prefix + ".patch_embed.proj.weight"], sd[prefix + ".patch_embed.proj.bias"], stride=ps)
    x = x.permute(0, 2, 3, 1)
    x = x + abs_pos(sd[prefix + ".pos_embed"], (x.shape[1], x.shape[2]))
    half = b["embed_dim"] // b["num_heads"] // 2
    rope_win = rope_tables(half, b["pt_hw_seq_len"], b["window_size"])
    rope_glb = rope_tables(half, b["pt_hw_seq_len"], b["img_size"] // ps)
    for i in range(b["depth"]):
        win = i in b["window_block_indexes"]
        x = vit_block(x, sd, f"{prefix}.blocks.{i}", b["num_heads"], b["window_size"] if win else 0,
                      rope_win if win else rope_glb)
    return x.permute(0, 3, 1, 2)


# ------------------------------------------------------------------------------------------------
# SimpleFeaturePyramid (vit_eva_clip.py:804-847, :871-922) + LastLevelMaxPool
# ------------------------------------------------------------------------------------------------
def ln2d(x, w, b, eps=1e-6):
    """detectron2 LayerNorm (channels-first)."""
    u = x.mean(1, keepdim=True)
    s = (x - u).pow(2).mean(1, keepdim=True)
    x = (x - u) / torch.sqrt(s + eps)
    return w[:, None, None] * x + b[:, None, None]


def sfp_forward(feat, sd, spec, prefix="backbone"):
    outs = {}
    for stage, scale in zip((2, 3, 4, 5), spec["backbone"]["scale_factors"]):
        p = f"{prefix}.simfp_{stage}"
        x = feat
        idx = 0
        if scale == 4.0:
            x = F.conv_transpose2d(x, sd[f"{p}.0.weight"], sd[f"{p}.0.bias"], stride=2)
            x = ln2d(x, sd[f"{p}.1.weight"], sd[f"{p}.1.bias"])
            x = F.gelu(x)
            x = F.conv_transpose2d(x, sd[f"{p}.3.weight"], sd[f"{p}.3.bias"], stride=2)
            idx = 4
        elif scale == 2.0:
            x = F.conv_transpose2d(x, sd[f"{p}.0.weight"], sd[f"{p}.0.bias"], stride=2)
            idx = 1
        elif scale == 0.5:
            x = F.max_pool2d(x, kernel_size=2, stride=2)
            idx = 1
        x = F.conv2d(x, sd[f"{p}.{idx}.weight"])
        x = ln2d(x, sd[f"{p}.{idx}.norm.weight"], sd[f"{p}.{idx}.norm.bias"])
        x = F.conv2d(x, sd[f"{p}.{idx + 1}.weight"], padding=1)
        x = ln2d(x, sd[f"{p}.{idx + 1}.norm.weight"], sd[f"{p}.{idx + 1}.norm.bias"])
        outs[f"p{stage}"] = x
    outs["p6"] = F.max_pool2d(outs["p5"], kernel_size=1, stride=2, padding=0)
    return outs


# ------------------------------------------------------------------------------------------------
# neck, masks, sine position embedding (detrex ChannelMapper / PositionEmbeddingSine;
# deformable_detr_segm_vl.py:375-392)
# ------------------------------------------------------------------------------------------------
def neck_forward(feats, sd, spec, prefix="neck"):
    outs = []
    for i, f in enumerate(("p2", "p3", "p4", "p5", "p6")):
        x = F.conv2d(feats[f], sd[f"{prefix}.convs.{i}.conv.weight"], sd[f"{prefix}.convs.{i}.conv.bias"])
        x = F.group_norm(x, spec["gn_groups"], sd[f"{prefix}.convs.{i}.norm.weight"], sd[f"{prefix}.convs.{i}.norm.bias"])
        outs.append(x)
    return outs


def sine_pos_embed(mask, num_pos_feats, temperature=10000, offset=-0.5, eps=1e-6, scale=2 * math.pi):
    not_mask = ~mask
    y = not_mask.cumsum(1, dtype=torch.float32)
    x = not_mask.cumsum(2, dtype=torch.float32)
    y = (y + offset) / (y[:, -1:, :] + eps) * scale
    x = (x + offset) / (x[:, :, -1:] + eps) * scale
    dim_t = torch.arange(num_pos_feats, dtype=torch.float32)
    dim_t = temperature ** (2 * torch.div(dim_t, 2, rounding_mode="floor") / num_pos_feats)
    px = x[:, :, :, None] / dim_t
    py = y[:, :, :, None] / dim_t
    B, H, W = mask.shape
    px = torch.stack((px[:, :, :, 0::2].sin(), px[:, :, :, 1::2].cos()), dim=4).view(B, H, W, -1)
    py = torch.stack((py[:, :, :, 0::2].sin(), py[:, :, :, 1::2].cos()), dim=4).view(B, H, W, -1)
    return torch.cat((py, px), dim=3).permute(0, 3, 1, 2)


# ------------------------------------------------------------------------------------------------
# layers: MSDA module, VLF, FFN, MHA
# ------------------------------------------------------------------------------------------------
def msda_module(query, value, identity, query_pos, key_padding_mask, reference_points, spatial_shapes, sd, p, spec):
    """MultiScaleDeformableAttention.forward, batch_first (multi_scale_deform_attn.py:260-358)."""
    Hh, L, P = spec["num_heads"], spec["num_levels"], spec["num_points"]
    if value is None:
        value = query
    if identity is None:
        identity = query
    if query_pos is not None:
        query = query + query_pos
    bs, nq, _ = query.shape
    nv = value.shape[1]
    value = _lin(value, sd, p + ".value_proj")
    if key_padding_mask is not None:
        value = value.masked_fill(key_padding_mask[..., None], 0.0)
    value = value.view(bs, nv, Hh, -1)
    off = _lin(query, sd, p + ".sampling_offsets").view(bs, nq, Hh, L, P, 2)
    aw = _lin(query, sd, p + ".attention_weights").view(bs, nq, Hh, L * P).softmax(-1).view(bs, nq, Hh, L, P)
    if reference_points.shape[-1] == 2:
        norm = torch.stack([spatial_shapes[..., 1], spatial_shapes[..., 0]], -1)
        loc = reference_points[:, :, None, :, None, :] + off / norm[None, None, None, :, None, :]
    else:
        loc = reference_points[:, :, None, :, None, :2] + off / P * reference_points[:, :, None, :, None, 2:] * 0.5
    out = msda_torch(value, spatial_shapes, loc, aw)
    return _lin(out, sd, p + ".output_proj") + identity


def vlf_layer(v, l, sd, p, spec):
    """BiAttentionBlock.forward + BiMultiHeadAttention.forward (fuse_helper.py:221-232, :67-166),
    stable_softmax_2d, both clamps, no attention masks (eval, 'name'/'phrase' prompts)."""
    nh = spec["vlf_heads"]
    E = spec["vlf_embed"]
    hd = E // nh
    v = _ln(v, sd, p + ".layer_norm_v")
    l = _ln(l, sd, p + ".layer_norm_l")
    bsz, tgt, _ = v.shape
    a = p + ".attn"

    def shape(t):
        return t.view(bsz, -1, nh, hd).transpose(1, 2).contiguous().view(bsz * nh, -1, hd)

    q = shape(_lin(v, sd, a + ".v_proj") * hd ** -0.5)
    k = shape(_lin(l, sd, a + ".l_proj"))
    vv = shape(_lin(v, sd, a + ".values_v_proj"))
    vl = shape(_lin(l, sd, a + ".values_l_proj"))
    w = torch.bmm(q, k.transpose(1, 2))
    w = w - w.max()
    w = torch.clamp(torch.clamp(w, min=-50000), max=50000)
    wT = w.transpose(1, 2)
    wl = wT - torch.max(wT, dim=-1, keepdim=True)[0]
    wl = torch.clamp(torch.clamp(wl, min=-50000), max=50000).softmax(dim=-1)
    wv = w.softmax(dim=-1)
    ov = torch.bmm(wv, vl).view(bsz, nh, tgt, hd).transpose(1, 2).reshape(bsz, tgt, E)
    ol = torch.bmm(wl, vv).view(bsz, nh, -1, hd).transpose(1, 2).reshape(bsz, -1, E)
    dv = _lin(ov, sd, a + ".out_v_proj")
    dl = _lin(ol, sd, a + ".out_l_proj")
    return v + sd[p + ".gamma_v"] * dv, l + sd[p + ".gamma_l"] * dl


def ffn(x, sd, p):
    """detrex FFN(num_fcs=2, ReLU, add_identity)."""
    return x + _lin(F.relu(_lin(x, sd, p + ".layers.0.0")), sd, p + ".layers.1")


def mha_self(x, pos, sd, p, nh):
    """detrex MultiheadAttention around nn.MultiheadAttention: q = k = x + pos, v = x, + identity."""
    E = x.shape[-1]
    w, b = sd[p + ".attn.in_proj_weight"], sd[p + ".attn.in_proj_bias"]
    qk = x + pos
    q = F.linear(qk, w[:E], b[:E])
    k = F.linear(qk, w[E:2 * E], b[E:2 * E])
    v = F.linear(x, w[2 * E:], b[2 * E:])
    B, N, _ = x.shape
    q, k, v = (t.view(B, N, nh, E // nh).transpose(1, 2) for t in (q, k, v))
    o = F.scaled_dot_product_attention(q, k, v)
    o = o.transpose(1, 2).reshape(B, N, E)
    return x + F.linear(o, sd[p + ".attn.out_proj.weight"], sd[p + ".attn.out_proj.bias"])


def mlp(x, sd, p, n):
    """detrex MLP: n Linear layers, ReLU between."""
    for i in range(n):
        x = _lin(x, sd, f"{p}.layers.{i}")
        if i < n - 1:
            x = F.relu(x)
    return x


def inverse_sigmoid(x, eps=1e-3):
    x = x.clamp(min=0, max=1)
    return torch.log(x.clamp(min=eps) / (1 - x).clamp(min=eps))


def box_cxcywh_to_xyxy(b):
    cx, cy, w, h = b.unbind(-1)
    return torch.stack([cx - 0.5 * w, cy - 0.5 * h, cx + 0.5 * w, cy + 0.5 * h], dim=-1)


# ------------------------------------------------------------------------------------------------
# transformer (deformable_transformer_vl.py:422-699)
# ------------------------------------------------------------------------------------------------
def valid_ratio(mask):
    """get_valid_ratio (:402-410)."""
    _, H, W = mask.shape
    vh = torch.sum(~mask[:, :, 0], 1).float() / H
    vw = torch.sum(~mask[:, 0, :], 1).float() / W
    return torch.stack([vw, vh], -1)


def encoder_reference_points(shapes, valid_ratios):
    """get_reference_points (:371-400)."""
    pts = []
    for lvl, (H, W) in enumerate(shapes):
        ry, rx = torch.meshgrid(torch.linspace(0.5, H - 0.5, H), torch.linspace(0.5, W - 0.5, W), indexing="ij")
        ry = ry.reshape(-1)[None] / (valid_ratios[:, None, lvl, 1] * H)
        rx = rx.reshape(-1)[None] / (valid_ratios[:, None, lvl, 0] * W)
        pts.append(torch.stack((rx, ry), -1))
    ref = torch.cat(pts, 1)
    return ref[:, :, None] * valid_ratios[:, None]


def encoder_output_proposals(memory, mask_flat, shapes, sd, p="transformer"):
    """gen_encoder_output_proposals (:321-369), no mask prompt."""
    N = memory.shape[0]
    props, level_ids = [], []
    cur = 0
    for lvl, (H, W) in enumerate(shapes):
        m = mask_flat[:, cur:cur + H * W].view(N, H, W, 1)
        vH = torch.sum(~m[:, :, 0, 0], 1)
        vW = torch.sum(~m[:, 0, :, 0], 1)
        gy, gx = torch.meshgrid(torch.linspace(0, H - 1, H), torch.linspace(0, W - 1, W), indexing="ij")
        grid = torch.cat([gx.unsqueeze(-1), gy.unsqueeze(-1)], -1)
        scale = torch.cat([vW.unsqueeze(-1), vH.unsqueeze(-1)], 1).view(N, 1, 1, 2)
        grid = (grid.unsqueeze(0).expand(N, -1, -1, -1) + 0.5) / scale
        wh = torch.ones_like(grid) * 0.05 * (2.0 ** lvl)
        props.append(torch.cat((grid, wh), -1).view(N, -1, 4))
        cur += H * W
        level_ids.append(torch.full((H * W,), lvl, dtype=torch.long))
    out = torch.cat(props, 1)
    valid = ((out > 0.01) & (out < 0.99)).all(-1, keepdim=True)
    out = torch.log(out / (1 - out))
    out = out.masked_fill(mask_flat.unsqueeze(-1), float("inf"))
    out = out.masked_fill(~valid, float("inf"))
    mem = memory.masked_fill(mask_flat.unsqueeze(-1), 0.0).masked_fill(~valid, 0.0)
    mem = _ln(_lin(mem, sd, p + ".enc_output"), sd, p + ".enc_output_norm")
    return mem, out, torch.cat(level_ids)


def select_proposals(logit, boxes_unact, level_ids, n_levels, topk, pre_nms_topk, nms_thresh):
    """The assign_first_stage branch (:569-625) for one image -> LongTensor[topk]."""
    boxes = box_cxcywh_to_xyxy(boxes_unact.sigmoid()).clamp(0, 1)
    pre = []
    for lvl in range(n_levels):
        lvl_mask = level_ids == lvl
        # reference: torch.topk(...)[1].  For levels with fewer tokens than pre_nms_topk the tail of that result is a
        # tie among zero scores whose order torch leaves implementation-defined (CPU != CUDA); the oracle and the
        # engine both use the stable order (lowest index first), one of the valid outcomes of the reference call.
        order = torch.sort(logit.sigmoid() * lvl_mask, descending=True, stable=True)[1]
        pre.append(order[: min(pre_nms_topk, logit.size(0))])
    pre = torch.cat(pre)
    post = torchvision.ops.boxes.batched_nms(boxes[pre], logit[pre], level_ids[pre], nms_thresh)
    keep = pre[post]
    if len(keep) < topk:
        keep = torch.sort(logit, descending=True, stable=True)[1][: min(topk, logit.size(0))]
    q_per_l = topk // n_levels
    ordered = level_ids[keep][None] == torch.arange(n_levels)[:, None]
    km = (ordered & (ordered.cumsum(1) <= q_per_l)).any(0)
    if km.sum() < topk:
        num_to_add = topk - km.sum()
        pad = (~km).nonzero()[:num_to_add]
        km[pad] = True
    return keep[km]


def proposal_pos_embed(proposals, num_pos_feats=128, temperature=10000):
    """get_proposal_pos_embed (:412-420)."""
    dim_t = torch.arange(num_pos_feats, dtype=torch.float32)
    dim_t = temperature ** (2 * torch.div(dim_t, 2, rounding_mode="floor") / num_pos_feats)
    proposals = proposals.sigmoid() * (2 * math.pi)
    pos = proposals[:, :, :, None] / dim_t
    return torch.stack((pos[:, :, :, 0::2].sin(), pos[:, :, :, 1::2].cos()), dim=4).flatten(2)


def transformer_forward(feats, masks, pos_embeds, query_l, sd, spec, p="transformer", taps=None):
    L = spec["num_levels"]
    shapes = [(int(f.shape[2]), int(f.shape[3])) for f in feats]
    feat = torch.cat([f.flatten(2).transpose(1, 2) for f in feats], 1)
    mask_flat = torch.cat([m.flatten(1) for m in masks], 1)
    lvl_pos = torch.cat([pe.flatten(2).transpose(1, 2) + sd[p + ".level_embeds"][i].view(1, 1, -1)
                         for i, pe in enumerate(pos_embeds)], 1)
    ss = torch.as_tensor(shapes, dtype=torch.long)
    vr = torch.stack([valid_ratio(m) for m in masks], 1)
    ref = encoder_reference_points(shapes, vr)

    # encoder (:84-115): [VLF] -> MSDA self-attn -> LN -> FFN -> LN
    x = feat
    for i in range(spec["enc_layers"]):
        if query_l is not None:
            x, query_l = vlf_layer(x, query_l, sd, f"{p}.encoder.vl_layers.{i}.b_attn", spec)
            if taps is not None:
                taps[f"vlf{i}.v"], taps[f"vlf{i}.l"] = x, query_l
        lp = f"{p}.encoder.layers.{i}"
        x = msda_module(x, None, None, lvl_pos, mask_flat, ref, ss, sd, lp + ".attentions.0", spec)
        x = _ln(x, sd, lp + ".norms.0")
        x = ffn(x, sd, lp + ".ffns.0")
        x = _ln(x, sd, lp + ".norms.1")
    memory = x
    bs, _, c = memory.shape

    # two-stage (:496-645)
    out_mem, out_props, level_ids = encoder_output_proposals(memory, mask_flat, shapes, sd, p)
    nd = spec["dec_layers"]
    enc_cls = _lin(out_mem, sd, f"{p}.decoder.class_embed.{nd}")
    enc_coord = mlp(out_mem, sd, f"{p}.decoder.bbox_embed.{nd}", 3) + out_props
    if spec["proposal_ambiguous"]:
        cls_all = torch.stack([enc_cls] + [_lin(out_mem, sd, f"{p}.decoder.class_embed_ambiguous.{j}")
                                           for j in range(spec["proposal_ambiguous"])], dim=1)
        coord_all = torch.stack([enc_coord] + [mlp(out_mem, sd, f"{p}.decoder.bbox_embed_ambiguous.{j}", 3) + out_props
                                               for j in range(spec["proposal_ambiguous"])], dim=1)
        idx = torch.argmax(cls_all, dim=1, keepdim=True)
        enc_cls = torch.gather(cls_all, 1, idx).squeeze(1)
        enc_coord = torch.gather(coord_all, 1, idx.repeat(1, 1, 1, 4)).squeeze(1)
    topk = spec["num_queries"]
    logit = enc_cls[..., 0]
    sel = torch.stack([select_proposals(logit[b], enc_coord[b], level_ids, L, topk, spec["pre_nms_topk"],
                                        spec["nms_thresh_enc"]) for b in range(bs)])
    topk_unact = torch.gather(enc_coord, 1, sel.unsqueeze(-1).repeat(1, 1, 4))
    reference = topk_unact.sigmoid()
    init_reference = reference
    pt = _ln(_lin(proposal_pos_embed(topk_unact), sd, p + ".pos_trans"), sd, p + ".pos_trans_norm")
    query_pos, query = torch.split(pt, c, dim=2)
    topk_feats = torch.stack([out_mem[b][sel[b]] for b in range(bs)])
    query = query + _ln(_lin(topk_feats, sd, p + ".pix_trans"), sd, p + ".pix_trans_norm")

    # decoder (:195-250)
    inter, inter_ref = [], []
    out = query
    for i in range(nd):
        ref_in = reference[:, :, None] * torch.cat([vr, vr], -1)[:, None]
        lp = f"{p}.decoder.layers.{i}"
        out = mha_self(out, query_pos, sd, lp + ".attentions.0", spec["num_heads"])
        out = _ln(out, sd, lp + ".norms.0")
        out = msda_module(out, memory, None, query_pos, mask_flat, ref_in, ss, sd, lp + ".attentions.1", spec)
        out = _ln(out, sd, lp + ".norms.1")
        out = ffn(out, sd, lp + ".ffns.0")
        out = _ln(out, sd, lp + ".norms.2")
        tmp = mlp(out, sd, f"{p}.decoder.bbox_embed.{i}", 3)
        reference = (tmp + inverse_sigmoid(reference)).sigmoid()
        inter.append(out)
        inter_ref.append(reference)
    return dict(inter_states=torch.stack(inter), init_reference=init_reference, inter_references=torch.stack(inter_ref),
                enc_outputs_class=enc_cls, enc_outputs_coord_unact=enc_coord, memory=memory, query_l=query_l,
                topk_proposals=sel, spatial_shapes=ss)


# ------------------------------------------------------------------------------------------------
# heads + inference (deformable_detr_segm_vl.py:482-503, :759-810; vision_language_align.py:27-52;
# fast_rcnn.py:97-201)
# ------------------------------------------------------------------------------------------------
def vl_align(x, emb, sd, p):
    emb = F.normalize(emb.to(x.dtype), p=2, dim=-1)
    tok = _lin(emb / 2.0, sd, p + ".dot_product_projection_text")
    bias = torch.matmul(emb, sd[p + ".bias_lang"]) + sd[p + ".bias0"]
    logit = torch.matmul(x, tok.transpose(-1, -2)) / sd[p + ".log_scale"].exp() + bias.unsqueeze(1)
    return torch.clamp(torch.clamp(logit, max=50000), min=-50000)


def fast_rcnn_inference_single(boxes, scores, image_shape, score_thresh, nms_thresh, topk):
    valid = torch.isfinite(boxes).all(dim=1) & torch.isfinite(scores).all(dim=1)
    if not valid.all():
        boxes, scores = boxes[valid], scores[valid]
    scores = scores[:, :-1]
    h, w = image_shape
    boxes = torch.stack((boxes[:, 0].clamp(0, w), boxes[:, 1].clamp(0, h), boxes[:, 2].clamp(0, w),
                         boxes[:, 3].clamp(0, h)), dim=-1)
    filt = scores > score_thresh
    inds = filt.nonzero()
    boxes = boxes[inds[:, 0]]
    scores = scores[filt]
    keep = torchvision.ops.boxes.batched_nms(boxes.float(), scores, inds[:, 1], nms_thresh)
    if topk >= 0:
        keep = keep[:topk]
    return boxes[keep], scores[keep], inds[keep, 1], inds[keep, 0]


def detector_postprocess(boxes, image_size, out_h, out_w):
    """detectron2 detector_postprocess for boxes: rescale, clip, drop empty."""
    sx, sy = out_w / image_size[1], out_h / image_size[0]
    boxes = boxes.clone()
    boxes[:, 0::2] *= sx
    boxes[:, 1::2] *= sy
    boxes = torch.stack((boxes[:, 0].clamp(0, out_w), boxes[:, 1].clamp(0, out_h), boxes[:, 2].clamp(0, out_w),
                         boxes[:, 3].clamp(0, out_h)), dim=-1)
    keep = ((boxes[:, 2] - boxes[:, 0]) > 0) & ((boxes[:, 3] - boxes[:, 1]) > 0)
    return boxes, keep


def build_text_inputs(text_feats, spec, phrase, batch, bank_reset=False):
    """deformable_detr_segm_vl.py:236-360: returns (features_l for the classifier, features_l_fusion).
    name prompt: classifier sees the N_t name features, fusion sees the zero 'name_prompt_fusion_feature';
    phrase prompt: both see [features ; zero bank rows] truncated to max(N_t, num_classes) (:320-326)."""
    if not phrase:
        fl = text_feats.unsqueeze(0).repeat(batch, 1, 1)
        fusion = torch.zeros(batch, 1, spec["lang_dim"])
        return fl, fusion
    fl = text_feats.to(torch.float32)
    if bank_reset:  # text_feature_bank_reset=True (:320-326); at eval_dataset_id=-1 without reset no bank rows are added
        bank = torch.zeros(spec["num_classes"], spec["lang_dim"])
        fl = torch.cat([fl, bank * 0], dim=0)[: max(text_feats.shape[0], spec["num_classes"])]
    fl = fl.unsqueeze(0).repeat(batch, 1, 1)
    return fl, fl


# ------------------------------------------------------------------------------------------------
# mask head, instance-mask post-processing, semantic branch (deformable_detr_segm_vl.py:728-750, :507-517,
# :563-603, :628-666, :875-918; detectron2 BitMasks.crop_and_resize / paste_masks_in_image / sem_seg_postprocess)
# ------------------------------------------------------------------------------------------------
def mask_features(memory, p2, shapes, sd):
    """maskdino_mask_features (:728-750), mask_encode_level 0, GroupNorm(32) convs without bias."""
    h, w = shapes[0]
    enc = memory[:, : h * w].permute(0, 2, 1).reshape(memory.shape[0], -1, h, w)
    x = F.group_norm(F.conv2d(p2, sd["lateral_conv.weight"]), 32, sd["lateral_conv.norm.weight"], sd["lateral_conv.norm.bias"])
    x = x + F.interpolate(enc, size=x.shape[-2:], mode="bilinear", align_corners=False)
    x = F.relu(F.group_norm(F.conv2d(x, sd["output_conv.weight"], padding=1), 32, sd["output_conv.norm.weight"],
                            sd["output_conv.norm.bias"]))
    return F.conv2d(x, sd["mask_conv.weight"])


def crop_and_resize(masks, boxes, size):
    """detectron2 BitMasks.crop_and_resize: ROIAlign(size, scale 1, sampling_ratio 0, aligned) on the float mask, >= 0.5."""
    rois = torch.cat([torch.arange(len(boxes), dtype=boxes.dtype)[:, None], boxes], dim=1)
    return torchvision.ops.roi_align(masks.to(torch.float32)[:, None], rois, (size, size), 1.0, 0, True).squeeze(1) >= 0.5


def paste_masks(masks, boxes, out_h, out_w, threshold=0.5):
    """detectron2 paste_masks_in_image (bilinear grid_sample, align_corners=False), whole image per mask."""
    out = torch.zeros(len(masks), out_h, out_w, dtype=torch.bool)
    for i in range(len(masks)):
        x0, y0, x1, y1 = boxes[i]
        iy = (torch.arange(out_h, dtype=torch.float32) + 0.5 - y0) / (y1 - y0) * 2 - 1
        ix = (torch.arange(out_w, dtype=torch.float32) + 0.5 - x0) / (x1 - x0) * 2 - 1
        grid = torch.stack([ix[None, :].expand(out_h, out_w), iy[:, None].expand(out_h, out_w)], dim=2)[None]
        out[i] = F.grid_sample(masks[i][None, None].float(), grid, align_corners=False)[0, 0] >= threshold
    return out


def forward(images, out_sizes, text_feats, sd, spec, phrase=False, taps=None, bank_reset=False, masks_on=False,
            semantic_on=False, topk=None):
    """Whole detection forward.  images: list of CHW float RGB 0..255; out_sizes: list of (height, width) the
    detections are rescaled to; text_feats [N_t, lang_dim].  masks_on / semantic_on: the reference's test_mask_on /
    semantic_on branches ("thing" entity).  topk: detections kept per image (default spec["test_topk"]; 1 for
    "expression" prompts, deformable_detr_segm_vl.py:184-193, which otherwise follow the phrase path).
    Returns list of dicts(boxes, scores, classes, query_index[, masks, sem_seg]) + the tap dict."""
    taps = {} if taps is None else taps
    with torch.no_grad():
        batch, img_masks, sizes = preprocess(images, spec)
        feat = vit_forward(batch, sd, spec)
        pyr = sfp_forward(feat, sd, spec)
        taps.update({f"backbone.{k}": v for k, v in pyr.items()})
        ml = neck_forward(pyr, sd, spec)
        taps.update({f"neck.{i}": v for i, v in enumerate(ml)})
        masks = [F.interpolate(img_masks[None], size=f.shape[-2:]).to(torch.bool).squeeze(0) for f in ml]
        pos = [sine_pos_embed(m, spec["embed_dim"] // 2) for m in masks]
        fl, fusion = build_text_inputs(text_feats, spec, phrase, len(images), bank_reset)
        tr = transformer_forward(ml, masks, pos, fusion, sd, spec, taps=taps)
        taps.update({k: tr[k] for k in ("memory", "inter_states", "init_reference", "inter_references",
                                        "enc_outputs_class", "enc_outputs_coord_unact", "topk_proposals")})
        if phrase:
            fl = 0.0 * fl + 1.0 * tr["query_l"]  # :448
        else:
            fl = 1.0 * fl + 0.0 * tr["query_l"]  # :446
        nd = spec["dec_layers"]
        hs = tr["inter_states"][nd - 1]
        reference = tr["init_reference"] if nd == 1 else tr["inter_references"][nd - 2]
        logits = vl_align(hs, fl, sd, f"class_embed.{nd - 1}")
        coord = (mlp(hs, sd, f"bbox_embed.{nd - 1}", 3) + inverse_sigmoid(reference)).sigmoid()
        taps["pred_logits"], taps["pred_boxes"] = logits, coord
        mask_pred = None
        if masks_on or semantic_on:
            shapes = [tuple(f.shape[-2:]) for f in ml]
            mf = mask_features(tr["memory"], pyr["p2"], shapes, sd)
            mask_pred = torch.einsum("bqc,bchw->bqhw", mlp(hs, sd, "mask_embed", 3), mf)  # last level only (:517)
            taps["mask_features"], taps["pred_masks"] = mf, mask_pred
        padded = tuple(batch.shape[-2:])
        results = []
        for b in range(len(images)):
            scores = torch.cat((logits[b].sigmoid(), torch.zeros(logits.shape[1], 1)), dim=1)
            h, w = sizes[b]
            boxes = box_cxcywh_to_xyxy(coord[b]) * torch.tensor([w, h, w, h], dtype=torch.float32)
            bx0, sc, cl, qi = fast_rcnn_inference_single(boxes, scores, (h, w), spec["test_score_thresh"],
                                                         spec["test_nms_thresh"], spec["test_topk"] if topk is None else topk)
            bx, keep = detector_postprocess(bx0, (h, w), out_sizes[b][0], out_sizes[b][1])
            r = dict(boxes=bx[keep], scores=sc[keep], classes=cl[keep], query_index=qi[keep])
            if masks_on:  # (:588-603) + detector_postprocess
                m = F.interpolate(mask_pred[b, qi][None], size=padded, mode="bilinear", align_corners=False)[0]
                m128 = crop_and_resize(m.sigmoid() > 0.5, bx0, 128).to(torch.float32)
                r["masks"] = paste_masks(m128[keep], bx[keep], out_sizes[b][0], out_sizes[b][1])
            if semantic_on:  # (:628-666, :875-918), semantic_post_nms with the detector's own thresholds
                cls = F.softmax(logits[b, qi].sigmoid() / 0.06, dim=-1)
                m = F.interpolate(mask_pred[b, qi][None], size=padded, mode="bilinear", align_corners=False)[0].sigmoid()
                sem = torch.einsum("qc,qhw->chw", cls, m)[:, :h, :w].expand(1, -1, -1, -1)
                r["sem_seg"] = F.interpolate(sem, size=tuple(out_sizes[b]), mode="bilinear", align_corners=False)[0]
            results.append(r)
    return results, taps
