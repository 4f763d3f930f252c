"""VisionLanguageFusion: bi-directional vision<->language cross attention of the VL encoder.
Mirror of ape/layers/vision_language_fusion.py:7-53 and ape/layers/fuse_helper.py:8-232 (same
constructor arguments, same parameter names: `b_attn.{layer_norm_v,layer_norm_l,gamma_v,gamma_l}`,
`b_attn.attn.{v_proj,l_proj,values_v_proj,values_l_proj,out_v_proj,out_l_proj}`).

Inference engine: dropout / drop-path / checkpointing arguments are accepted and ignored."""
import torch
import torch.nn as nn


class BiMultiHeadAttention(nn.Module):
    def __init__(self, v_dim, l_dim, embed_dim, num_heads, dropout=0.1, stable_softmax_2d=False,
                 clamp_min_for_underflow=True, clamp_max_for_overflow=True, use_attention_mask_v=False):
        super().__init__()
        self.embed_dim, self.num_heads, self.head_dim = embed_dim, num_heads, embed_dim // num_heads
        self.v_dim, self.l_dim = v_dim, l_dim
        assert self.head_dim * num_heads == embed_dim
        self.scale = self.head_dim ** (-0.5)
        self.v_proj = nn.Linear(v_dim, embed_dim)
        self.l_proj = nn.Linear(l_dim, embed_dim)
        self.values_v_proj = nn.Linear(v_dim, embed_dim)
        self.values_l_proj = nn.Linear(l_dim, embed_dim)
        self.out_v_proj = nn.Linear(embed_dim, v_dim)
        self.out_l_proj = nn.Linear(embed_dim, l_dim)
        self.stable_softmax_2d = stable_softmax_2d
        self.clamp_min_for_underflow = clamp_min_for_underflow
        self.clamp_max_for_overflow = clamp_max_for_overflow
        self.use_attention_mask_v = use_attention_mask_v
        for m in (self.v_proj, self.l_proj, self.values_v_proj, self.values_l_proj, self.out_v_proj, self.out_l_proj):
            nn.init.xavier_uniform_(m.weight)
            m.bias.data.fill_(0)

    def _clamp(self, w):
        if self.clamp_min_for_underflow:
            w = torch.clamp(w, min=-50000)
        if self.clamp_max_for_overflow:
            w = torch.clamp(w, max=50000)
        return w

    def forward(self, v, l, attention_mask_v=None, attention_mask_l=None):
        """fuse_helper.py:67-166 (v, l already layer-normed by the block)."""
        bsz, tgt_len, _ = v.shape
        nh, hd = self.num_heads, self.head_dim

        def heads(t):
            return t.view(bsz, -1, nh, hd).transpose(1, 2).reshape(bsz * nh, -1, hd)

        q = heads(self.v_proj(v) * self.scale)
        k = heads(self.l_proj(l))
        val_v = heads(self.values_v_proj(v))
        val_l = heads(self.values_l_proj(l))
        w = torch.bmm(q, k.transpose(1, 2))
        if self.stable_softmax_2d:
            w = w - w.max()
        w = self._clamp(w)
        wT = w.transpose(1, 2)
        wl = self._clamp(wT - torch.max(wT, dim=-1, keepdim=True)[0])
        if attention_mask_v is not None and self.use_attention_mask_v:
            mv = attention_mask_v[:, None, None, :].repeat(1, nh, 1, 1).flatten(0, 1)
            wl = wl.masked_fill(mv, float("-inf"))
        wl = wl.softmax(dim=-1)
        if attention_mask_l is not None:
            ml = attention_mask_l[:, None, None, :].repeat(1, nh, 1, 1).flatten(0, 1)
            w = w.masked_fill(ml, float("-inf"))
        wv = w.softmax(dim=-1)
        out_v = torch.bmm(wv, val_l).view(bsz, nh, tgt_len, hd).transpose(1, 2).reshape(bsz, tgt_len, self.embed_dim)
        out_l = torch.bmm(wl, val_v).view(bsz, nh, -1, hd).transpose(1, 2).reshape(bsz, -1, self.embed_dim)
        return self.out_v_proj(out_v), self.out_l_proj(out_l)


class BiAttentionBlock(nn.Module):
    def __init__(self, v_dim, l_dim, embed_dim, num_heads, dropout=0.1, drop_path=0.0, init_values=1e-4,
                 stable_softmax_2d=False, clamp_min_for_underflow=True, clamp_max_for_overflow=True,
                 use_attention_mask_v=False):
        super().__init__()
        self.layer_norm_v = nn.LayerNorm(v_dim)
        self.layer_norm_l = nn.LayerNorm(l_dim)
        self.attn = BiMultiHeadAttention(v_dim, l_dim, embed_dim, num_heads, dropout, stable_softmax_2d,
                                         clamp_min_for_underflow, clamp_max_for_overflow, use_attention_mask_v)
        self.gamma_v = nn.Parameter(init_values * torch.ones(v_dim))
        self.gamma_l = nn.Parameter(init_values * torch.ones(l_dim))

    def forward(self, v, l, attention_mask_v=None, attention_mask_l=None):
        # fuse_helper.py:221-232 — note the residual is added to the *normalised* v / l
        v = self.layer_norm_v(v)
        l = self.layer_norm_l(l)
        dv, dl = self.attn(v, l, attention_mask_v=attention_mask_v, attention_mask_l=attention_mask_l)
        return v + self.gamma_v * dv, l + self.gamma_l * dl


class VisionLanguageFusion(nn.Module):
    def __init__(self, v_dim, l_dim, embed_dim, num_heads, dropout=0.1, drop_path=0.0, init_values=1e-4,
                 stable_softmax_2d=False, clamp_min_for_underflow=True, clamp_max_for_overflow=True,
                 use_checkpoint=False, use_attention_mask_v=False):
        super().__init__()
        self.use_checkpoint = use_checkpoint
        self.b_attn = BiAttentionBlock(v_dim, l_dim, embed_dim, num_heads, dropout, drop_path, init_values,
                                       stable_softmax_2d, clamp_min_for_underflow, clamp_max_for_overflow,
                                       use_attention_mask_v)

    def forward(self, v, l, attention_mask_v=None, attention_mask_l=None):
        return self.b_attn(v, l, attention_mask_v, attention_mask_l)
