"""EVA-02 ViT backbone + SimpleFeaturePyramid of APE-L_D.

Mirror of ape/modeling/backbone/vit_eva_clip.py (`ViT` :570-754, `Block` :383-567, `Attention`
:135-319, `SwiGLU` :101-132, `SimpleFeaturePyramid` :757-922) and utils_eva02.py (`PatchEmbed`
:190-216, `get_abs_pos` :158-187, `VisionRotaryEmbeddingFast` :307-346, window partition :19-63):
same constructor arguments, same parameter / buffer names, so `DetectionCheckpointer.load` fills
them.  Only the configuration APE uses is implemented (sub-LN, naive SwiGLU, 2-D RoPE, q/v bias,
window + global blocks, no rel-pos bias, pre-norm, no layer scale); other switches raise."""
import math
from functools import partial

import torch
import torch.nn as nn
import torch.nn.functional as F

from ..layers.common import ConvNorm, LayerNorm2d


class ShapeSpec:
    def __init__(self, channels=None, height=None, width=None, stride=None):
        self.channels, self.height, self.width, self.stride = channels, height, width, stride


class VisionRotaryEmbeddingFast(nn.Module):
    """utils_eva02.py:307-346: cos/sin tables (ft_seq_len^2, 2*dim) for 2-D rotary embedding."""

    def __init__(self, dim, pt_seq_len=16, ft_seq_len=None, theta=10000):
        super().__init__()
        freqs = 1.0 / (theta ** (torch.arange(0, dim, 2)[: (dim // 2)].float() / dim))
        if ft_seq_len is None:
            ft_seq_len = pt_seq_len
        t = torch.arange(ft_seq_len) / ft_seq_len * pt_seq_len
        freqs = torch.einsum("i,f->if", t, freqs).repeat_interleave(2, dim=-1)
        fh = freqs[:, None, :].expand(ft_seq_len, ft_seq_len, -1)
        fw = freqs[None, :, :].expand(ft_seq_len, ft_seq_len, -1)
        freqs = torch.cat([fh, fw], dim=-1)
        self.register_buffer("freqs_cos", freqs.cos().reshape(-1, freqs.shape[-1]))
        self.register_buffer("freqs_sin", freqs.sin().reshape(-1, freqs.shape[-1]))

    def forward(self, t):
        x = t.reshape(*t.shape[:-1], -1, 2)
        x1, x2 = x.unbind(dim=-1)
        rot = torch.stack((-x2, x1), dim=-1).flatten(-2)
        return t * self.freqs_cos + rot * self.freqs_sin


class PatchEmbed(nn.Module):
    def __init__(self, kernel_size=(16, 16), stride=(16, 16), padding=(0, 0), in_chans=3, embed_dim=768):
        super().__init__()
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=kernel_size, stride=stride, padding=padding)

    def forward(self, x):
        return self.proj(x).permute(0, 2, 3, 1)


def get_abs_pos(abs_pos, has_cls_token, hw):
    h, w = hw
    if has_cls_token:
        abs_pos = abs_pos[:, 1:]
    size = int(math.sqrt(abs_pos.shape[1]))
    assert size * size == abs_pos.shape[1]
    if size != h or size != w:
        new = F.interpolate(abs_pos.reshape(1, size, size, -1).permute(0, 3, 1, 2), size=(h, w), mode="bicubic",
                            align_corners=False)
        return new.permute(0, 2, 3, 1)
    return abs_pos.reshape(1, h, w, -1)


def window_partition(x, ws):
    B, H, W, C = x.shape
    pad_h, pad_w = (ws - H % ws) % ws, (ws - W % ws) % ws
    if pad_h > 0 or pad_w > 0:
        x = F.pad(x, (0, 0, 0, pad_w, 0, pad_h))
    Hp, Wp = H + pad_h, W + pad_w
    x = x.view(B, Hp // ws, ws, Wp // ws, ws, C)
    return x.permute(0, 1, 3, 2, 4, 5).contiguous().view(-1, ws, ws, C), (Hp, Wp)


def window_unpartition(win, ws, pad_hw, hw):
    Hp, Wp = pad_hw
    H, W = hw
    B = win.shape[0] // (Hp * Wp // ws // ws)
    x = win.view(B, Hp // ws, Wp // ws, ws, ws, -1).permute(0, 1, 3, 2, 4, 5).contiguous().view(B, Hp, Wp, -1)
    if Hp > H or Wp > W:
        x = x[:, :H, :W, :].contiguous()
    return x


class SwiGLU(nn.Module):
    def __init__(self, in_features, hidden_features, norm_layer):
        super().__init__()
        self.w1 = nn.Linear(in_features, hidden_features)
        self.w2 = nn.Linear(in_features, hidden_features)
        self.ffn_ln = norm_layer(hidden_features)
        self.w3 = nn.Linear(hidden_features, in_features)

    def forward(self, x):
        return self.w3(self.ffn_ln(F.silu(self.w1(x)) * self.w2(x)))


class Attention(nn.Module):
    def __init__(self, dim, num_heads, rope, norm_layer):
        super().__init__()
        self.num_heads = num_heads
        head_dim = dim // num_heads
        self.scale = head_dim ** -0.5
        self.q_proj = nn.Linear(dim, dim, bias=False)
        self.k_proj = nn.Linear(dim, dim, bias=False)
        self.v_proj = nn.Linear(dim, dim, bias=False)
        self.q_bias = nn.Parameter(torch.zeros(dim))
        self.v_bias = nn.Parameter(torch.zeros(dim))
        self.inner_attn_ln = norm_layer(dim)
        self.proj = nn.Linear(dim, dim)
        self.rope = rope

    def forward(self, x):
        B, H, W, C = x.shape
        N = H * W
        x = x.reshape(B, N, C)
        q = F.linear(x, self.q_proj.weight, self.q_bias)
        k = F.linear(x, self.k_proj.weight, None)
        v = F.linear(x, self.v_proj.weight, self.v_bias)
        q = q.reshape(B, N, self.num_heads, -1).permute(0, 2, 1, 3)
        k = k.reshape(B, N, self.num_heads, -1).permute(0, 2, 1, 3)
        v = v.reshape(B, N, self.num_heads, -1).permute(0, 2, 1, 3)
        q = self.rope(q).type_as(v)
        k = self.rope(k).type_as(v)
        o = F.scaled_dot_product_attention(q, k, v, dropout_p=0.0, scale=self.scale)
        o = o.permute(0, 2, 1, 3).reshape(B, N, -1)
        return self.proj(self.inner_attn_ln(o)).view(B, H, W, C)


class Block(nn.Module):
    def __init__(self, dim, num_heads, mlp_ratio, norm_layer, window_size, rope):
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.attn = Attention(dim, num_heads, rope, norm_layer)
        self.norm2 = norm_layer(dim)
        self.mlp = SwiGLU(dim, int(dim * mlp_ratio), norm_layer)
        self.window_size = window_size

    def forward(self, x):
        shortcut = x
        x = self.norm1(x)
        if self.window_size > 0:
            H, W = x.shape[1], x.shape[2]
            x, pad_hw = window_partition(x, self.window_size)
        x = self.attn(x)
        if self.window_size > 0:
            x = window_unpartition(x, self.window_size, pad_hw, (H, W))
        x = shortcut + x
        return x + self.mlp(self.norm2(x))


class ViT(nn.Module):
    def __init__(self, img_size=1024, patch_size=16, in_chans=3, embed_dim=768, depth=12, num_heads=12,
                 mlp_ratio=4.0, qkv_bias=False, qk_scale=None, drop_rate=0.0, attn_drop_rate=0.0, drop_path_rate=0.0,
                 norm_layer=partial(nn.LayerNorm, eps=1e-6), init_values=None, use_abs_pos=True, use_rel_pos=False,
                 rope=False, postnorm=False, pt_hw_seq_len=16, intp_freq=False, naiveswiglu=False, subln=False,
                 window_size=0, window_block_indexes=(), residual_block_indexes=(), use_act_checkpoint=False,
                 pretrain_img_size=224, pretrain_use_cls_token=True, out_feature="last_feat", xattn=False,
                 frozen_stages=-1):
        super().__init__()
        if not (rope and naiveswiglu and subln and qkv_bias and use_abs_pos and intp_freq) or postnorm or init_values \
                or len(residual_block_indexes) or qk_scale is not None:
            raise NotImplementedError("ape_b200.ViT implements the EVA-02 configuration APE uses "
                                      "(rope, naiveswiglu, subln, qkv_bias, abs pos, intp_freq; pre-norm)")
        self.pretrain_use_cls_token = pretrain_use_cls_token
        self.patch_embed = PatchEmbed((patch_size, patch_size), (patch_size, patch_size), in_chans=in_chans,
                                      embed_dim=embed_dim)
        num_patches = (pretrain_img_size // patch_size) ** 2
        self.pos_embed = nn.Parameter(torch.zeros(1, num_patches + (1 if pretrain_use_cls_token else 0), embed_dim))
        half = embed_dim // num_heads // 2
        self.rope_win = VisionRotaryEmbeddingFast(half, pt_hw_seq_len, window_size)
        self.rope_glb = VisionRotaryEmbeddingFast(half, pt_hw_seq_len, img_size // patch_size)
        self.blocks = nn.ModuleList([
            Block(embed_dim, num_heads, mlp_ratio, norm_layer, window_size if i in window_block_indexes else 0,
                  self.rope_win if i in window_block_indexes else self.rope_glb)
            for i in range(depth)])
        self._out_feature_channels = {out_feature: embed_dim}
        self._out_feature_strides = {out_feature: patch_size}
        self._out_features = [out_feature]
        nn.init.trunc_normal_(self.pos_embed, std=0.02)
        self.apply(self._init_weights)

    @staticmethod
    def _init_weights(m):
        if isinstance(m, nn.Linear):
            nn.init.trunc_normal_(m.weight, std=0.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    def output_shape(self):
        return {n: ShapeSpec(channels=self._out_feature_channels[n], stride=self._out_feature_strides[n])
                for n in self._out_features}

    def forward(self, x):
        x = self.patch_embed(x)
        x = x + get_abs_pos(self.pos_embed, self.pretrain_use_cls_token, (x.shape[1], x.shape[2])).to(x.dtype)
        for blk in self.blocks:
            x = blk(x)
        return {self._out_features[0]: x.permute(0, 3, 1, 2)}


class LastLevelMaxPool(nn.Module):
    def __init__(self):
        super().__init__()
        self.num_levels = 1
        self.in_feature = "p5"

    def forward(self, x):
        return [F.max_pool2d(x, kernel_size=1, stride=2, padding=0)]


class SimpleFeaturePyramid(nn.Module):
    def __init__(self, net, in_feature, out_channels, scale_factors, top_block=None, norm="LN", square_pad=0):
        super().__init__()
        assert norm == "LN"
        self.scale_factors = scale_factors
        shapes = net.output_shape()
        strides = [int(shapes[in_feature].stride / s) for s in scale_factors]
        dim = shapes[in_feature].channels
        self.stages = []
        for idx, scale in enumerate(scale_factors):
            out_dim = dim
            if scale == 4.0:
                layers = [nn.ConvTranspose2d(dim, dim // 2, kernel_size=2, stride=2), LayerNorm2d(dim // 2), nn.GELU(),
                          nn.ConvTranspose2d(dim // 2, dim // 4, kernel_size=2, stride=2)]
                out_dim = dim // 4
            elif scale == 2.0:
                layers = [nn.ConvTranspose2d(dim, dim // 2, kernel_size=2, stride=2)]
                out_dim = dim // 2
            elif scale == 1.0:
                layers = []
            elif scale == 0.5:
                layers = [nn.MaxPool2d(kernel_size=2, stride=2)]
            else:
                raise NotImplementedError(f"scale_factor={scale} is not supported yet.")
            layers.extend([ConvNorm(out_dim, out_channels, 1, bias=False, norm=LayerNorm2d(out_channels)),
                           ConvNorm(out_channels, out_channels, 3, padding=1, bias=False, norm=LayerNorm2d(out_channels))])
            seq = nn.Sequential(*layers)
            stage = int(math.log2(strides[idx]))
            self.add_module(f"simfp_{stage}", seq)
            self.stages.append(seq)
        self.net = net
        self.in_feature = in_feature
        self.top_block = top_block
        self._out_feature_strides = {"p{}".format(int(math.log2(s))): s for s in strides}
        if top_block is not None:
            for s in range(stage, stage + top_block.num_levels):
                self._out_feature_strides["p{}".format(s + 1)] = 2 ** (s + 1)
        self._out_features = list(self._out_feature_strides.keys())
        self._out_feature_channels = {k: out_channels for k in self._out_features}
        self._size_divisibility = strides[-1]
        self._square_pad = square_pad

    @property
    def size_divisibility(self):
        return 0  # detectron2 Backbone default; the reference's SFP does not override it

    @property
    def padding_constraints(self):
        return {"size_divisiblity": self._size_divisibility, "square_size": self._square_pad}

    def output_shape(self):
        return {n: ShapeSpec(channels=self._out_feature_channels[n], stride=self._out_feature_strides[n])
                for n in self._out_features}

    def forward(self, x):
        feats = self.net(x)
        f = feats[self.in_feature]
        results = [stage(f) for stage in self.stages]
        if self.top_block is not None:
            src = feats[self.top_block.in_feature] if self.top_block.in_feature in feats else \
                results[self._out_features.index(self.top_block.in_feature)]
            results.extend(self.top_block(src))
        return dict(zip(self._out_features, results))
