"""oracle/refshim.py — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Stand-ins for the third-party packages the reference imports on the detection forward path but
which are absent from this image (detectron2 @017abbf, detrex @776058e, fvcore, timm, fairscale),
so that the reference's own .py files can be executed UNMODIFIED from /root/reference to generate
golden vectors (tests/golden/gen_*.py).  Only the build container has /root/reference; nothing
here is importable by the product (ape_b200/) and nothing here runs on the GPU box except as
the restated semantics that oracle/ape_forward.py also follows.

The semantics below restate detectron2 / detrex behaviour from knowledge of those libraries at
the commits the reference pins (requirements.txt:10-12).  They are not covered by any reference
test ("parity unpinned" for these third-party pieces).

Usage:
    from oracle import refshim
    refshim.install()
    vl = refshim.load("ape.modeling.ape_deta.deformable_transformer_vl")
"""
import copy
import importlib
import math
import os
import sys
import types
import warnings
from dataclasses import dataclass
from typing import Any, Dict, List, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F
import torchvision

REF = os.environ.get("APE_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "ape"))


# --------------------------------------------------------------------------------------------
# detectron2.layers
# --------------------------------------------------------------------------------------------
@dataclass
class ShapeSpec:
    channels: Optional[int] = None
    height: Optional[int] = None
    width: Optional[int] = None
    stride: Optional[int] = None


class Conv2d(nn.Conv2d):
    """detectron2.layers.Conv2d: conv -> norm -> activation."""

    def __init__(self, *args, **kwargs):
        norm = kwargs.pop("norm", None)
        activation = kwargs.pop("activation", None)
        super().__init__(*args, **kwargs)
        self.norm = norm
        self.activation = activation

    def forward(self, x):
        x = F.conv2d(x, self.weight, self.bias, self.stride, self.padding, self.dilation, self.groups)
        if self.norm is not None:
            x = self.norm(x)
        if self.activation is not None:
            x = self.activation(x)
        return x


class LayerNorm2d(nn.Module):
    """detectron2.layers.batch_norm.LayerNorm: channels-first LN over C, eps 1e-6."""

    def __init__(self, normalized_shape, eps=1e-6):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(normalized_shape))
        self.bias = nn.Parameter(torch.zeros(normalized_shape))
        self.eps = eps
        self.normalized_shape = (normalized_shape,)

    def forward(self, x):
        u = x.mean(1, keepdim=True)
        s = (x - u).pow(2).mean(1, keepdim=True)
        x = (x - u) / torch.sqrt(s + self.eps)
        return self.weight[:, None, None] * x + self.bias[:, None, None]


def get_norm(norm, out_channels):
    if norm is None:
        return None
    if isinstance(norm, str):
        if len(norm) == 0:
            return None
        norm = {"GN": lambda c: nn.GroupNorm(32, c), "LN": lambda c: LayerNorm2d(c)}[norm]
    return norm(out_channels)


class CNNBlockBase(nn.Module):
    def __init__(self, in_channels, out_channels, stride):
        super().__init__()
        self.in_channels, self.out_channels, self.stride = in_channels, out_channels, stride


def move_device_like(src, dst):
    return src.to(dst.device)


def batched_nms(boxes, scores, idxs, iou_threshold):
    assert boxes.shape[-1] == 4
    return torchvision.ops.boxes.batched_nms(boxes.float(), scores, idxs, iou_threshold)


# --------------------------------------------------------------------------------------------
# detectron2.structures
# --------------------------------------------------------------------------------------------
class Boxes:
    def __init__(self, tensor):
        if not isinstance(tensor, torch.Tensor):
            tensor = torch.as_tensor(tensor, dtype=torch.float32)
        else:
            tensor = tensor.to(torch.float32)
        if tensor.numel() == 0:
            tensor = tensor.reshape((-1, 4)).to(dtype=torch.float32)
        self.tensor = tensor

    def to(self, device):
        return Boxes(self.tensor.to(device=device))

    def clip(self, box_size):
        h, w = box_size
        x1 = self.tensor[:, 0].clamp(min=0, max=w)
        y1 = self.tensor[:, 1].clamp(min=0, max=h)
        x2 = self.tensor[:, 2].clamp(min=0, max=w)
        y2 = self.tensor[:, 3].clamp(min=0, max=h)
        self.tensor = torch.stack((x1, y1, x2, y2), dim=-1)

    def nonempty(self, threshold=0.0):
        box = self.tensor
        return ((box[:, 2] - box[:, 0]) > threshold) & ((box[:, 3] - box[:, 1]) > threshold)

    def scale(self, sx, sy):
        self.tensor[:, 0::2] *= sx
        self.tensor[:, 1::2] *= sy

    def __getitem__(self, item):
        if isinstance(item, int):
            return Boxes(self.tensor[item].view(1, -1))
        return Boxes(self.tensor[item])

    def __len__(self):
        return self.tensor.shape[0]

    @property
    def device(self):
        return self.tensor.device


class Instances:
    def __init__(self, image_size, **kwargs):
        self._image_size = image_size
        self._fields: Dict[str, Any] = {}
        for k, v in kwargs.items():
            self.set(k, v)

    @property
    def image_size(self):
        return self._image_size

    def __setattr__(self, name, val):
        if name.startswith("_"):
            super().__setattr__(name, val)
        else:
            self.set(name, val)

    def __getattr__(self, name):
        if name == "_fields" or name not in self._fields:
            raise AttributeError(name)
        return self._fields[name]

    def set(self, name, value):
        self._fields[name] = value

    def has(self, name):
        return name in self._fields

    def get(self, name):
        return self._fields[name]

    def get_fields(self):
        return self._fields

    def to(self, *args, **kwargs):
        ret = Instances(self._image_size)
        for k, v in self._fields.items():
            if hasattr(v, "to"):
                v = v.to(*args, **kwargs)
            ret.set(k, v)
        return ret

    def __getitem__(self, item):
        ret = Instances(self._image_size)
        for k, v in self._fields.items():
            ret.set(k, v[item])
        return ret

    def __len__(self):
        for v in self._fields.values():
            return v.__len__()
        raise NotImplementedError("Empty Instances does not support __len__!")


class ImageList:
    def __init__(self, tensor, image_sizes):
        self.tensor = tensor
        self.image_sizes = image_sizes

    def __len__(self):
        return len(self.image_sizes)

    @staticmethod
    def from_tensors(tensors, size_divisibility=0, pad_value=0.0, padding_constraints=None):
        image_sizes = [(im.shape[-2], im.shape[-1]) for im in tensors]
        max_size = torch.tensor(image_sizes).max(0).values
        if padding_constraints is not None:
            square_size = padding_constraints.get("square_size", 0)
            if square_size > 0:
                max_size[0] = max_size[1] = square_size
            if "size_divisibility" in padding_constraints:  # (the reference misspells the key)
                size_divisibility = padding_constraints["size_divisibility"]
        if size_divisibility > 1:
            stride = size_divisibility
            max_size = (max_size + (stride - 1)).div(stride, rounding_mode="floor") * stride
        if len(tensors) == 1:
            h, w = image_sizes[0]
            padded = F.pad(tensors[0], [0, int(max_size[-1]) - w, 0, int(max_size[-2]) - h], value=pad_value).unsqueeze_(0)
        else:
            shape = [len(tensors)] + list(tensors[0].shape[:-2]) + [int(max_size[0]), int(max_size[1])]
            padded = tensors[0].new_full(shape, pad_value)
            for i, img in enumerate(tensors):
                padded[i, ..., : img.shape[-2], : img.shape[-1]].copy_(img)
        return ImageList(padded.contiguous(), image_sizes)


class BitMasks:
    def __init__(self, tensor):
        self.tensor = tensor.to(torch.bool)

    def crop_and_resize(self, boxes, mask_size):
        from torchvision.ops import roi_align

        device = self.tensor.device
        batch_inds = torch.arange(len(boxes), device=device).to(dtype=boxes.dtype)[:, None]
        rois = torch.cat([batch_inds, boxes], dim=1)
        bit_masks = self.tensor.to(dtype=torch.float32)
        rois = rois.to(device=device)
        output = roi_align(bit_masks[:, None, :, :], rois, (mask_size, mask_size), 1.0, 0, True).squeeze(1)
        return output >= 0.5


def detector_postprocess(results, output_height, output_width, mask_threshold=0.5):
    new_size = (output_height, output_width)
    scale_x, scale_y = (output_width / results.image_size[1], output_height / results.image_size[0])
    results = Instances(new_size, **results.get_fields())
    output_boxes = results.pred_boxes
    output_boxes.scale(scale_x, scale_y)
    output_boxes.clip(results.image_size)
    results = results[output_boxes.nonempty()]
    if results.has("pred_masks"):
        # ROIMasks(results.pred_masks[:, 0]).to_bitmasks(boxes, H, W, mask_threshold) -> paste_masks_in_image
        results.pred_masks = paste_masks_in_image(results.pred_masks[:, 0, :, :], results.pred_boxes.tensor,
                                                  (output_height, output_width), threshold=mask_threshold)
    return results


def _do_paste_mask(masks, boxes, img_h, img_w, skip_empty=True):
    """detectron2/layers/mask_ops.py::_do_paste_mask (restated)."""
    device = masks.device
    if skip_empty:
        x0_int, y0_int = torch.clamp(boxes.min(dim=0).values.floor()[:2] - 1, min=0).to(dtype=torch.int32)
        x1_int = torch.clamp(boxes[:, 2].max().ceil() + 1, max=img_w).to(dtype=torch.int32)
        y1_int = torch.clamp(boxes[:, 3].max().ceil() + 1, max=img_h).to(dtype=torch.int32)
    else:
        x0_int, y0_int = 0, 0
        x1_int, y1_int = img_w, img_h
    x0, y0, x1, y1 = torch.split(boxes, 1, dim=1)
    N = masks.shape[0]
    img_y = torch.arange(int(y0_int), int(y1_int), device=device, dtype=torch.float32) + 0.5
    img_x = torch.arange(int(x0_int), int(x1_int), device=device, dtype=torch.float32) + 0.5
    img_y = (img_y - y0) / (y1 - y0) * 2 - 1
    img_x = (img_x - x0) / (x1 - x0) * 2 - 1
    gx = img_x[:, None, :].expand(N, img_y.size(1), img_x.size(1))
    gy = img_y[:, :, None].expand(N, img_y.size(1), img_x.size(1))
    grid = torch.stack([gx, gy], dim=3)
    if not masks.dtype.is_floating_point:
        masks = masks.float()
    img_masks = F.grid_sample(masks, grid.to(masks.dtype), align_corners=False)
    if skip_empty:
        return img_masks[:, 0], (slice(int(y0_int), int(y1_int)), slice(int(x0_int), int(x1_int)))
    return img_masks[:, 0], ()


def paste_masks_in_image(masks, boxes, image_shape, threshold=0.5):
    """detectron2/layers/mask_ops.py::paste_masks_in_image (restated): CPU = one mask per chunk with skip_empty."""
    N = len(masks)
    img_h, img_w = int(image_shape[0]), int(image_shape[1])
    if N == 0:
        return masks.new_empty((0, img_h, img_w), dtype=torch.uint8)
    device = boxes.device
    num_chunks = N if device.type == "cpu" else max(1, int(math.ceil(N * img_h * img_w * 4 / 1024 ** 3)))
    chunks = torch.chunk(torch.arange(N, device=device), num_chunks)
    img_masks = torch.zeros(N, img_h, img_w, device=device, dtype=torch.bool if threshold >= 0 else torch.uint8)
    for inds in chunks:
        masks_chunk, spatial_inds = _do_paste_mask(masks[inds, None, :, :], boxes[inds], img_h, img_w,
                                                   skip_empty=device.type == "cpu")
        if threshold >= 0:
            masks_chunk = (masks_chunk >= threshold).to(dtype=torch.bool)
        else:
            masks_chunk = (masks_chunk * 255).to(dtype=torch.uint8)
        img_masks[(inds,) + spatial_inds] = masks_chunk
    return img_masks


def sem_seg_postprocess(result, img_size, output_height, output_width):
    result = result[:, : img_size[0], : img_size[1]].expand(1, -1, -1, -1)
    return F.interpolate(result, size=(output_height, output_width), mode="bilinear", align_corners=False)[0]


# --------------------------------------------------------------------------------------------
# detectron2.modeling.backbone
# --------------------------------------------------------------------------------------------
class Backbone(nn.Module):
    @property
    def size_divisibility(self):
        return 0

    @property
    def padding_constraints(self):
        return {}

    def output_shape(self):
        return {name: ShapeSpec(channels=self._out_feature_channels[name], stride=self._out_feature_strides[name])
                for name in self._out_features}


def _assert_strides_are_log2_contiguous(strides):
    for i, stride in enumerate(strides[1:], 1):
        assert stride == 2 * strides[i - 1], "Strides {} {} are not log2 contiguous".format(stride, strides[i - 1])


class LastLevelMaxPool(nn.Module):
    def __init__(self):
        super().__init__()
        self.num_levels = 1
        self.in_feature = "p5"

    def forward(self, x):
        return [F.max_pool2d(x, kernel_size=1, stride=2, padding=0)]


# --------------------------------------------------------------------------------------------
# detrex.layers / detrex.utils / detrex.modeling.neck
# --------------------------------------------------------------------------------------------
class MLP(nn.Module):
    def __init__(self, input_dim, hidden_dim, output_dim, num_layers):
        super().__init__()
        self.num_layers = num_layers
        h = [hidden_dim] * (num_layers - 1)
        self.layers = nn.ModuleList(nn.Linear(n, k) for n, k in zip([input_dim] + h, h + [output_dim]))

    def forward(self, x):
        for i, layer in enumerate(self.layers):
            x = F.relu(layer(x)) if i < self.num_layers - 1 else layer(x)
        return x


class FFN(nn.Module):
    def __init__(self, embed_dim=256, feedforward_dim=1024, output_dim=None, num_fcs=2,
                 activation=nn.ReLU(inplace=True), ffn_drop=0.0, fc_bias=True, add_identity=True):
        super().__init__()
        assert num_fcs >= 2
        self.embed_dim, self.feedforward_dim, self.num_fcs, self.activation = embed_dim, feedforward_dim, num_fcs, activation
        output_dim = embed_dim if output_dim is None else output_dim
        layers = []
        in_channels = embed_dim
        for _ in range(num_fcs - 1):
            layers.append(nn.Sequential(nn.Linear(in_channels, feedforward_dim, bias=fc_bias), self.activation, nn.Dropout(ffn_drop)))
            in_channels = feedforward_dim
        layers.append(nn.Linear(feedforward_dim, output_dim, bias=fc_bias))
        layers.append(nn.Dropout(ffn_drop))
        self.layers = nn.Sequential(*layers)
        self.add_identity = add_identity

    def forward(self, x, identity=None):
        out = self.layers(x)
        if not self.add_identity:
            return out
        if identity is None:
            identity = x
        return identity + out


class MultiheadAttention(nn.Module):
    def __init__(self, embed_dim, num_heads, attn_drop=0.0, proj_drop=0.0, batch_first=False, **kwargs):
        super().__init__()
        self.embed_dim, self.num_heads, self.batch_first = embed_dim, num_heads, batch_first
        self.attn = nn.MultiheadAttention(embed_dim=embed_dim, num_heads=num_heads, dropout=attn_drop,
                                          batch_first=batch_first, **kwargs)
        self.proj_drop = nn.Dropout(proj_drop)

    def forward(self, query, key=None, value=None, identity=None, query_pos=None, key_pos=None,
                attn_mask=None, key_padding_mask=None, **kwargs):
        if key is None:
            key = query
        if value is None:
            value = key
        if identity is None:
            identity = query
        if key_pos is None:
            if query_pos is not None:
                if query_pos.shape == key.shape:
                    key_pos = query_pos
                else:
                    warnings.warn("position encoding of key is missing in MultiheadAttention.")
        if query_pos is not None:
            query = query + query_pos
        if key_pos is not None:
            key = key + key_pos
        out = self.attn(query=query, key=key, value=value, attn_mask=attn_mask, key_padding_mask=key_padding_mask)[0]
        return identity + self.proj_drop(out)


class BaseTransformerLayer(nn.Module):
    def __init__(self, attn, ffn, norm, operation_order=None):
        super().__init__()
        assert set(operation_order).issubset({"self_attn", "norm", "cross_attn", "ffn"})
        num_attn = operation_order.count("self_attn") + operation_order.count("cross_attn")
        if isinstance(attn, nn.Module):
            attn = [copy.deepcopy(attn) for _ in range(num_attn)]
        else:
            assert len(attn) == num_attn
        self.num_attn = num_attn
        self.operation_order = operation_order
        self.pre_norm = operation_order[0] == "norm"
        self.attentions = nn.ModuleList()
        index = 0
        for op in operation_order:
            if op in ["self_attn", "cross_attn"]:
                self.attentions.append(attn[index])
                index += 1
        self.embed_dim = self.attentions[0].embed_dim
        self.ffns = nn.ModuleList()
        for _ in range(operation_order.count("ffn")):
            self.ffns.append(copy.deepcopy(ffn))
        self.norms = nn.ModuleList()
        for _ in range(operation_order.count("norm")):
            self.norms.append(copy.deepcopy(norm))

    def forward(self, query, key=None, value=None, query_pos=None, key_pos=None, attn_masks=None,
                query_key_padding_mask=None, key_padding_mask=None, **kwargs):
        norm_index = attn_index = ffn_index = 0
        identity = query
        if attn_masks is None:
            attn_masks = [None for _ in range(self.num_attn)]
        elif isinstance(attn_masks, torch.Tensor):
            attn_masks = [copy.deepcopy(attn_masks) for _ in range(self.num_attn)]
        for layer in self.operation_order:
            if layer == "self_attn":
                temp_key = temp_value = query
                query = self.attentions[attn_index](
                    query, temp_key, temp_value, identity if self.pre_norm else None, query_pos=query_pos,
                    key_pos=query_pos, attn_mask=attn_masks[attn_index], key_padding_mask=query_key_padding_mask,
                    **kwargs)
                attn_index += 1
                identity = query
            elif layer == "norm":
                query = self.norms[norm_index](query)
                norm_index += 1
            elif layer == "cross_attn":
                query = self.attentions[attn_index](
                    query, key, value, identity if self.pre_norm else None, query_pos=query_pos, key_pos=key_pos,
                    attn_mask=attn_masks[attn_index], key_padding_mask=key_padding_mask, **kwargs)
                attn_index += 1
                identity = query
            elif layer == "ffn":
                query = self.ffns[ffn_index](query, identity if self.pre_norm else None)
                ffn_index += 1
        return query


class TransformerLayerSequence(nn.Module):
    def __init__(self, transformer_layers=None, num_layers=None):
        super().__init__()
        self.num_layers = num_layers
        self.layers = nn.ModuleList()
        if isinstance(transformer_layers, nn.Module):
            for _ in range(num_layers):
                self.layers.append(copy.deepcopy(transformer_layers))
        else:
            assert isinstance(transformer_layers, list) and len(transformer_layers) == num_layers

    def forward(self):
        raise NotImplementedError()


class PositionEmbeddingSine(nn.Module):
    def __init__(self, num_pos_feats=64, temperature=10000, scale=2 * math.pi, eps=1e-6, offset=0.0, normalize=False):
        super().__init__()
        if scale is not None and normalize is False:
            raise ValueError("normalize should be True if scale is passed")
        self.num_pos_feats, self.temperature, self.normalize = num_pos_feats, temperature, normalize
        self.scale, self.eps, self.offset = scale, eps, offset

    def forward(self, mask):
        assert mask is not None
        not_mask = ~mask
        y_embed = not_mask.cumsum(1, dtype=torch.float32)
        x_embed = not_mask.cumsum(2, dtype=torch.float32)
        if self.normalize:
            y_embed = (y_embed + self.offset) / (y_embed[:, -1:, :] + self.eps) * self.scale
            x_embed = (x_embed + self.offset) / (x_embed[:, :, -1:] + self.eps) * self.scale
        dim_t = torch.arange(self.num_pos_feats, dtype=torch.float32, device=mask.device)
        dim_t = self.temperature ** (2 * torch.div(dim_t, 2, rounding_mode="floor") / self.num_pos_feats)
        pos_x = x_embed[:, :, :, None] / dim_t
        pos_y = y_embed[:, :, :, None] / dim_t
        B, H, W = mask.size()
        pos_x = torch.stack((pos_x[:, :, :, 0::2].sin(), pos_x[:, :, :, 1::2].cos()), dim=4).view(B, H, W, -1)
        pos_y = torch.stack((pos_y[:, :, :, 0::2].sin(), pos_y[:, :, :, 1::2].cos()), dim=4).view(B, H, W, -1)
        return torch.cat((pos_y, pos_x), dim=3).permute(0, 3, 1, 2)


class ConvNormAct(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size=1, stride=1, padding=0, dilation=1, groups=1,
                 bias=True, norm_layer=None, activation=None, **kwargs):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size=kernel_size, stride=stride, padding=padding,
                              dilation=dilation, groups=groups, bias=bias, **kwargs)
        self.norm = norm_layer
        self.activation = activation

    def forward(self, x):
        x = self.conv(x)
        if self.norm is not None:
            x = self.norm(x)
        if self.activation is not None:
            x = self.activation(x)
        return x


class ChannelMapper(nn.Module):
    def __init__(self, input_shapes, in_features, out_channels, kernel_size=1, stride=1, bias=True, groups=1,
                 dilation=1, norm_layer=None, activation=None, num_outs=None, **kwargs):
        super().__init__()
        self.extra_convs = None
        in_channels_per_feature = [input_shapes[f].channels for f in in_features]
        if num_outs is None:
            num_outs = len(input_shapes)
        self.convs = nn.ModuleList()
        for in_channel in in_channels_per_feature:
            self.convs.append(ConvNormAct(in_channel, out_channels, kernel_size=kernel_size, stride=stride,
                                          padding=(kernel_size - 1) // 2, bias=bias, groups=groups, dilation=dilation,
                                          norm_layer=copy.deepcopy(norm_layer), activation=copy.deepcopy(activation)))
        if num_outs > len(in_channels_per_feature):
            self.extra_convs = nn.ModuleList()
            for i in range(len(in_channels_per_feature), num_outs):
                in_channel = in_channels_per_feature[-1] if i == len(in_channels_per_feature) else out_channels
                self.extra_convs.append(ConvNormAct(in_channel, out_channels, kernel_size=3, stride=2, padding=1,
                                                    bias=bias, groups=groups, dilation=dilation,
                                                    norm_layer=copy.deepcopy(norm_layer),
                                                    activation=copy.deepcopy(activation)))
        self.input_shapes, self.in_features, self.out_channels = input_shapes, in_features, out_channels

    def forward(self, inputs):
        assert len(inputs) == len(self.convs)
        outs = [self.convs[i](inputs[self.in_features[i]]) for i in range(len(inputs))]
        if self.extra_convs:
            for i in range(len(self.extra_convs)):
                outs.append(self.extra_convs[i](inputs[self.in_features[-1]] if i == 0 else outs[-1]))
        return tuple(outs)


def inverse_sigmoid(x, eps=1e-3):
    x = x.clamp(min=0, max=1)
    x1 = x.clamp(min=eps)
    x2 = (1 - x).clamp(min=eps)
    return torch.log(x1 / x2)


def box_cxcywh_to_xyxy(bbox):
    cx, cy, w, h = bbox.unbind(-1)
    return torch.stack([(cx - 0.5 * w), (cy - 0.5 * h), (cx + 0.5 * w), (cy + 0.5 * h)], dim=-1)


def box_xyxy_to_cxcywh(bbox):
    x0, y0, x1, y1 = bbox.unbind(-1)
    return torch.stack([(x0 + x1) / 2, (y0 + y1) / 2, (x1 - x0), (y1 - y0)], dim=-1)


# --------------------------------------------------------------------------------------------
# fvcore / timm / misc
# --------------------------------------------------------------------------------------------
def c2_xavier_fill(module):
    nn.init.kaiming_uniform_(module.weight, a=1)
    if module.bias is not None:
        nn.init.constant_(module.bias, 0)


def c2_msra_fill(module):
    nn.init.kaiming_normal_(module.weight, mode="fan_out", nonlinearity="relu")
    if module.bias is not None:
        nn.init.constant_(module.bias, 0)


class DropPath(nn.Module):
    def __init__(self, drop_prob=0.0):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        assert not self.training or self.drop_prob == 0.0, "refshim DropPath: eval only"
        return x


class _Metadata(types.SimpleNamespace):
    def get(self, key, default=None):
        return getattr(self, key, default)


class _MetadataCatalog:
    def __init__(self):
        self._d = {}

    def get(self, name):
        if name not in self._d:
            self._d[name] = _Metadata(name=name)
        return self._d[name]


MetadataCatalog = _MetadataCatalog()


def retry_if_cuda_oom(func):
    return func


def _unavailable(name):
    def f(*a, **k):
        raise NotImplementedError(f"refshim: {name} is not restated")

    return f


# --------------------------------------------------------------------------------------------
# installation
# --------------------------------------------------------------------------------------------
def _mod(name, **attrs):
    m = sys.modules.get(name)
    if m is None:
        m = types.ModuleType(name)
        sys.modules[name] = m
        if "." in name:
            parent, child = name.rsplit(".", 1)
            setattr(_mod(parent), child, m)
    for k, v in attrs.items():
        setattr(m, k, v)
    return m


def _pkg(name, path):
    m = _mod(name)
    m.__path__ = [path]
    return m


_installed = False


def install():
    """Register the stand-in modules and the `ape` package skeleton (package __init__ files of the
    reference are NOT executed: they import training-only code that needs more of detectron2)."""
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError(f"{REF} not present: the reference can only be executed in the build container")
    _mod("detectron2")
    _mod("detectron2.layers", Conv2d=Conv2d, ShapeSpec=ShapeSpec, get_norm=get_norm, move_device_like=move_device_like,
         CNNBlockBase=CNNBlockBase, batched_nms=batched_nms)
    _mod("detectron2.structures", Boxes=Boxes, Instances=Instances, ImageList=ImageList, BitMasks=BitMasks)
    _mod("detectron2.modeling", GeneralizedRCNN=object, detector_postprocess=detector_postprocess)
    _mod("detectron2.modeling.backbone", Backbone=Backbone)
    _mod("detectron2.modeling.backbone.fpn", _assert_strides_are_log2_contiguous=_assert_strides_are_log2_contiguous,
         LastLevelMaxPool=LastLevelMaxPool)
    _mod("detectron2.modeling.meta_arch")
    _mod("detectron2.modeling.meta_arch.panoptic_fpn",
         combine_semantic_and_instance_outputs=_unavailable("combine_semantic_and_instance_outputs"))
    _mod("detectron2.modeling.postprocessing", detector_postprocess=detector_postprocess,
         sem_seg_postprocess=sem_seg_postprocess)
    _mod("detectron2.modeling.roi_heads")
    _mod("detectron2.modeling.roi_heads.fast_rcnn", fast_rcnn_inference=_unavailable("d2 fast_rcnn_inference"))
    _mod("detectron2.utils")
    _mod("detectron2.utils.events", get_event_storage=_unavailable("get_event_storage"))
    _mod("detectron2.utils.memory", retry_if_cuda_oom=retry_if_cuda_oom)
    _mod("detectron2.data")
    _mod("detectron2.data.detection_utils", convert_image_to_rgb=_unavailable("convert_image_to_rgb"))
    _mod("detectron2.data.catalog", MetadataCatalog=MetadataCatalog)
    _mod("detrex")
    _mod("detrex.layers", MLP=MLP, FFN=FFN, BaseTransformerLayer=BaseTransformerLayer,
         MultiheadAttention=MultiheadAttention, TransformerLayerSequence=TransformerLayerSequence,
         PositionEmbeddingSine=PositionEmbeddingSine, box_cxcywh_to_xyxy=box_cxcywh_to_xyxy,
         box_xyxy_to_cxcywh=box_xyxy_to_cxcywh)
    _mod("detrex.utils", inverse_sigmoid=inverse_sigmoid)
    _mod("detrex.modeling")
    _mod("detrex.modeling.neck", ChannelMapper=ChannelMapper)
    _mod("fvcore")
    _mod("fvcore.nn")
    _mod("fvcore.nn.weight_init", c2_xavier_fill=c2_xavier_fill, c2_msra_fill=c2_msra_fill)
    _mod("timm")
    _mod("timm.models")
    _mod("timm.models.layers", DropPath=DropPath)

    ape = _pkg("ape", os.path.join(REF, "ape"))
    _mod("ape._C")  # keeps the real MSDA class (multi_scale_deform_attn.py:415-423)
    _pkg("ape.layers", os.path.join(REF, "ape", "layers"))
    _pkg("ape.modeling", os.path.join(REF, "ape", "modeling"))
    _pkg("ape.modeling.ape_deta", os.path.join(REF, "ape", "modeling", "ape_deta"))
    _pkg("ape.modeling.backbone", os.path.join(REF, "ape", "modeling", "backbone"))
    _pkg("ape.modeling.text", os.path.join(REF, "ape", "modeling", "text"))
    _pkg("ape.utils", os.path.join(REF, "ape", "utils"))
    # what `from ape.layers import X` expects (ape/layers/__init__.py re-exports these)
    layers = sys.modules["ape.layers"]
    for sub, names in (("multi_scale_deform_attn", ["MultiScaleDeformableAttention", "multi_scale_deformable_attn_pytorch"]),
                       ("vision_language_align", ["VisionLanguageAlign"]),
                       ("vision_language_fusion", ["VisionLanguageFusion"]),
                       ("zero_shot_fc", ["ZeroShotFC"])):
        m = importlib.import_module(f"ape.layers.{sub}")
        for n in names:
            setattr(layers, n, getattr(m, n))
    _installed = True
    return ape


def load(name):
    install()
    return importlib.import_module(name)
