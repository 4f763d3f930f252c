"""oracle/ref_model.py — TEST INFRASTRUCTURE, NOT PRODUCT CODE (build container only).

Builds the REFERENCE's own model objects (DeformableDETRSegmVL + ViT/SimpleFeaturePyramid + VL
transformer), executed unmodified from /root/reference under oracle/refshim.py, from one of the
plain-dict specs in ape_b200/configs.py — i.e. what detectron2's `instantiate(cfg.model)` would
build from configs/…/ape_deta_vitl_eva02_clip_vlf_lsj1024_cp_16x4_1080k.py, with the reference's
portable recipe (README.md:136-145, demo/app.py:680-694): pytorch_attn=True, xattn=False,
fp32 text features, no activation checkpointing.  Used by tests/golden/gen_model_golden.py.
"""
from functools import partial

import torch
import torch.nn as nn

from oracle import refshim


class FakeCriterion(nn.Module):
    """Stands in for DeformableCriterion: only the two attributes __init__/forward read
    (deformable_detr.py:102,283-291)."""

    loss_class_type = "focal_loss"

    def __init__(self, num_classes):
        super().__init__()
        self.num_classes = num_classes


class FakeLanguageModel:
    """model_language.forward_text stand-in: seeded text features (randn(N_t, C), seed 2)."""

    def __init__(self, lang_dim, n_max=8192, seed=2):
        g = torch.Generator().manual_seed(seed)
        self.bank = torch.randn(n_max, lang_dim, generator=g)

    def forward_text(self, text_list, cache=False):
        return {"last_hidden_state_eot": self.bank[: len(text_list)].clone()}


def randomize_degenerate_parameters(model, seed=1):
    """The reference zero-initialises several matrices (multi_scale_deform_attn.py:194,208-209,
    deformable_detr.py:119-120) and sets norms to identity; a parity test on such weights is
    degenerate.  Deterministically perturb every constant tensor (sorted by name, one generator)."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in sorted(model.state_dict().items()):
            if not p.is_floating_point() or p.numel() < 2 or "name_prompt_fusion_feature" in name:
                continue
            if "features_phrase_bank" in name or "freqs_" in name or "pixel_" in name:
                continue
            if bool((p == p.flatten()[0]).all()):
                p.add_(torch.randn(p.shape, generator=g) * 0.02)


def build_reference_model(spec, num_text=None, seed=0, test_mask_on=False, semantic_on=False):
    """-> (model_vision in eval mode, class-name list of length num_text)"""
    refshim.install()
    ti = spec["backbone"].get("variant", "eva_clip") == "eva02"
    vit_mod = refshim.load("ape.modeling.backbone.vit_eva02" if ti else "ape.modeling.backbone.vit_eva_clip")
    tr = refshim.load("ape.modeling.ape_deta.deformable_transformer_vl")
    segm = refshim.load("ape.modeling.ape_deta.deformable_detr_segm_vl")
    from ape.layers import VisionLanguageFusion  # noqa  (reference class, via refshim)

    torch.manual_seed(seed)
    b = spec["backbone"]
    net = vit_mod.ViT(
        img_size=b["img_size"], patch_size=b["patch_size"], embed_dim=b["embed_dim"], depth=b["depth"],
        num_heads=b["num_heads"], drop_path_rate=0.0, window_size=b["window_size"], mlp_ratio=b["mlp_ratio"],
        qkv_bias=True, norm_layer=partial(nn.LayerNorm, eps=1e-6), window_block_indexes=b["window_block_indexes"],
        residual_block_indexes=[], use_rel_pos=True, out_feature="last_feat", use_act_checkpoint=False, xattn=False,
        rope=True, pt_hw_seq_len=b["pt_hw_seq_len"], intp_freq=True, naiveswiglu=not ti, subln=not ti,
        pretrain_img_size=b["pretrain_img_size"], pretrain_use_cls_token=True, **({"swiglu": True} if ti else {}))
    backbone = vit_mod.SimpleFeaturePyramid(
        net=net, in_feature="last_feat", out_channels=b["out_channels"], scale_factors=b["scale_factors"],
        top_block=refshim.LastLevelMaxPool(), norm="LN", square_pad=b["square_pad"])
    E = spec["embed_dim"]
    feats = ["p2", "p3", "p4", "p5", "p6"]
    shapes = {f: refshim.ShapeSpec(channels=b["out_channels"]) for f in feats}
    neck = refshim.ChannelMapper(input_shapes=shapes, in_features=feats, out_channels=E, num_outs=5, kernel_size=1,
                                 norm_layer=nn.GroupNorm(num_groups=spec["gn_groups"], num_channels=E))
    vl_layer = VisionLanguageFusion(v_dim=E, l_dim=spec["lang_dim"], embed_dim=spec["vlf_embed"],
                                    num_heads=spec["vlf_heads"], dropout=0.1, drop_path=0.0,
                                    init_values=spec["vlf_init"], stable_softmax_2d=True, clamp_min_for_underflow=True,
                                    clamp_max_for_overflow=True, use_checkpoint=False)
    transformer = tr.DeformableDetrTransformerVL(
        encoder=tr.DeformableDetrTransformerEncoderVL(
            embed_dim=E, num_heads=spec["num_heads"], feedforward_dim=spec["ffn_dim"], attn_dropout=0.0,
            ffn_dropout=0.0, num_layers=spec["enc_layers"], post_norm=False, num_feature_levels=spec["num_levels"],
            vl_layer=vl_layer, use_act_checkpoint=False, pytorch_attn=True),
        decoder=tr.DeformableDetrTransformerDecoderVL(
            embed_dim=E, num_heads=spec["num_heads"], feedforward_dim=spec["ffn_dim"], attn_dropout=0.0,
            ffn_dropout=0.0, num_layers=spec["dec_layers"], return_intermediate=True,
            num_feature_levels=spec["num_levels"], pytorch_attn=True),
        as_two_stage=True, num_feature_levels=spec["num_levels"], two_stage_num_proposals=spec["num_queries"],
        assign_first_stage=True, pre_nms_topk=spec["pre_nms_topk"], nms_thresh_enc=spec["nms_thresh_enc"],
        proposal_ambiguous=spec["proposal_ambiguous"])
    n_text = num_text if num_text is not None else spec["num_classes"]
    names = [f"c{i}" for i in range(n_text)]
    meta = refshim.MetadataCatalog.get(f"fake_{spec['name']}_{n_text}")
    meta.thing_classes = names
    model = segm.DeformableDETRSegmVL(
        instance_on=True, semantic_on=semantic_on, panoptic_on=False, input_shapes=shapes, mask_in_features=["p2"],
        mask_encode_level=0, stuff_dataset_learn_thing=False, stuff_prob_thing=0.9, name_prompt_fusion_type="zero",
        test_mask_on=test_mask_on,
        backbone=backbone, position_embedding=refshim.PositionEmbeddingSine(num_pos_feats=E // 2, temperature=10000,
                                                                           normalize=True, offset=-0.5),
        neck=neck, transformer=transformer, embed_dim=E, num_classes=spec["num_classes"],
        num_queries=spec["num_queries"], criterion=[FakeCriterion(spec["num_classes"])],
        pixel_mean=list(spec["pixel_mean"]), pixel_std=list(spec["pixel_std"]), aux_loss=True, with_box_refine=True,
        as_two_stage=True, select_box_nums_for_evaluation=spec["test_topk"], input_format="RGB",
        dataset_names=[meta.name], dataset_metas=[meta.name], dataset_prompts=["name"],
        embed_dim_language=spec["lang_dim"], text_feature_bank=True, text_feature_reduce_before_fusion=True,
        text_feature_batch_repeat=True, expression_cumulative_gt_class=True,
        test_nms_thresh=spec["test_nms_thresh"], test_score_thresh=spec["test_score_thresh"])
    model.set_model_language(FakeLanguageModel(spec["lang_dim"]))
    randomize_degenerate_parameters(model)
    model.eval()
    return model, names


def synthetic_image(h, w, seed=0):
    """randint(0,256) as float32, CHW RGB."""
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, 256, (3, h, w), generator=g).to(torch.float32)
