from .multi_scale_deform_attn import MultiScaleDeformableAttention  # noqa: F401
