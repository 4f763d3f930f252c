from .multi_scale_deform_attn import MultiScaleDeformableAttention  # noqa: F401
from .vision_language_align import VisionLanguageAlign  # noqa: F401
from .vision_language_fusion import VisionLanguageFusion  # noqa: F401
