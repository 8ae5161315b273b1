"""B200-native (sm_100a) implementation of the multimodal forward/backward hot path of
IamJerryXu/Multimodal-Diagnosis-HAM-Spine, behind the reference's own nn.Module constructors."""
from . import _lib  # noqa: F401
from .encoder import ImageEncoder, TextEncoder  # noqa: F401
from .model import MultimodalBaselineModel  # noqa: F401

__all__ = ["ImageEncoder", "TextEncoder", "MultimodalBaselineModel"]
