from .attention import MultiHeadCrossAttention_v2, SelfAttention, compute_kl_divergence  # noqa: F401
from .bert import BertEncoder  # noqa: F401
from .model_resnet import Resnet50WithOurs  # noqa: F401
