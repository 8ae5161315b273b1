"""ConNeXT family (ConNexT/models/): KAN (block/kan1.py), sparsely-gated MoE head (block/moe.py) and the
ConvNeXt + BERT cross-attention classifier (ourmodel.py) on the B200 kernels."""
from .kan1 import KAN1, KANLinear  # noqa: F401
from .moe import MoE  # noqa: F401
