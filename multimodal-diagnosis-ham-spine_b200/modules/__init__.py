from .fusion_blocks import (  # noqa: F401
    BasicTransformerBlock, FusionModule, CrossAttentionBlock, MultiScaleFusionModule, ConcatFusionModule,
    WeightedConcatFusionModule, HadamardFusionModule, BilinearFusionModule)
from .heads import ResidualBlock, ResidualClassifier, AttentionPoolingClassifier, MLPHead, build_kan_head  # noqa: F401
from .gating import DualExpertGate  # noqa: F401
