"""Algorithm enums of the reference (``spconv/core.py:25-35``) plus the activation enum the
reference takes from ``cumm.tensorview.gemm.Activation``."""
from enum import Enum


class ConvAlgo(Enum):
    Native = 0
    MaskImplicitGemm = 1
    MaskSplitImplicitGemm = 2


class AlgoHint(Enum):
    NoHint = 0b000
    Fowrard = 0b001          # (sic) spelling kept from the reference
    BackwardInput = 0b010
    BackwardWeight = 0b100


class Activation(Enum):
    """``tv.gemm.Activation`` (values match include/spconv_b200.h ``spx_act``)."""
    None_ = 0
    ReLU = 1
    Sigmoid = 2
    LeakyReLU = 3
