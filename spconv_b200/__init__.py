"""spconv_b200 -- B200-native (sm_100a) drop-in for the rulebook -> implicit-GEMM hot path of
traveller59/spconv.  ``import spconv_b200.pytorch as spconv`` mirrors ``import spconv.pytorch``.
"""
from .core import Activation, AlgoHint, ConvAlgo  # noqa: F401

__version__ = "0.1.0"
