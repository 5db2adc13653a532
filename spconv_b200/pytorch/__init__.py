"""``import spconv_b200.pytorch as spconv`` -- the ``spconv.pytorch`` surface of the hot path."""
from ..core import Activation, AlgoHint, ConvAlgo  # noqa: F401
from . import functional, ops  # noqa: F401
from .conv import (SparseConv1d, SparseConv2d, SparseConv3d, SparseConv4d,  # noqa: F401
                   SparseConvolution, SparseConvTranspose1d, SparseConvTranspose2d,
                   SparseConvTranspose3d, SparseConvTranspose4d, SparseInverseConv1d,
                   SparseInverseConv2d, SparseInverseConv3d, SparseInverseConv4d, SubMConv1d,
                   SubMConv2d, SubMConv3d, SubMConv4d)
from .core import (CUDAKernelTimer, ImplicitGemmIndiceData, IndiceData,  # noqa: F401
                   SparseConvTensor, scatter_nd)
from .modules import (RemoveGrid, SparseBatchNorm, SparseIdentity, SparseModule,  # noqa: F401
                      SparseReLU, SparseSequential, ToDense, assign_name_for_sparse_modules)
from .pool import (SparseAvgPool1d, SparseAvgPool2d, SparseAvgPool3d, SparseGlobalAvgPool,  # noqa: F401
                   SparseGlobalMaxPool, SparseMaxPool1d, SparseMaxPool2d, SparseMaxPool3d, SparseMaxPool4d)
from .utils_fuse import (fuse_act, fuse_bn, fuse_bn_act_sequential, fuse_bn_weights)  # noqa: F401
from . import quantized  # noqa: F401
from .graph import GraphedStep, graph_capture  # noqa: F401
from .utils import PointToVoxel, gather_features_by_pc_voxel_id  # noqa: F401
from .prefetch import RulebookPrefetcher  # noqa: F401
