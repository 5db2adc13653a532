"""Eval-time BatchNorm / activation folding for sparse convs -- the reference's "fused BN/act"
(``example/fuse_bn_act.py:36-86``, same math in ``spconv/pytorch/quantization/utils.py:5-52``).

BatchNorm folds into the KRSC filter and the bias; the activation becomes the conv's ``act_type`` and
then runs in the implicit-GEMM kernel epilogue together with the bias (``spx_implicit_gemm_fwd``),
so a ``SubMConv3d -> BatchNorm1d -> ReLU`` block is ONE kernel launch at inference time.
"""
from __future__ import annotations

import copy
from typing import Optional, Tuple

import torch
from torch import nn

from ..core import Activation
from .conv import SparseConvolution
from .modules import SparseSequential


def fuse_bn_weights(conv_w_OKI: torch.Tensor, conv_b: Optional[torch.Tensor], bn_rm: torch.Tensor,
                    bn_rv: torch.Tensor, bn_eps: float, bn_w: Optional[torch.Tensor],
                    bn_b: Optional[torch.Tensor]) -> Tuple[nn.Parameter, nn.Parameter]:
    """``W'[k] = W[k] * gamma[k] / sqrt(var[k] + eps)``, ``b' = (b - mean) * gamma / sqrt(var + eps) + beta``
    for a KRSC filter ``[K, *ksize, C]`` (the output channel is the LEADING axis, so no permutes)."""
    if conv_b is None:
        conv_b = torch.zeros_like(bn_rm)
    if bn_w is None:
        bn_w = torch.ones_like(bn_rm)
    if bn_b is None:
        bn_b = torch.zeros_like(bn_rm)
    scale = bn_w * torch.rsqrt(bn_rv + bn_eps)
    w = conv_w_OKI * scale.reshape([-1] + [1] * (conv_w_OKI.ndim - 1)).to(conv_w_OKI.dtype)
    b = (conv_b - bn_rm) * scale + bn_b
    return nn.Parameter(w.contiguous()), nn.Parameter(b.to(conv_w_OKI.dtype))


fuse_spconv_bn_weights = fuse_bn_weights          # name used by spconv/pytorch/quantization/utils.py


def fuse_bn(conv: SparseConvolution, bn: nn.modules.batchnorm._BatchNorm) -> SparseConvolution:
    """A conv ``C`` with ``C(x) == bn(conv(x))`` in inference mode (``fuse_bn_act.py:58-70``)."""
    assert not (conv.training or bn.training), "Fusion only for eval!"
    fused = copy.deepcopy(conv)
    fused.weight, fused.bias = fuse_bn_weights(fused.weight, fused.bias, bn.running_mean, bn.running_var,
                                               bn.eps, bn.weight, bn.bias)
    return fused


fuse_spconv_bn_eval = fuse_bn


def fuse_act(conv: SparseConvolution, act: nn.Module) -> SparseConvolution:
    """Moves the activation into the conv's kernel epilogue (``fuse_bn_act.py:72-86``)."""
    assert not conv.training, "Fusion only for eval!"
    fused = copy.deepcopy(conv)
    if isinstance(act, nn.ReLU):
        fused.act_type = Activation.ReLU
    elif isinstance(act, nn.Sigmoid):
        fused.act_type = Activation.Sigmoid
    elif isinstance(act, nn.LeakyReLU):
        fused.act_type = Activation.LeakyReLU
        fused.act_alpha = act.negative_slope
    else:
        raise NotImplementedError(type(act))
    return fused


fuse_act_net = fuse_act
fuse_spconv_act_eval = fuse_act


def fuse_bn_act_sequential(net: SparseSequential) -> SparseSequential:
    """Folds every ``conv -> BatchNorm1d [-> ReLU | LeakyReLU | Sigmoid]`` run of an eval-mode
    :class:`SparseSequential` (recursively).  The reference does this with a torch.fx graph rewrite
    (``fuse_bn_act.py:97-160``); sequential containers are what SECOND-style encoders use."""
    assert not net.training, "Fusion only for eval!"
    mods = list(net._modules.items())
    out = []
    i = 0
    while i < len(mods):
        name, m = mods[i]
        if isinstance(m, SparseSequential):
            out.append((name, fuse_bn_act_sequential(m)))
            i += 1
            continue
        if isinstance(m, SparseConvolution) and not m.conv1x1:
            j = i + 1
            if j < len(mods) and isinstance(mods[j][1], nn.modules.batchnorm._BatchNorm):
                m = fuse_bn(m, mods[j][1])
                j += 1
            if (j < len(mods) and isinstance(mods[j][1], (nn.ReLU, nn.LeakyReLU, nn.Sigmoid))
                    and m.act_type == Activation.None_):
                m = fuse_act(m, mods[j][1])
                j += 1
            out.append((name, m))
            i = j
            continue
        out.append((name, m))
        i += 1
    from collections import OrderedDict
    return SparseSequential(OrderedDict(out)).eval()
