"""Autograd layer over :mod:`spconv_b200.pytorch.ops`.

Public entry points keep the reference's names and positional argument order
(``spconv/pytorch/functional.py:423-429``): ``indice_conv``, ``indice_inverse_conv``,
``indice_subm_conv``, ``implicit_gemm``.  One generic :class:`_NativeConv` serves the three
ConvAlgo.Native variants (they differ only in the ``inverse`` / ``subm`` flags,
reference classes at ``functional.py:59-189,293-357``).
"""
from __future__ import annotations

import sys
from typing import List, Optional

import numpy as np
import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from ..core import Activation
from . import ops
from .core import CUDAKernelTimer

# AMP: inputs are cast to fp16 inside autocast regions, like the reference (functional.py:44-56)
_amp_fwd = torch.amp.custom_fwd(cast_inputs=torch.float16, device_type="cuda")
_amp_bwd = torch.amp.custom_bwd(device_type="cuda")


def _report(tag: str, **shapes) -> None:
    """Reference convention: print a one-line context to stderr before re-raising."""
    info = ",".join(f"{k}={v}" for k, v in shapes.items())
    print(f"[Exception|{tag}]{info}", file=sys.stderr)


class _NativeConv(Function):
    """features, filters, indice_pairs, indice_pair_num, num_activate_out, algo, timer, bias,
    act_alpha, act_beta, act_type, inverse, subm"""

    @staticmethod
    @_amp_fwd
    def forward(ctx, features, filters, indice_pairs, indice_pair_num, num_activate_out, algo,
                timer, bias, act_alpha, act_beta, act_type, inverse, subm):
        ctx.save_for_backward(indice_pairs, indice_pair_num, features, filters)
        ctx.spx = (algo, timer, inverse, subm)
        ctx.spx_scope = timer.snapshot()
        try:
            return ops.indice_conv(features, filters, indice_pairs, indice_pair_num,
                                   num_activate_out, inverse, subm, algo=algo, timer=timer,
                                   bias=bias, act_alpha=act_alpha, act_beta=act_beta,
                                   act_type=act_type)
        except Exception:
            _report("indice_conv", feat=tuple(features.shape), w=tuple(filters.shape),
                    pair=tuple(indice_pairs.shape), act=num_activate_out, algo=algo,
                    inverse=inverse, subm=subm)
            raise

    @staticmethod
    @once_differentiable
    @_amp_bwd
    def backward(ctx, grad_output):
        indice_pairs, indice_pair_num, features, filters = ctx.saved_tensors
        algo, timer, inverse, subm = ctx.spx
        try:
            with timer.scoped(ctx.spx_scope):
                din, dw = ops.indice_conv_backward(features, filters, grad_output, indice_pairs,
                                                   indice_pair_num, inverse, subm, algo=algo,
                                                   timer=timer)
        except Exception:
            _report("indice_conv_backward", feat=tuple(features.shape), w=tuple(filters.shape),
                    pair=tuple(indice_pairs.shape), do=tuple(grad_output.shape))
            raise
        return (din, dw) + (None,) * 11


class SparseImplicitGemmFunction(Function):
    """Masked implicit GEMM with autograd (reference ``functional.py:191-290``)."""

    @staticmethod
    @_amp_fwd
    def forward(ctx, features: torch.Tensor, filters: torch.Tensor, pair_fwd: torch.Tensor,
                pair_bwd: torch.Tensor, pair_mask_fwd_splits: List[torch.Tensor],
                pair_mask_bwd_splits: List[torch.Tensor],
                mask_argsort_fwd_splits: List[torch.Tensor],
                mask_argsort_bwd_splits: List[torch.Tensor], num_activate_out: int,
                masks: List[np.ndarray], is_train: bool, is_subm: bool,
                timer: CUDAKernelTimer = CUDAKernelTimer(False),
                fp32_accum: Optional[bool] = None, bias: Optional[torch.Tensor] = None,
                act_alpha: float = 0.0, act_beta: float = 0.0, act_type=Activation.None_):
        try:
            out, mask_out, mask_width = ops.implicit_gemm(
                features, filters, pair_fwd, pair_mask_fwd_splits, mask_argsort_fwd_splits,
                num_activate_out, masks, is_train, is_subm, timer, fp32_accum, bias, act_alpha,
                act_beta, act_type)
        except Exception:
            _report("implicit_gemm", feat=tuple(features.shape), w=tuple(filters.shape),
                    pair=tuple(pair_fwd.shape), act=num_activate_out, issubm=is_subm,
                    istrain=is_train)
            raise
        ctx.save_for_backward(features, filters, pair_fwd, pair_bwd)
        ctx.spx = dict(mask_width=mask_width, mask_out=mask_out, timer=timer, masks=masks,
                       scope=timer.snapshot(),
                       is_subm=is_subm, fp32_accum=fp32_accum,
                       mask_fwd=pair_mask_fwd_splits, mask_bwd=pair_mask_bwd_splits,
                       sort_fwd=mask_argsort_fwd_splits, sort_bwd=mask_argsort_bwd_splits)
        return out

    @staticmethod
    @once_differentiable
    @_amp_bwd
    def backward(ctx, grad_output):
        features, filters, pair_fwd, pair_bwd = ctx.saved_tensors
        s = ctx.spx
        try:
            with s["timer"].scoped(s["scope"]):
                din, dw = ops.implicit_gemm_backward(
                    features, filters, grad_output, pair_fwd, pair_bwd, s["mask_fwd"], s["mask_bwd"],
                    s["sort_fwd"], s["sort_bwd"], mask_output_fwd=s["mask_out"], masks=s["masks"],
                    mask_width=s["mask_width"], is_subm=s["is_subm"], timer=s["timer"],
                    fp32_accum=s["fp32_accum"])
        except Exception:
            _report("implicit_gemm_backward", feat=tuple(features.shape), w=tuple(filters.shape),
                    pair=tuple(pair_fwd.shape), issubm=s["is_subm"], do=tuple(grad_output.shape))
            raise
        return (din, dw) + (None,) * 16


class SparseMaxPoolFunction(Function):
    """ConvAlgo.Native max pooling (reference ``functional.py:360-378``)."""

    @staticmethod
    @_amp_fwd
    def forward(ctx, features, indice_pairs, indice_pair_num, num_activate_out):
        out = ops.indice_maxpool(features, indice_pairs, indice_pair_num, num_activate_out)
        ctx.save_for_backward(indice_pairs, indice_pair_num, features, out)
        return out

    @staticmethod
    @once_differentiable
    @_amp_bwd
    def backward(ctx, grad_output):
        indice_pairs, indice_pair_num, features, out = ctx.saved_tensors
        return ops.indice_maxpool_backward(features, out, grad_output, indice_pairs,
                                           indice_pair_num), None, None, None


class SparseMaxPoolImplicitGemmFunction(Function):
    """Reference ``functional.py:381-400``."""

    @staticmethod
    @_amp_fwd
    def forward(ctx, features, indice_pairs_fwd, indice_pairs_bwd, num_activate_out):
        out = ops.indice_maxpool_implicit_gemm(features, indice_pairs_fwd, num_activate_out)
        ctx.save_for_backward(indice_pairs_bwd, features, out)
        return out

    @staticmethod
    @once_differentiable
    @_amp_bwd
    def backward(ctx, grad_output):
        indice_pairs_bwd, features, out = ctx.saved_tensors
        return ops.indice_maxpool_implicit_gemm_backward(features, out, grad_output,
                                                         indice_pairs_bwd), None, None, None


class SparseAvgPoolImplicitGemmFunction(Function):
    """Reference ``functional.py:403-423``."""

    @staticmethod
    @_amp_fwd
    def forward(ctx, features, indice_pairs_fwd, indice_pairs_bwd, num_activate_out, calc_count):
        out, count = ops.indice_avgpool_implicit_gemm(features, indice_pairs_fwd, num_activate_out,
                                                      calc_count)
        ctx.save_for_backward(indice_pairs_bwd, count)
        return out

    @staticmethod
    @once_differentiable
    @_amp_bwd
    def backward(ctx, grad_output):
        indice_pairs_bwd, count = ctx.saved_tensors
        return ops.indice_avgpool_implicit_gemm_backward(grad_output, indice_pairs_bwd,
                                                         count), None, None, None, None


def _native(features, filters, indice_pairs, indice_pair_num, num_activate_out, algo, timer, bias,
            act_alpha, act_beta, act_type, inverse, subm):
    if timer is None:
        timer = CUDAKernelTimer(False)
    return _NativeConv.apply(features, filters, indice_pairs, indice_pair_num, num_activate_out,
                             algo, timer, bias, act_alpha, act_beta, act_type, inverse, subm)


def indice_conv(features, filters, indice_pairs, indice_pair_num, num_activate_out, algo,
                timer=None, bias=None, act_alpha=0.0, act_beta=0.0, act_type=Activation.None_):
    return _native(features, filters, indice_pairs, indice_pair_num, num_activate_out, algo, timer,
                   bias, act_alpha, act_beta, act_type, False, False)


def indice_inverse_conv(features, filters, indice_pairs, indice_pair_num, num_activate_out, algo,
                        timer=None, bias=None, act_alpha=0.0, act_beta=0.0,
                        act_type=Activation.None_):
    return _native(features, filters, indice_pairs, indice_pair_num, num_activate_out, algo, timer,
                   bias, act_alpha, act_beta, act_type, True, False)


def indice_subm_conv(features, filters, indice_pairs, indice_pair_num, num_activate_out, algo,
                     timer=None, bias=None, act_alpha=0.0, act_beta=0.0,
                     act_type=Activation.None_):
    return _native(features, filters, indice_pairs, indice_pair_num, num_activate_out, algo, timer,
                   bias, act_alpha, act_beta, act_type, False, True)


implicit_gemm = SparseImplicitGemmFunction.apply
indice_maxpool = SparseMaxPoolFunction.apply
indice_maxpool_implicit_gemm = SparseMaxPoolImplicitGemmFunction.apply
indice_avgpool_implicit_gemm = SparseAvgPoolImplicitGemmFunction.apply
