"""Operator layer: the drop-in boundary of the hot path.

Same function names, argument order and return shapes as ``spconv/pytorch/ops.py``
(``get_indice_pairs`` :132, ``get_indice_pairs_implicit_gemm`` :329, ``indice_conv`` :811,
``indice_conv_backward`` :1103, ``implicit_gemm`` :1450, ``implicit_gemm_backward`` :1667),
implemented as thin calls into the C-ABI library (``include/spconv_b200.h``) on the current
CUDA stream.  CUDA tensors only: there is deliberately no CPU path in the product.

Differences a reference user can observe (all documented in DESIGN.md):
  * rulebooks are deterministic and bit-equal to the reference's CPU order;
  * ``mask_width`` is always 128 (the tcgen05 tile height);
  * ``ConvAlgo.MaskSplitImplicitGemm`` builds the reference's two mask splits (offsets
    ``[0, kv - kv//2)`` and the rest, ``ops.py:494-503``), sorts each on its own and runs one kernel
    pass per split, summing the partial outputs (the reference accumulates with beta = 1).
"""
from __future__ import annotations

import ctypes
import functools
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from .. import _cabi
from ..constants import SPCONV_ALLOW_TF32, SPCONV_B200_FUSED_RULEBOOK, SPCONV_DO_SORT
from ..core import Activation, ConvAlgo
from .core import CUDAKernelTimer, ThrustSortAllocator

INT32_MAX = 2147483647
MASK_WIDTH = 128

_DTYPE_CODE = {
    torch.float32: _cabi.SPX_F32,
    torch.float16: _cabi.SPX_F16,
    torch.bfloat16: _cabi.SPX_BF16,
    torch.int8: _cabi.SPX_I8,
}


# ---------------------------------------------------------------------------- small helpers
def _lib():
    return _cabi.load()


_raw_stream = torch._C._cuda_getCurrentRawStream
_cur_device = torch._C._cuda_getDevice


def _stream() -> int:
    """cudaStream_t of torch's current stream on the current device.  (``torch.cuda.current_stream()``
    costs ~4 us of Python per call -- 76 calls per encoder step were a fifth of the eager host time.)"""
    return _raw_stream(_cur_device())


def _prod(xs) -> int:
    r = 1
    for x in xs:
        r *= int(x)
    return r


_SIDE_STREAMS: dict = {}


def _side_stream(device: torch.device) -> "torch.cuda.Stream":
    """One auxiliary stream per device for the input-gradient kernel (see implicit_gemm_backward)."""
    key = torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device()
    st = _SIDE_STREAMS.get(key)
    if st is None:
        st = _SIDE_STREAMS[key] = torch.cuda.Stream(device=key)
    return st


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    if t is None or t.numel() == 0:
        return None
    return t.data_ptr()


def _require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(
            f"spconv_b200: {what} must be a CUDA tensor. This engine has no CPU path "
            "(the CPU restatement under oracle/ is test infrastructure only).")


def _bytes(n: int, device, alloc: Optional[ThrustSortAllocator] = None) -> torch.Tensor:
    if alloc is not None:
        return alloc.get(n)
    return torch.empty(max(int(n), 1), dtype=torch.uint8, device=device)


def _act_code(act_type) -> int:
    if act_type is None:
        return _cabi.SPX_ACT_NONE
    if isinstance(act_type, Activation):
        return act_type.value
    return int(getattr(act_type, "value", act_type))


def get_conv_output_size(input_size, kernel_size, stride, padding, dilation):
    """``(in + 2p - d(k-1) - 1) // s + 1`` per axis; ``k == -1`` means global (size 1)."""
    return [1 if k == -1 else (i + 2 * p - d * (k - 1) - 1) // s + 1
            for i, k, s, p, d in zip(input_size, kernel_size, stride, padding, dilation)]


def get_deconv_output_size(input_size, kernel_size, stride, padding, dilation, output_padding):
    out = []
    for i, k, s, p, op in zip(input_size, kernel_size, stride, padding, output_padding):
        if k == -1:
            raise ValueError("deconv don't support kernel_size < 0")
        out.append((i - 1) * s - 2 * p + k + op)
    return out


_VANISHED = ("Your points vanished here, this usually because you provide conv params that "
             "may ignore some input points. Example: spatial_shape=[8, 200, 200] -> "
             "[3, 3, 3] kernel, [2, 2, 2] stride, no padding -> out shape [3, 99, 99]; "
             "points in z=7/8 are dropped when no point lies in z<=6.")


def _out_shape(spatial_shape, ksize, stride, padding, dilation, out_padding, subm, transpose):
    if subm:
        shape = list(spatial_shape)
    elif transpose:
        shape = get_deconv_output_size(spatial_shape, ksize, stride, padding, dilation, out_padding)
    else:
        shape = get_conv_output_size(spatial_shape, ksize, stride, padding, dilation)
    if any(x <= 0 for x in shape):
        raise ValueError(
            f"your out spatial shape {shape} reach zero!!! input shape: {spatial_shape}")
    return shape


_GEO_CACHE: dict = {}


def _geometry(indices, batch_size, spatial_shape, out_shape, ksize, stride, padding, dilation,
              transpose):
    ndim = indices.shape[1] - 1
    if not (1 <= ndim <= _cabi.SPX_MAX_NDIM):
        raise RuntimeError(f"unsupported ndim {ndim}")
    key = (ndim, int(batch_size), tuple(spatial_shape), tuple(out_shape), tuple(ksize), tuple(stride),
           tuple(padding), tuple(dilation), bool(transpose))
    geo = _GEO_CACHE.get(key)
    if geo is None:                      # the struct is immutable once built: layers re-use it every step
        if len(_GEO_CACHE) > 4096:
            _GEO_CACHE.clear()
        geo = _GEO_CACHE[key] = _cabi.make_geometry(ndim, batch_size, spatial_shape, out_shape, ksize, stride,
                                                    padding, dilation, transpose)
    return geo


_ZERO_COUNTS: dict = {}


def _zero_counts(kv: int, device) -> torch.Tensor:
    """``indice_num_per_loc`` of the implicit-GEMM rulebooks: not consumed by the GEMM, kept
    for the shape of the reference's 9-tuple; one read-only zeros tensor per (kv, device)."""
    key = (kv, device)
    t = _ZERO_COUNTS.get(key)
    if t is None:
        t = _ZERO_COUNTS[key] = torch.zeros((kv,), dtype=torch.int32, device=device)
    return t


def _conv_rulebook(geo, indices, n_in, kv, words, want_masks, alloc):
    """Two-phase regular-conv rulebook; returns (out_inds, pair_fwd, pair_bwd, mask_fwd, mask_bwd)."""
    lib = _lib()
    dev = indices.device
    ws_bytes = lib.spx_rulebook_workspace_size(ctypes.byref(geo), n_in, 0, 0)
    ws = _bytes(ws_bytes, dev, alloc)
    m_host = ctypes.c_int64(0)
    _cabi.check(lib.spx_conv_rulebook_stage1(ctypes.byref(geo), _ptr(indices), n_in,
                                             ctypes.byref(m_host), ws.data_ptr(), ws.numel(),
                                             _stream()), "conv_rulebook_stage1")
    m = int(m_host.value)
    if m == 0:
        raise ValueError(_VANISHED)
    ndim = indices.shape[1] - 1
    out_inds = torch.empty((m, ndim + 1), dtype=torch.int32, device=dev)
    pair_fwd = torch.empty((kv, m), dtype=torch.int32, device=dev)
    pair_bwd = torch.empty((kv, n_in), dtype=torch.int32, device=dev)
    mask_fwd = torch.empty((1, m, words), dtype=torch.int32, device=dev) if want_masks else None
    mask_bwd = torch.empty((1, n_in, words), dtype=torch.int32, device=dev) if want_masks else None
    _cabi.check(lib.spx_conv_rulebook_stage2(ctypes.byref(geo), _ptr(indices), n_in, m,
                                             out_inds.data_ptr(), pair_fwd.data_ptr(),
                                             pair_bwd.data_ptr(), _ptr(mask_fwd), _ptr(mask_bwd),
                                             ws.data_ptr(), ws.numel(), _stream()),
                "conv_rulebook_stage2")
    return out_inds, pair_fwd, pair_bwd, mask_fwd, mask_bwd


def _conv_rulebook_all(geo, indices, n_in, kv, words, is_train, do_sort, alloc, indice_num_per_loc, masks):
    """Regular-conv implicit-GEMM rulebook in two native calls (stage 1 with its host read-back of the
    output count, then stage 2 + both mask sorts + both tile tables)."""
    lib = _lib()
    dev = indices.device
    ws = _bytes(lib.spx_conv_rulebook_all_workspace_size(ctypes.byref(geo), n_in), dev, alloc)
    m_host = ctypes.c_int64(0)
    _cabi.check(lib.spx_conv_rulebook_stage1(ctypes.byref(geo), _ptr(indices), n_in, ctypes.byref(m_host),
                                             ws.data_ptr(), ws.numel(), _stream()), "conv_rulebook_stage1")
    m = int(m_host.value)
    if m == 0:
        raise ValueError(_VANISHED)
    ndim = indices.shape[1] - 1
    out_inds = torch.empty((m, ndim + 1), dtype=torch.int32, device=dev)
    pair_fwd = torch.empty((kv, m), dtype=torch.int32, device=dev)
    pair_bwd = torch.empty((kv, n_in), dtype=torch.int32, device=dev)
    mask_fwd = torch.empty((1, m, words), dtype=torch.int32, device=dev)
    mask_bwd = torch.empty((1, n_in, words), dtype=torch.int32, device=dev)
    sort_fwd = torch.empty((1, m), dtype=torch.int32, device=dev)
    sort_bwd = torch.empty((1, n_in), dtype=torch.int32, device=dev) if is_train else None
    t_fwd, tm_fwd = _alloc_tile_tables(m, kv, dev)
    t_bwd, tm_bwd = _alloc_tile_tables(n_in, kv, dev) if is_train else (None, None)
    _cabi.check(lib.spx_conv_rulebook_stage2_all(
        ctypes.byref(geo), _ptr(indices), n_in, m, out_inds.data_ptr(), pair_fwd.data_ptr(), pair_bwd.data_ptr(),
        mask_fwd.data_ptr(), mask_bwd.data_ptr(), sort_fwd.data_ptr(), _ptr(sort_bwd), int(bool(do_sort)),
        _ptr(t_fwd), _ptr(tm_fwd), _ptr(t_bwd), _ptr(tm_bwd), ws.data_ptr(), ws.numel(), _stream()),
        "conv_rulebook_stage2_all")
    sf = sort_fwd[0]
    sf._spx_tile_cache = (_tile_key(pair_fwd, sf, m), t_fwd, tm_fwd)
    if not is_train:
        return (out_inds, indice_num_per_loc, pair_fwd, pair_bwd, [mask_fwd[0]], [], [sf], [], masks)
    sb = sort_bwd[0]
    sb._spx_tile_cache = (_tile_key(pair_bwd, sb, n_in), t_bwd, tm_bwd)
    return (out_inds, indice_num_per_loc, pair_fwd, pair_bwd, [mask_fwd[0]], [mask_bwd[0]], [sf], [sb], masks)


def _argsort_masks(mask: torch.Tensor, kv: int, do_sort: bool, alloc) -> torch.Tensor:
    """mask [1, n, words] -> argsort [1, n]; mask is left sorted (thrust::sort_by_key semantics,
    ``spconv/csrc/sparse/all.py:935-1000``)."""
    lib = _lib()
    n, words = mask.shape[1], mask.shape[2]
    argsort = torch.empty((1, n), dtype=torch.int32, device=mask.device)
    if n == 0:
        return argsort
    ws_bytes = lib.spx_mask_argsort_workspace_size(n, words)
    ws = _bytes(ws_bytes, mask.device, alloc)
    _cabi.check(lib.spx_mask_argsort(mask.data_ptr(), argsort.data_ptr(), n, words, kv,
                                     int(bool(do_sort)), ws.data_ptr(), ws.numel(), _stream()),
                "mask_argsort")
    return argsort


def _split_and_sort(mask: torch.Tensor, masks: List[np.ndarray], kv: int, do_sort: bool, alloc):
    """mask [1, n, 1] (unsorted) -> per split j: (mask & masks[j]) sorted in place + its argsort
    (``ops.py:494-503,538-549``: every split is sorted on its own)."""
    mask_s, sort_s = [], []
    for m in masks:
        const = torch.from_numpy(m.view(np.int32).copy()).to(mask.device)
        part = torch.bitwise_and(mask, const.view(1, 1, 1)).contiguous()
        sort_s.append(_argsort_masks(part, kv, do_sort, alloc)[0])
        mask_s.append(part[0])
    return mask_s, sort_s


# ---------------------------------------------------------------------------- rulebooks
def get_indice_pairs(indices: torch.Tensor, batch_size: int, spatial_shape: List[int],
                     algo: ConvAlgo, ksize: List[int], stride: List[int], padding: List[int],
                     dilation: List[int], out_padding: List[int], subm: bool = False,
                     transpose: bool = False, num_out_act_bound: int = -1):
    """ConvAlgo.Native rulebook: ``(out_inds [M, ndim+1], pairs [2, kv, N], indice_pair_num [kv])``
    with pair ORDER equal to the reference CPU implementation
    (``spconv/csrc/sparse/indices.py:1640-1778``)."""
    _require_cuda(indices, "indices")
    lib = _lib()
    dev = indices.device
    indices = indices.contiguous()
    n_in = indices.shape[0]
    kv = _prod(ksize)
    out_shape = _out_shape(spatial_shape, ksize, stride, padding, dilation, out_padding, subm,
                           transpose)
    geo = _geometry(indices, batch_size, spatial_shape, out_shape, ksize, stride, padding,
                    dilation, transpose)
    pairs = torch.empty((2, kv, n_in), dtype=torch.int32, device=dev)
    num = torch.empty((kv,), dtype=torch.int32, device=dev)
    if subm:
        for k in ksize:
            if k % 2 != 1:
                raise RuntimeError("subm only support odd ksize")
        pair_fwd = torch.empty((kv, n_in), dtype=torch.int32, device=dev)
        pair_bwd = torch.empty((kv, n_in), dtype=torch.int32, device=dev)
        ws = _bytes(lib.spx_rulebook_workspace_size(ctypes.byref(geo), n_in, 0, 1), dev)
        _cabi.check(lib.spx_subm_rulebook(ctypes.byref(geo), _ptr(indices), n_in,
                                          pair_fwd.data_ptr(), pair_bwd.data_ptr(), None, None,
                                          ws.data_ptr(), ws.numel(), _stream()), "subm_rulebook")
        out_inds = indices
    else:
        out_inds, pair_fwd, pair_bwd, _, _ = _conv_rulebook(geo, indices, n_in, kv,
                                                            (kv + 31) // 32, False, None)
    ws2 = _bytes(lib.spx_native_pairs_workspace_size(n_in, kv), dev)
    _cabi.check(lib.spx_native_pairs(_ptr(pair_bwd), n_in, kv, int(subm), pairs.data_ptr(),
                                     num.data_ptr(), ws2.data_ptr(), ws2.numel(), _stream()),
                "native_pairs")
    return out_inds, pairs, num


def get_indice_pairs_implicit_gemm(indices: torch.Tensor, batch_size: int,
                                   spatial_shape: List[int], algo: ConvAlgo, ksize: List[int],
                                   stride: List[int], padding: List[int], dilation: List[int],
                                   out_padding: List[int], subm: bool = False,
                                   transpose: bool = False, is_train: bool = True,
                                   alloc: Optional[ThrustSortAllocator] = None,
                                   timer: CUDAKernelTimer = CUDAKernelTimer(False),
                                   num_out_act_bound: int = -1,
                                   direct_table: bool = True,
                                   do_sort=SPCONV_DO_SORT):
    """Masked implicit-GEMM rulebook.  Returns the reference's 9-tuple
    ``(out_inds, indice_num_per_loc, pair_fwd, pair_bwd, [mask_fwd], [mask_bwd],
    [argsort_fwd], [argsort_bwd], masks)`` (``ops.py:329-359``)."""
    _require_cuda(indices, "indices")
    assert algo in (ConvAlgo.MaskImplicitGemm, ConvAlgo.MaskSplitImplicitGemm), "TODO"
    lib = _lib()
    dev = indices.device
    indices = indices.contiguous()
    n_in = indices.shape[0]
    kv = _prod(ksize)
    words = (kv + 31) // 32
    if kv > 128:
        raise NotImplementedError("masked implicit gemm supports kernel volume <= 128")
    out_shape = _out_shape(spatial_shape, ksize, stride, padding, dilation, out_padding, subm,
                           transpose)
    geo = _geometry(indices, batch_size, spatial_shape, out_shape, ksize, stride, padding,
                    dilation, transpose)
    is_split = algo == ConvAlgo.MaskSplitImplicitGemm
    if is_split:
        assert words == 1, "Not Implemented"                    # reference: ops.py:495
        remain = kv - kv // 2
        masks = [np.array([(1 << remain) - 1], dtype=np.uint32),
                 np.array([((1 << (kv // 2)) - 1) << remain], dtype=np.uint32)]
    else:
        masks = [np.array([0xffffffff], dtype=np.uint32)]
    # per-offset pair counts are not consumed by the GEMM; kept for API shape
    indice_num_per_loc = _zero_counts(kv, dev)
    if subm:
        for k in ksize:
            if k % 2 != 1:
                raise RuntimeError("subm only support odd ksize")
        pair = torch.empty((2 if is_train else 1, kv, n_in), dtype=torch.int32, device=dev)
        pair_mask = torch.empty((1, n_in, words), dtype=torch.int32, device=dev)
        if SPCONV_B200_FUSED_RULEBOOK and not timer.enable and not is_split and n_in:
            # one native call: hash + probe + mask sort + tile table (the separate calls below are kept
            # for profiling regions and the mask-split algo)
            mask_argsort = torch.empty((1, n_in), dtype=torch.int32, device=dev)
            table, tile_mask = _alloc_tile_tables(n_in, kv, dev)
            ws = _bytes(lib.spx_subm_rulebook_all_workspace_size(ctypes.byref(geo), n_in), dev, alloc)
            _cabi.check(lib.spx_subm_rulebook_all(ctypes.byref(geo), _ptr(indices), n_in, pair[0].data_ptr(),
                                                  pair[1].data_ptr() if is_train else None, _ptr(pair_mask),
                                                  _ptr(mask_argsort), int(bool(do_sort)), _ptr(table), _ptr(tile_mask),
                                                  ws.data_ptr(), ws.numel(), _stream()), "subm_rulebook_all")
            pair_fwd, argsort_view = pair[0], mask_argsort[0]
            argsort_view._spx_tile_cache = (_tile_key(pair_fwd, argsort_view, n_in), table, tile_mask)
            return (indices, indice_num_per_loc, pair_fwd, pair[1] if is_train else torch.Tensor(),
                    [pair_mask[0]], [], [argsort_view], [], masks)
        # row-major by-product of the probe kernel; consumed (and dropped) by the first tile-table build
        rows = None
        if n_in and lib.spx_subm_row_table_supported(ctypes.byref(geo)):
            rows = torch.empty((n_in, 32), dtype=torch.int32, device=dev)
        with timer.record("gen_subm_inds", _stream()):
            ws = _bytes(lib.spx_rulebook_workspace_size(ctypes.byref(geo), n_in, 0, 1), dev, alloc)
            _cabi.check(lib.spx_subm_rulebook(ctypes.byref(geo), _ptr(indices), n_in,
                                              pair[0].data_ptr(),
                                              pair[1].data_ptr() if is_train and n_in else None,
                                              _ptr(pair_mask), _ptr(rows), ws.data_ptr(), ws.numel(),
                                              _stream()), "subm_rulebook")
        pair_bwd = pair[1] if is_train else torch.Tensor()
        if is_split:
            with timer.record("gen_subm_inds_sort", _stream()):
                mask_s, sort_s = _split_and_sort(pair_mask, masks, kv, do_sort, alloc)
            return (indices, indice_num_per_loc, pair[0], pair_bwd, mask_s, [], sort_s, [], masks)
        with timer.record("gen_subm_inds_sort", _stream()):
            mask_argsort = _argsort_masks(pair_mask, kv, do_sort, alloc)
        argsort_view = mask_argsort[0]
        if rows is not None:
            argsort_view._spx_row_table = (pair[0].data_ptr(), rows)
        return (indices, indice_num_per_loc, pair[0], pair_bwd, [pair_mask[0]], [],
                [argsort_view], [], masks)
    if SPCONV_B200_FUSED_RULEBOOK and not timer.enable and not is_split and n_in:
        return _conv_rulebook_all(geo, indices, n_in, kv, words, is_train, do_sort, alloc, indice_num_per_loc, masks)
    with timer.record("gen_conv_inds", _stream()):
        out_inds, pair_fwd, pair_bwd, mask_fwd, mask_bwd = _conv_rulebook(
            geo, indices, n_in, kv, words, True, alloc)
    if is_split:
        with timer.record("gen_conv_inds_sort", _stream()):
            mf, sf = _split_and_sort(mask_fwd, masks, kv, do_sort, alloc)
            mb, sb = _split_and_sort(mask_bwd, masks, kv, do_sort, alloc) if is_train else ([], [])
        return (out_inds, indice_num_per_loc, pair_fwd, pair_bwd, mf, mb, sf, sb, masks)
    with timer.record("gen_conv_inds_sort", _stream()):
        argsort_fwd = _argsort_masks(mask_fwd, kv, do_sort, alloc)
        if is_train:
            argsort_bwd = _argsort_masks(mask_bwd, kv, do_sort, alloc)
    if is_train:
        return (out_inds, indice_num_per_loc, pair_fwd, pair_bwd, [mask_fwd[0]], [mask_bwd[0]],
                [argsort_fwd[0]], [argsort_bwd[0]], masks)
    return (out_inds, indice_num_per_loc, pair_fwd, pair_bwd, [mask_fwd[0]], [], [argsort_fwd[0]],
            [], masks)


# ---------------------------------------------------------------------------- GEMM descriptor
def _f32_mode() -> int:
    return _cabi.SPX_F32_TF32 if SPCONV_ALLOW_TF32 else _cabi.SPX_F32_EXACT


def _tile_key(pair: torch.Tensor, argsort: Optional[torch.Tensor], rows: int):
    return (pair.data_ptr(), tuple(pair.shape), pair._version,
            None if argsort is None else (argsort.data_ptr(), argsort._version), int(rows))


def _alloc_tile_tables(rows: int, kv: int, device):
    tiles = max((int(rows) + MASK_WIDTH - 1) // MASK_WIDTH, 1)
    # layout of include/spconv_b200.h: blocks + schedule records + scheduler scratch (== spx_tile_table_elems)
    table = torch.empty((tiles * (kv + 1) * 128 + tiles * 8 + 64,), dtype=torch.int32, device=device)
    tile_mask = torch.empty((tiles, (kv + 31) // 32), dtype=torch.int32, device=device)
    return table, tile_mask


def _tile_tables(pair: torch.Tensor, mask: Optional[torch.Tensor], argsort: Optional[torch.Tensor],
                 rows: int, kv: int, owner: Optional[torch.Tensor] = None):
    """Tile-blocked gather table + per-tile OR masks for (pair, mask, argsort)
    (``spx_build_tile_table``).  Built once per rulebook and cached on ``owner`` (the argsort
    tensor that lives in the cached ``ImplicitGemmIndiceData``), so forward, input-gradient and
    weight-gradient of every layer sharing the ``indice_key`` reuse it."""
    key = _tile_key(pair, argsort, rows)
    if owner is not None:
        hit = getattr(owner, "_spx_tile_cache", None)
        if hit is not None and hit[0] == key:
            return hit[1], hit[2]
    lib = _lib()
    table, tile_mask = _alloc_tile_tables(rows, kv, pair.device)
    row_table = None
    hint = getattr(owner, "_spx_row_table", None) if owner is not None else None
    if hint is not None and hint[0] == pair.data_ptr() and hint[1].shape[0] == int(rows):
        row_table = hint[1]
    _cabi.check(lib.spx_build_tile_table(_ptr(pair), int(pair.stride(0)), kv, _ptr(argsort), _ptr(mask),
                                         int(rows), _ptr(row_table), _ptr(table), _ptr(tile_mask), _stream()),
                "build_tile_table")
    if hint is not None:
        owner._spx_row_table = None                # one-shot: 128 B per voxel are not kept alive
    if owner is not None:
        owner._spx_tile_cache = (key, table, tile_mask)
    return table, tile_mask


def _desc(dtype, kv, c_in, c_out, n_in, n_out, pair, mask, argsort, reverse=False, tiles=None):
    d = _cabi.GemmDesc()
    d.dtype = _DTYPE_CODE[dtype]
    d.f32_mode = _f32_mode()
    d.kv, d.c_in, d.c_out = int(kv), int(c_in), int(c_out)
    d.n_in, d.n_out = int(n_in), int(n_out)
    d.pair = _ptr(pair)
    d.pair_stride = int(pair.stride(0)) if pair is not None and pair.dim() == 2 else 0
    d.mask = _ptr(mask)
    d.argsort = _ptr(argsort)
    d.reverse_offsets = int(bool(reverse))
    if tiles is not None:
        d.tile_table = _ptr(tiles[0])
        d.tile_mask = _ptr(tiles[1])
    return d


def _check_filter(features, filters):
    if filters.dtype != features.dtype:
        raise RuntimeError(f"features ({features.dtype}) and filters ({filters.dtype}) must have the same dtype")
    if features.dtype not in _DTYPE_CODE:
        raise RuntimeError(f"unsupported dtype {features.dtype}")
    kv = _prod(filters.shape[1:-1])
    return kv, int(filters.shape[-1]), int(filters.shape[0])


def _first(split_list):
    if isinstance(split_list, (list, tuple)):
        return split_list[0] if len(split_list) else None
    return split_list


# ---------------------------------------------------------------------------- masked implicit GEMM
def implicit_gemm(features: torch.Tensor, filters: torch.Tensor, pair_fwd: torch.Tensor,
                  pair_mask_fwd_splits: List[torch.Tensor],
                  mask_argsort_fwd_splits: List[torch.Tensor], num_activate_out: int,
                  masks: List[np.ndarray], is_train: bool, is_subm: bool,
                  timer: CUDAKernelTimer = CUDAKernelTimer(False),
                  fp32_accum: Optional[bool] = None, bias: Optional[torch.Tensor] = None,
                  act_alpha: float = 0.0, act_beta: float = 0.0,
                  act_type=Activation.None_, output_scale: float = 1.0,
                  scale: Optional[torch.Tensor] = None, output_add: Optional[torch.Tensor] = None,
                  output_add_scale: float = 0.0, output_dtype: Optional[torch.dtype] = None):
    """Forward masked implicit GEMM -> ``(out [M, K], mask_output_fwd, mask_width)``
    (``ops.py:1450-1469`` / ``convops.py:2075-2243``).  Accumulation is always fp32 in TMEM
    (``fp32_accum`` is accepted and ignored)."""
    _require_cuda(features, "features")
    lib = _lib()
    features = features.contiguous()
    filters = filters.contiguous()
    kv, c_in, c_out = _check_filter(features, filters)
    assert features.shape[1] == c_in, "channel size mismatch"
    n_in, n_out = features.shape[0], int(num_activate_out)
    n_splits = len(pair_mask_fwd_splits) if isinstance(pair_mask_fwd_splits, (list, tuple)) else 1
    if n_splits > 1:
        return _implicit_gemm_splits(features, filters, pair_fwd, pair_mask_fwd_splits,
                                     mask_argsort_fwd_splits, n_out, is_train, timer, bias, act_alpha,
                                     act_type, output_add, output_dtype)
    mask = _first(pair_mask_fwd_splits)
    argsort = _first(mask_argsort_fwd_splits)
    is_int8 = features.dtype == torch.int8
    if output_dtype is None:
        output_dtype = features.dtype
    words = (kv + 31) // 32
    with timer.record("tile_table", _stream()):
        tiles = _tile_tables(pair_fwd, mask, argsort, n_out, kv, owner=argsort) if n_out else None
    # mask_output_fwd (per-128-row OR of the sorted masks) is the tile table's mask block
    mask_output = tiles[1].view(1, -1, words) if (is_train and tiles is not None) else torch.Tensor()
    d = _desc(features.dtype, kv, c_in, c_out, n_in, n_out, pair_fwd, mask, argsort, tiles=tiles)
    if is_int8:
        assert scale is not None, "int8 implicit gemm needs the per-channel scale"
        out = torch.empty((n_out, c_out), dtype=output_dtype, device=features.device)
        # reference int8 epilogue (convops.py:2176-2206): per-channel `scale` multiplies the int32
        # accumulator, the residual enters with beta = output_add_scale / output_scale
        scale_f = scale.float().contiguous()
        bias_f = bias.float().contiguous() if bias is not None else None
        with timer.record("implicit_gemm_int8", _stream()):
            _cabi.check(lib.spx_implicit_gemm_fwd_int8(
                ctypes.byref(d), _ptr(features), _ptr(filters), _ptr(out), _DTYPE_CODE[output_dtype],
                _ptr(scale_f), _ptr(bias_f), _ptr(output_add),
                float(output_add_scale) / float(output_scale if output_scale else 1.0),
                _act_code(act_type), float(act_alpha), _stream()), "implicit_gemm_fwd_int8")
        return out, mask_output, MASK_WIDTH
    out = torch.empty((n_out, c_out), dtype=features.dtype, device=features.device)
    if bias is not None:
        bias = bias.to(features.dtype).contiguous()
    with timer.record("implicit_gemm", _stream()):
        _cabi.check(lib.spx_implicit_gemm_fwd(ctypes.byref(d), _ptr(features), _ptr(filters),
                                              _ptr(out), _ptr(bias), _act_code(act_type),
                                              float(act_alpha), None, _stream()), "implicit_gemm_fwd")
    if output_add is not None:
        out = out + output_add
    if output_dtype != out.dtype:
        out = out.to(output_dtype)
    return out, mask_output, MASK_WIDTH


def _implicit_gemm_splits(features, filters, pair_fwd, mask_splits, argsort_splits, n_out, is_train, timer,
                          bias, act_alpha, act_type, output_add, output_dtype):
    """ConvAlgo.MaskSplitImplicitGemm forward: one kernel pass per mask split (each visits only its
    own offsets, rows in that split's sorted order), partial outputs summed, bias / activation after
    the last split (the reference fuses them into the last pass, ``convops.py:2196-2234``)."""
    lib = _lib()
    if features.dtype == torch.int8:
        raise NotImplementedError("int8 + MaskSplitImplicitGemm: use ConvAlgo.MaskImplicitGemm")
    kv, c_in, c_out = _check_filter(features, filters)
    n_in = features.shape[0]
    words = (kv + 31) // 32
    out = None
    tile_masks = []
    for mask, argsort in zip(mask_splits, argsort_splits):
        tiles = _tile_tables(pair_fwd, mask, argsort, n_out, kv, owner=argsort) if n_out else None
        d = _desc(features.dtype, kv, c_in, c_out, n_in, n_out, pair_fwd, mask, argsort, tiles=tiles)
        part = torch.empty((n_out, c_out), dtype=features.dtype, device=features.device)
        with timer.record("implicit_gemm", _stream()):
            _cabi.check(lib.spx_implicit_gemm_fwd(ctypes.byref(d), _ptr(features), _ptr(filters), _ptr(part),
                                                  None, _cabi.SPX_ACT_NONE, 0.0, None, _stream()),
                        "implicit_gemm_fwd(split)")
        out = part if out is None else out.add_(part)
        if tiles is not None:
            tile_masks.append(tiles[1].view(1, -1, words))
    if (bias is not None or _act_code(act_type) != _cabi.SPX_ACT_NONE) and n_out:
        bias_add_act_inplace(out, bias, act_type, act_alpha)
    if output_add is not None:
        out = out + output_add
    if output_dtype is not None and output_dtype != out.dtype:
        out = out.to(output_dtype)
    mask_output = torch.cat(tile_masks, 0) if (is_train and tile_masks) else torch.Tensor()
    return out, mask_output, MASK_WIDTH


def implicit_gemm_backward(features: torch.Tensor, filters: torch.Tensor, out_bp: torch.Tensor,
                           pair_fwd: torch.Tensor, pair_bwd: torch.Tensor,
                           pair_mask_fwd_splits: List[torch.Tensor],
                           pair_mask_bwd_splits: List[torch.Tensor],
                           mask_argsort_fwd_splits: List[torch.Tensor],
                           mask_argsort_bwd_splits: List[torch.Tensor],
                           mask_output_fwd: Optional[torch.Tensor], masks: List[np.ndarray],
                           mask_width: int, is_subm: bool,
                           timer: CUDAKernelTimer = CUDAKernelTimer(False),
                           fp32_accum: Optional[bool] = None):
    """Input gradient + weight gradient of the masked implicit GEMM -> ``(din, dfilters)``
    (``ops.py:1667-1681`` / ``convops.py:2247-2440``)."""
    _require_cuda(features, "features")
    lib = _lib()
    features = features.contiguous()
    filters = filters.contiguous()
    out_bp = out_bp.contiguous()
    if out_bp.dtype != features.dtype:
        out_bp = out_bp.to(features.dtype)
    kv, c_in, c_out = _check_filter(features, filters)
    n_in, n_out = features.shape[0], out_bp.shape[0]
    n_splits = len(pair_mask_fwd_splits) if isinstance(pair_mask_fwd_splits, (list, tuple)) else 1
    if n_splits > 1:
        # MaskSplitImplicitGemm: gradients are sums over the splits (each split only visits its offsets)
        din = dfilters = None
        for j in range(n_splits):
            bwd_m = [pair_mask_bwd_splits[j]] if pair_mask_bwd_splits else []
            bwd_s = [mask_argsort_bwd_splits[j]] if mask_argsort_bwd_splits else []
            di, dw = implicit_gemm_backward(features, filters, out_bp, pair_fwd, pair_bwd,
                                            [pair_mask_fwd_splits[j]], bwd_m, [mask_argsort_fwd_splits[j]],
                                            bwd_s, None, masks, mask_width, is_subm, timer, fp32_accum)
            # dW of split j is only meaningful on ITS offsets: the tcgen05 kernel leaves the others zero, the
            # generic FMA kernel (odd channel counts) walks the whole pair table -- mask either way
            keep = torch.tensor([(int(masks[j][k >> 5]) >> (k & 31)) & 1 for k in range(kv)], dtype=dw.dtype,
                                device=dw.device).view(1, kv, 1)
            dw = dw.view(c_out, kv, c_in) * keep
            din = di if din is None else din.add_(di)
            dfilters = dw if dfilters is None else dfilters.add_(dw)
        return din, dfilters.view(filters.shape)
    din = torch.empty_like(features)
    dfilters = torch.empty_like(filters)
    mask_fwd, argsort_fwd = _first(pair_mask_fwd_splits), _first(mask_argsort_fwd_splits)
    tiles_fwd = _tile_tables(pair_fwd, mask_fwd, argsort_fwd, n_out, kv, owner=argsort_fwd) if n_out else None
    if is_subm:
        # SubM pairs are symmetric: walk the FORWARD table/mask and flip the filter offset
        # (the reference's reverse_mask, convops.py:2412)
        d_dg = _desc(features.dtype, kv, c_in, c_out, n_in, n_out, pair_fwd, mask_fwd, argsort_fwd,
                     reverse=True, tiles=tiles_fwd)
    else:
        mask_bwd, argsort_bwd = _first(pair_mask_bwd_splits), _first(mask_argsort_bwd_splits)
        tiles_bwd = _tile_tables(pair_bwd, mask_bwd, argsort_bwd, n_in, kv, owner=argsort_bwd) if n_in else None
        d_dg = _desc(features.dtype, kv, c_in, c_out, n_in, n_out, pair_bwd, mask_bwd, argsort_bwd,
                     tiles=tiles_bwd)
    d_wg = _desc(features.dtype, kv, c_in, c_out, n_in, n_out, pair_fwd, mask_fwd, argsort_fwd,
                 tiles=tiles_fwd)
    ws_bytes = lib.spx_implicit_gemm_wgrad_workspace_size(ctypes.byref(d_wg))
    ws = _bytes(ws_bytes, features.device)

    def run_dgrad():
        with timer.record("implicit_gemm_dgrad", _stream()):
            _cabi.check(lib.spx_implicit_gemm_dgrad(ctypes.byref(d_dg), _ptr(out_bp), _ptr(filters),
                                                    _ptr(din), _stream()), "implicit_gemm_dgrad")

    def run_wgrad():
        with timer.record("implicit_gemm_wgrad", _stream()):
            if _PEERS is not None and not (_PEER_TRIAGE & 2):
                # data-parallel: the kernel that reduces the split-K partials pushes this rank's fp32 dW into every
                # rank's exchange buffer (csrc/peer.cu); finish_exchange() below writes dfilters
                _cabi.check(lib.spx_implicit_gemm_wgrad_push(
                    ctypes.byref(d_wg), _ptr(features), _ptr(out_bp), _ptr(dfilters), ws.data_ptr(), ws.numel(),
                    ctypes.byref(_PEERS.group), _stream()), "implicit_gemm_wgrad_push")
            else:
                _cabi.check(lib.spx_implicit_gemm_wgrad(ctypes.byref(d_wg), _ptr(features), _ptr(out_bp),
                                                        _ptr(dfilters), ws.data_ptr(), ws.numel(),
                                                        _stream()), "implicit_gemm_wgrad")

    def finish_exchange():
        if _PEER_TRIAGE & 1:
            return
        with timer.record("implicit_gemm_wgrad_exchange", _stream()):
            _cabi.check(lib.spx_peer_finish(ctypes.byref(_PEERS.group), _ptr(dfilters), dfilters.numel(),
                                            _DTYPE_CODE[dfilters.dtype], _PEERS.scale, _stream()), "peer_finish")

    if _PEERS is not None:
        if timer.enable or not (n_in and n_out) or not torch._C._cuda_isCurrentStreamCapturing():
            # eager: one stream (a fork / join per layer costs more host time than it hides); the input gradient
            # between publish and finish hides the NVLink latency
            run_wgrad()
            run_dgrad()
            finish_exchange()
            return din, dfilters
        # captured: the two gradients stay parallel branches as in the single-GPU graph (measured: running them one
        # after the other costs 17 us per config-2 step, more than the whole exchange).  Weight gradient + publish
        # on the caller's stream, input gradient on the forked one, the receive side (wait, TMA pull, rank-order
        # sum: a few small CTAs) behind it: every dependent kernel on the caller's stream costs ~8 us of launch and
        # queueing when the next cloud's rulebook kernels share the GPU.
        main = torch.cuda.current_stream()
        side = _side_stream(features.device)
        side.wait_stream(main)
        run_wgrad()
        with torch.cuda.stream(side):
            run_dgrad()
            side.wait_stream(main)          # the publish
            finish_exchange()
        main.wait_stream(side)
        return din, dfilters
    if _WGRAD_HOOK is not None and n_in and n_out:
        # Data-parallel overlap: weight gradient FIRST, then the hook (typically the all-reduce of dW) on
        # a forked stream while the input gradient -- which the hook does not need -- runs on this one.
        main = torch.cuda.current_stream()
        side = _side_stream(features.device)
        run_wgrad()
        side.wait_stream(main)
        with torch.cuda.stream(side):
            _WGRAD_HOOK(dfilters)
        run_dgrad()
        main.wait_stream(side)
        return din, dfilters
    if timer.enable or not (n_in and n_out) or not torch._C._cuda_isCurrentStreamCapturing():
        # eager launches are host-bound (and an eager fork/join per call measured slower, not
        # faster); profiling regions stay one kernel each
        run_dgrad()
        run_wgrad()
        return din, dfilters
    # Under CUDA-graph capture the two independent gradients become parallel branches.  The weight-gradient kernel (one CTA per SM, statically
    # assigned tiles) has a long tail -- its CTAs finish over a ~15 us window -- so it goes first on
    # the caller's stream and the dynamically scheduled input-gradient kernel is queued on a forked
    # stream: its CTAs fill the SMs the weight gradient has already left.  Joined before returning.
    main = torch.cuda.current_stream()
    side = _side_stream(features.device)
    side.wait_stream(main)
    run_wgrad()
    with torch.cuda.stream(side):
        run_dgrad()
    main.wait_stream(side)
    return din, dfilters


_WGRAD_HOOK = None


def set_wgrad_hook(fn) -> None:
    """``fn(dfilters)`` is called on a forked stream right after every weight-gradient launch of
    :func:`implicit_gemm_backward`, while the input gradient of the same layer runs on the caller's
    stream (joined before the op returns).  This is the place for a data-parallel all-reduce of the
    layer's dW: it overlaps the rest of the backward pass instead of trailing it (DDP-style hook).
    ``None`` removes the hook."""
    global _WGRAD_HOOK
    _WGRAD_HOOK = fn


_PEERS = None
_PEER_TRIAGE = 0          # bench.py --peer-triage (timing experiments only; results are wrong when set)


def set_peer_group(peers) -> None:
    """Data-parallel mode: with a :class:`spconv_b200.pytorch.dist.PeerGroup` installed, every weight
    gradient computed by :func:`implicit_gemm_backward` / :func:`indice_conv_backward` is returned
    already summed (x ``peers.scale``) over the ranks -- the exchange is the tail of the
    weight-gradient kernel (NVLink peer stores, ``csrc/peer.cu``), there is no separate all-reduce.
    Every rank must run the same sequence of layers.  ``None`` switches it off."""
    global _PEERS
    _PEERS = peers


def peer_allreduce_(t: torch.Tensor) -> torch.Tensor:
    """In-place sum (x scale) of a small tensor over the installed peer group (bias gradients, weight
    gradients of layers that do not run the implicit-GEMM kernels); no-op without a group."""
    if _PEERS is None or t.numel() == 0:
        return t
    assert t.is_contiguous(), "peer_allreduce_: contiguous tensor"
    _cabi.check(_lib().spx_peer_allreduce(ctypes.byref(_PEERS.group), t.data_ptr(), t.numel(), _DTYPE_CODE[t.dtype],
                                          _PEERS.scale, _stream()), "peer_allreduce")
    return t


# ---------------------------------------------------------------------------- ConvAlgo.Native
def _native_tables(indice_pairs, indice_pair_num, n_in, n_out, kv, subm, inverse, need_fwd,
                   need_bwd):
    """compact pairs [2, kv, L] -> dense gather tables + row masks (visited in natural order)."""
    lib = _lib()
    dev = indice_pairs.device
    words = (kv + 31) // 32
    t_fwd = torch.empty((kv, n_out), dtype=torch.int32, device=dev) if need_fwd else None
    m_fwd = torch.empty((n_out, words), dtype=torch.int32, device=dev) if need_fwd else None
    t_bwd = torch.empty((kv, n_in), dtype=torch.int32, device=dev) if need_bwd else None
    m_bwd = torch.empty((n_in, words), dtype=torch.int32, device=dev) if need_bwd else None
    _cabi.check(lib.spx_pairs_to_table(_ptr(indice_pairs), _ptr(indice_pair_num), kv,
                                       int(indice_pairs.shape[2]), n_in, n_out, int(subm),
                                       int(inverse), _ptr(t_fwd), _ptr(t_bwd), _ptr(m_fwd),
                                       _ptr(m_bwd), _stream()), "pairs_to_table")
    return t_fwd, m_fwd, t_bwd, m_bwd


def indice_conv(features: torch.Tensor, filters: torch.Tensor, indice_pairs: torch.Tensor,
                indice_pair_num: torch.Tensor, num_activate_out: int, inverse: bool = False,
                subm: bool = False, algo: ConvAlgo = ConvAlgo.Native,
                timer: CUDAKernelTimer = CUDAKernelTimer(False),
                bias: Optional[torch.Tensor] = None, act_alpha: float = 0.0,
                act_beta: float = 0.0, act_type=Activation.None_):
    """Gather-GEMM-scatter forward over a compact rulebook (``ops.py:811-823`` /
    ``convops.py:1504-1747``).  The compact pairs are scattered into a dense gather table on the
    device (no ``indice_pair_num.cpu()`` sync) and the output-stationary implicit-GEMM kernel
    accumulates every offset in TMEM -- no atomics, deterministic."""
    _require_cuda(features, "features")
    lib = _lib()
    features = features.contiguous()
    filters = filters.contiguous()
    indice_pairs = indice_pairs.contiguous()
    kv, c_in, c_out = _check_filter(features, filters)
    assert features.shape[1] == c_in, "channel size mismatch"
    assert indice_pairs.shape[1] == kv, "indice_pairs / filter kernel volume mismatch"
    n_in, n_out = features.shape[0], int(num_activate_out)
    with timer.record("indice_conv_table", _stream()):
        t_fwd, m_fwd, _, _ = _native_tables(indice_pairs, indice_pair_num, n_in, n_out, kv, subm,
                                            inverse, True, False)
    out = torch.empty((n_out, c_out), dtype=features.dtype, device=features.device)
    if bias is not None:
        bias = bias.to(features.dtype).contiguous()
    tiles = _tile_tables(t_fwd, m_fwd, None, n_out, kv) if n_out else None
    d = _desc(features.dtype, kv, c_in, c_out, n_in, n_out, t_fwd, m_fwd, None, tiles=tiles)
    with timer.record("indice_conv", _stream()):
        _cabi.check(lib.spx_implicit_gemm_fwd(ctypes.byref(d), _ptr(features), _ptr(filters),
                                              _ptr(out), _ptr(bias), _act_code(act_type),
                                              float(act_alpha), None, _stream()),
                    "implicit_gemm_fwd(native)")
    return out


def indice_conv_backward(features: torch.Tensor, filters: torch.Tensor, out_bp: torch.Tensor,
                         indice_pairs: torch.Tensor, indice_pair_num: torch.Tensor,
                         inverse: bool = False, subm: bool = False,
                         algo: ConvAlgo = ConvAlgo.Native,
                         timer: CUDAKernelTimer = CUDAKernelTimer(False)):
    """Backward of :func:`indice_conv` -> ``(din, dfilters)`` (``ops.py:1103-1111`` /
    ``convops.py:1751-2071``)."""
    _require_cuda(features, "features")
    lib = _lib()
    features = features.contiguous()
    filters = filters.contiguous()
    out_bp = out_bp.contiguous()
    if out_bp.dtype != features.dtype:
        out_bp = out_bp.to(features.dtype)
    indice_pairs = indice_pairs.contiguous()
    kv, c_in, c_out = _check_filter(features, filters)
    n_in, n_out = features.shape[0], out_bp.shape[0]
    t_fwd, m_fwd, t_bwd, m_bwd = _native_tables(indice_pairs, indice_pair_num, n_in, n_out, kv,
                                                subm, inverse, True, True)
    din = torch.empty_like(features)
    dfilters = torch.empty_like(filters)
    tiles_bwd = _tile_tables(t_bwd, m_bwd, None, n_in, kv) if n_in else None
    tiles_fwd = _tile_tables(t_fwd, m_fwd, None, n_out, kv) if n_out else None
    d_dg = _desc(features.dtype, kv, c_in, c_out, n_in, n_out, t_bwd, m_bwd, None, tiles=tiles_bwd)
    with timer.record("indice_conv_dgrad", _stream()):
        _cabi.check(lib.spx_implicit_gemm_dgrad(ctypes.byref(d_dg), _ptr(out_bp), _ptr(filters),
                                                _ptr(din), _stream()), "implicit_gemm_dgrad(native)")
    d_wg = _desc(features.dtype, kv, c_in, c_out, n_in, n_out, t_fwd, m_fwd, None, tiles=tiles_fwd)
    ws = _bytes(lib.spx_implicit_gemm_wgrad_workspace_size(ctypes.byref(d_wg)), features.device)
    with timer.record("indice_conv_wgrad", _stream()):
        if _PEERS is not None:
            _cabi.check(lib.spx_implicit_gemm_wgrad_allreduce(
                ctypes.byref(d_wg), _ptr(features), _ptr(out_bp), _ptr(dfilters), ws.data_ptr(), ws.numel(),
                ctypes.byref(_PEERS.group), _PEERS.scale, _stream()), "implicit_gemm_wgrad_allreduce(native)")
        else:
            _cabi.check(lib.spx_implicit_gemm_wgrad(ctypes.byref(d_wg), _ptr(features), _ptr(out_bp),
                                                    _ptr(dfilters), ws.data_ptr(), ws.numel(),
                                                    _stream()), "implicit_gemm_wgrad(native)")
    return din, dfilters


# ---------------------------------------------------------------------------- pooling
_POOL_MAX, _POOL_MAX_ZERO_FLOOR, _POOL_MEAN = 0, 1, 2


def _pool_check(features: torch.Tensor):
    _require_cuda(features, "features")
    if features.dtype not in _DTYPE_CODE:
        raise RuntimeError(f"unsupported dtype {features.dtype}")
    c = features.shape[1]
    if (c * features.element_size()) % 16:
        raise RuntimeError(f"pooling needs channels * element size to be a multiple of 16 bytes, got {c} x "
                           f"{features.element_size()}")


def _pool_fwd(mode, features, table, n_out, count_out=None):
    features = features.contiguous()
    _pool_check(features)
    out = torch.empty((int(n_out), features.shape[1]), dtype=features.dtype, device=features.device)
    _cabi.check(_lib().spx_indice_pool_fwd(mode, _ptr(features), _ptr(out), _ptr(table), int(table.stride(0)),
                                           int(table.shape[0]), int(n_out), int(features.shape[1]),
                                           _DTYPE_CODE[features.dtype], _ptr(count_out), _stream()),
                "indice_pool_fwd")
    return out


def _pool_bwd(mode, features, out_features, out_bp, table_bwd, n_in, count_out=None):
    out_bp = out_bp.contiguous()
    _pool_check(out_bp)
    din = torch.empty((int(n_in), out_bp.shape[1]), dtype=out_bp.dtype, device=out_bp.device)
    _cabi.check(_lib().spx_indice_pool_bwd(mode, _ptr(features), _ptr(out_features), _ptr(out_bp), _ptr(din),
                                           _ptr(table_bwd), int(table_bwd.stride(0)), int(table_bwd.shape[0]),
                                           int(n_in), int(out_bp.shape[1]), _DTYPE_CODE[out_bp.dtype],
                                           _ptr(count_out), _stream()), "indice_pool_bwd")
    return din


def indice_maxpool(features: torch.Tensor, indice_pairs: torch.Tensor, indice_pair_num: torch.Tensor,
                   num_activate_out):
    """ConvAlgo.Native max pooling over compact pairs (``ops.py:1899-1936``).  The reference raises a
    zero-initialised output per offset, i.e. the result is ``max(0, max over inputs)``; the compact
    pairs are scattered into a dense table on the device (no ``indice_pair_num.cpu()`` sync) and
    one kernel reduces every output row."""
    kv = int(indice_pairs.shape[1])
    t_fwd, _, _, _ = _native_tables(indice_pairs.contiguous(), indice_pair_num, features.shape[0],
                                    int(num_activate_out), kv, False, False, True, False)
    return _pool_fwd(_POOL_MAX_ZERO_FLOOR, features, t_fwd, num_activate_out)


def indice_maxpool_backward(features, out_features, out_bp, indice_pairs, indice_pair_num):
    """``din[i] += dout[o]`` where ``x[i] == y[o]`` (``ops.py:1939-1972``)."""
    kv = int(indice_pairs.shape[1])
    _, _, t_bwd, _ = _native_tables(indice_pairs.contiguous(), indice_pair_num, features.shape[0],
                                    out_features.shape[0], kv, False, False, False, True)
    return _pool_bwd(_POOL_MAX, features.contiguous(), out_features.contiguous(), out_bp, t_bwd,
                     features.shape[0])


def indice_maxpool_implicit_gemm(features: torch.Tensor, indice_pairs: torch.Tensor, num_activate_out):
    """Max pooling through the dense forward table ``pair_fwd [kv, M]`` (``ops.py:1975-2006``)."""
    return _pool_fwd(_POOL_MAX, features, indice_pairs, num_activate_out)


def indice_maxpool_implicit_gemm_backward(features, out_features, out_bp, indice_pairs):
    """``indice_pairs`` is the backward table ``pair_bwd [kv, N]`` (``ops.py:2009-2030``)."""
    return _pool_bwd(_POOL_MAX, features.contiguous(), out_features.contiguous(), out_bp, indice_pairs,
                     features.shape[0])


def indice_avgpool_implicit_gemm(features: torch.Tensor, indice_pairs: torch.Tensor, num_activate_out,
                                 calc_count: bool):
    """Mean over the valid neighbours + their count (``ops.py:2033-2074``)."""
    count_out = torch.Tensor()
    if calc_count:
        count_out = torch.empty((int(num_activate_out),), dtype=torch.int32, device=features.device)
    out = _pool_fwd(_POOL_MEAN, features, indice_pairs, num_activate_out, count_out if calc_count else None)
    return out, count_out


def indice_avgpool_implicit_gemm_backward(out_bp, indice_pairs, count_out):
    """``din[i] = sum_o dout[o] * count[o]`` -- the reference multiplies by the neighbour count
    (``maxpool.py:262-300``); kept so gradients equal the reference's (``ops.py:2077-2096``)."""
    return _pool_bwd(_POOL_MEAN, None, None, out_bp, indice_pairs, indice_pairs.shape[1], count_out)


def global_pool_rearrange(coords: torch.Tensor, batch_size: int):
    """Row indices of every sample: ``(out_indices [batch, N], counts [batch])`` (``ops.py:2108-2124``)."""
    _require_cuda(coords, "coords")
    coords = coords.contiguous()
    n = coords.shape[0]
    out_indices = torch.empty((batch_size, n), dtype=torch.int32, device=coords.device)
    counts = torch.empty((batch_size,), dtype=torch.int32, device=coords.device)
    _cabi.check(_lib().spx_global_pool_rearrange(_ptr(coords), n, int(coords.shape[1]), int(batch_size),
                                                 _ptr(out_indices) if n else out_indices.data_ptr(),
                                                 counts.data_ptr(), _stream()), "global_pool_rearrange")
    return out_indices, counts


# ---------------------------------------------------------------------------- misc
def bias_add_act_inplace(x: torch.Tensor, bias: Optional[torch.Tensor], act_type=Activation.None_,
                         act_alpha: float = 0.0, act_beta: float = 0.0) -> torch.Tensor:
    """``InferenceOps.bias_add_act_inplace`` (``inference.py:166-252``)."""
    _require_cuda(x, "x")
    assert x.is_contiguous() and x.dim() == 2
    if bias is not None:
        bias = bias.to(x.dtype).contiguous()
    _cabi.check(_lib().spx_bias_act_inplace(_ptr(x), _ptr(bias), x.shape[0], x.shape[1],
                                            _DTYPE_CODE[x.dtype], _act_code(act_type),
                                            float(act_alpha), _stream()), "bias_act_inplace")
    return x


def maximum_value_int_(ten: torch.Tensor, value: int):
    """running max of the active-voxel count (``ops.py`` maximum_value_int_)."""
    ten.clamp_(min=int(value))
    return ten


def last_kernel_family() -> int:
    """0 none, 1 generic FMA kernels, 2 tcgen05 kernels (what served the last GEMM call)."""
    return int(_lib().spx_last_kernel_family())


def launch_count(reset: bool = False) -> int:
    return int(_lib().spx_launch_count(int(reset)))
