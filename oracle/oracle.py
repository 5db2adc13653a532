"""CPU oracle for the rulebook -> gather-GEMM-scatter hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module; the product package
(``spconv_b200``) never does.

Parity status: **pinned to the reference itself.**  ``make_ref.py`` extracts the reference's own
CPU rulebook and gather/scatter C++ from /root/reference and compiles it into
``oracle/_ref/libspconv_ref.so``; ``tests/test_oracle_ref.py`` holds the C restatement
(``spconv_oracle.c``) to it bit for bit (pair ORDER included).  The conv loops below follow
``convops.py:1534-1633,1769-1860`` step by step -- ``GatherCPU::gather`` -> ``torch.mm`` ->
``GatherCPU::scatter_add`` with the reference's gather/scatter code when ``_ref`` is built -- and
are additionally pinned by the reference's dense-convolution equivalence test
(``tests/test_oracle.py``).  Unpinnable (no reference implementation exists): bf16, and the
implicit-GEMM tables, which the reference only builds on the GPU (derived per SURVEY A.5).

What is restated (paths relative to /root/reference):

* rulebooks ............ ``spconv_oracle.c`` (C, compiled with gcc by :func:`build`)
* buffer shapes / fill .. ``spconv/csrc/sparse/all.py:2064-2127`` (``get_indice_pairs``)
* "points vanished" .... ``spconv/pytorch/ops.py:54-70,260-262``
* CPU conv fwd ......... ``spconv/csrc/sparse/convops.py:1534-1633`` + mm callbacks
                         ``spconv/pytorch/cppcore.py:232-262`` (``buf @ W_k^T``)
* CPU conv bwd ......... ``spconv/csrc/sparse/convops.py:1769-1860`` + ``cppcore.py:290-348``
* SubM mirror rule ..... ``spconv/pytorch/ops.py:962-968`` (``nhot = num[kv-1-k]``)
* implicit-GEMM tables . derived (the reference builds them on GPU only,
                         ``spconv/csrc/sparse/indices.py:807-874,600-721``); SURVEY A.5
* int8 epilogue ........ ``test/test_all_algo.py:272-287``
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "spconv_oracle.c")
_LIB = os.path.join(_HERE, "_build", "libspconv_oracle.so")
_REF_LIB = os.path.join(_HERE, "_ref", "libspconv_ref.so")
_lib: Optional[ctypes.CDLL] = None
_ref: Optional[ctypes.CDLL] = None
_ref_tried = False


def build(force: bool = False) -> str:
    """Compile the C restatement with gcc (seconds)."""
    if (not force and os.path.exists(_LIB)
            and os.path.getmtime(_LIB) >= os.path.getmtime(_SRC)):
        return _LIB
    os.makedirs(os.path.dirname(_LIB), exist_ok=True)
    cmd = ["gcc", "-O2", "-fPIC", "-shared", "-std=c11", "-o", _LIB, _SRC]
    subprocess.run(cmd, check=True)
    return _LIB


def _load() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(_SRC):
            try:
                build()
            except Exception:
                if not os.path.exists(_LIB):
                    raise
        _lib = ctypes.CDLL(_LIB)
        _lib.orc_subm_rulebook.restype = ctypes.c_int
        _lib.orc_conv_rulebook.restype = ctypes.c_int
    return _lib


def build_ref(force: bool = False) -> Optional[str]:
    """Compile ``oracle/_ref/libspconv_ref.so`` -- the reference's OWN CPU rulebook and
    gather/scatter code, extracted from /root/reference by ``oracle/make_ref.py``.  Returns the
    path, or None when neither the reference tree nor a prebuilt library is present."""
    from . import make_ref
    return make_ref.build(force=force)


def ref_lib() -> Optional[ctypes.CDLL]:
    """The compiled reference code (``oracle/_ref``) or None.  On the GPU box /root/reference does
    not exist; the library built here travels with the snapshot."""
    global _ref, _ref_tried
    if _ref is None and not _ref_tried:
        _ref_tried = True
        path = None
        try:
            path = build_ref()
        except Exception:
            path = _REF_LIB if os.path.exists(_REF_LIB) else None
        if path and os.path.exists(path):
            _ref = ctypes.CDLL(path)
            _ref.ref_subm_rulebook.restype = ctypes.c_int
            _ref.ref_conv_rulebook.restype = ctypes.c_int
            _ref.ref_num_threads.restype = ctypes.c_int
    return _ref


def have_ref() -> bool:
    return ref_lib() is not None


def _iarr(v: Sequence[int]):
    return (ctypes.c_int * len(v))(*[int(x) for x in v])


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


def get_conv_output_size(input_size, kernel_size, stride, padding, dilation) -> List[int]:
    out = []
    for i in range(len(input_size)):
        if kernel_size[i] == -1:
            out.append(1)
        else:
            out.append((input_size[i] + 2 * padding[i] - dilation[i] * (kernel_size[i] - 1) - 1)
                       // stride[i] + 1)
    return out


def get_deconv_output_size(input_size, kernel_size, stride, padding, dilation,
                           output_padding) -> List[int]:
    out = []
    for i in range(len(input_size)):
        if kernel_size[i] == -1:
            raise ValueError("deconv don't support kernel_size < 0")
        out.append((input_size[i] - 1) * stride[i] - 2 * padding[i] + kernel_size[i]
                   + output_padding[i])
    return out


def get_indice_pairs(indices: np.ndarray, batch_size: int, spatial_shape: Sequence[int],
                     ksize: Sequence[int], stride: Sequence[int], padding: Sequence[int],
                     dilation: Sequence[int], out_padding: Sequence[int], subm: bool = False,
                     transpose: bool = False, impl: str = "port"
                     ) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Native rulebook in the reference's CPU order.

    Returns ``(out_inds [M, ndim+1], pairs [2, kv, N], indice_num_per_loc [kv])`` exactly
    as ``ops.get_indice_pairs`` does on a CPU tensor (``ops.py:132-170``).
    ``impl="port"``: the C restatement (``spconv_oracle.c``); ``impl="ref"``: the reference's own
    C++ (``oracle/_ref``, see ``make_ref.py``) -- ``tests/test_oracle_ref.py`` pins one to the other.
    """
    if impl == "ref":
        lib = ref_lib()
        if lib is None:
            raise RuntimeError("oracle/_ref is not built (needs /root/reference or a prebuilt library)")
        f_subm, f_conv = lib.ref_subm_rulebook, lib.ref_conv_rulebook
    else:
        lib = _load()
        f_subm, f_conv = lib.orc_subm_rulebook, lib.orc_conv_rulebook
    indices = np.ascontiguousarray(indices, dtype=np.int32)
    n, ndim = indices.shape[0], indices.shape[1] - 1
    kv = int(np.prod(ksize))
    if not subm:
        if transpose:
            out_shape = get_deconv_output_size(spatial_shape, ksize, stride, padding, dilation,
                                               out_padding)
        else:
            out_shape = get_conv_output_size(spatial_shape, ksize, stride, padding, dilation)
    else:
        out_shape = list(spatial_shape)
    if any(x == 0 for x in out_shape):
        raise ValueError(
            f"your out spatial shape {out_shape} reach zero!!! input shape: {spatial_shape}")
    pairs = np.full((2, kv, n), -1, dtype=np.int32)
    num = np.zeros((kv,), dtype=np.int32)
    if subm:
        ret = f_subm(_ptr(indices), n, ndim, int(batch_size), _iarr(spatial_shape),
                     _iarr(ksize), _iarr(dilation), _ptr(pairs), _ptr(num))
        if ret == -2:
            raise RuntimeError("subm only support odd ksize")
        if ret < 0:
            raise RuntimeError(f"oracle subm rulebook failed ({ret})")
        return indices, pairs, num
    out_inds = np.empty((max(kv * n, 1), ndim + 1), dtype=np.int32)
    num_act = f_conv(_ptr(indices), n, ndim, int(batch_size), _iarr(out_shape),
                     _iarr(spatial_shape), _iarr(ksize), _iarr(stride),
                     _iarr(padding), _iarr(dilation), int(bool(transpose)),
                     _ptr(pairs), _ptr(out_inds), _ptr(num))
    if num_act < 0:
        raise RuntimeError(f"oracle conv rulebook failed ({num_act})")
    if num_act == 0:
        raise ValueError("Your points vanished here, this usually because you provide "
                         "conv params that may ignore some input points. Example: "
                         "spatial_shape=[8, 200, 200] -> stride=2 conv")
    return out_inds[:num_act].copy(), pairs, num


def _pair_counts(num: np.ndarray, kv: int, n_in: int, subm: bool) -> np.ndarray:
    """valid length of pairs[:, k, :] for every k (SubM mirror rule, ops.py:962-968)."""
    cnt = np.array(num, dtype=np.int64).copy()
    if subm:
        for k in range(kv):
            if k > kv // 2:
                cnt[k] = num[kv - 1 - k]
            elif k == kv // 2:
                cnt[k] = n_in
    return cnt


def implicit_gemm_tables(pairs: np.ndarray, num: np.ndarray, n_in: int, n_out: int, subm: bool,
                         do_sort: bool = True):
    """Derive the masked-implicit-GEMM artefacts from a native rulebook.

    ``pair_fwd[k, o] = i``, ``pair_bwd[k, i] = o`` (-1 = none); ``mask_*`` has bit ``k % 32``
    of word ``k // 32`` set iff the entry exists; ``argsort_*`` is the STABLE ascending
    argsort of the masks and the returned masks are in sorted order, as
    ``thrust::sort_by_key`` leaves them (``all.py:935-1000`` sorts keys in place).
    """
    kv = pairs.shape[1]
    words = (kv + 31) // 32
    cnt = _pair_counts(num, kv, n_in, subm)
    pair_fwd = np.full((kv, n_out), -1, dtype=np.int32)
    pair_bwd = np.full((kv, n_in), -1, dtype=np.int32)
    mask_fwd = np.zeros((n_out, words), dtype=np.uint32)
    mask_bwd = np.zeros((n_in, words), dtype=np.uint32)
    for k in range(kv):
        i_inds = pairs[0, k, :cnt[k]]
        o_inds = pairs[1, k, :cnt[k]]
        pair_fwd[k, o_inds] = i_inds
        pair_bwd[k, i_inds] = o_inds
        bit = np.uint32(1 << (k % 32))
        mask_fwd[o_inds, k // 32] |= bit
        mask_bwd[i_inds, k // 32] |= bit

    def sort_masks(mask):
        n = mask.shape[0]
        if not do_sort:
            return mask.copy(), np.arange(n, dtype=np.int32)
        # thrust tuple compare: word 0 is the most significant key
        order = np.arange(n)
        for w in range(words - 1, -1, -1):
            order = order[np.argsort(mask[order, w], kind="stable")]
        return mask[order].copy(), order.astype(np.int32)

    mask_fwd_sorted, argsort_fwd = sort_masks(mask_fwd)
    mask_bwd_sorted, argsort_bwd = sort_masks(mask_bwd)
    return {
        "pair_fwd": pair_fwd, "pair_bwd": pair_bwd,
        "mask_fwd_unsorted": mask_fwd, "mask_bwd_unsorted": mask_bwd,
        "mask_fwd": mask_fwd_sorted, "mask_bwd": mask_bwd_sorted,
        "argsort_fwd": argsort_fwd, "argsort_bwd": argsort_bwd,
    }


def _mm(a: np.ndarray, b: np.ndarray, out: Optional[np.ndarray] = None) -> np.ndarray:
    """The reference's mm callback is ``torch.mm`` on the host BLAS (``cppcore.py:232-348``);
    transposed operands are passed as strided views (no copies), as ``tensor.T`` is there."""
    try:
        import torch
    except Exception:                                   # pragma: no cover
        return np.matmul(a, b, out=out)
    ta, tb = torch.from_numpy(a), torch.from_numpy(b)    # from_numpy keeps numpy's strides
    if out is None:
        return torch.mm(ta, tb).numpy()
    torch.mm(ta, tb, out=torch.from_numpy(out))
    return out


def gather_rows(buf: np.ndarray, src: np.ndarray, inds: np.ndarray) -> None:
    """``GatherCPU::gather`` (``gather.py:30-56``): buf[i] = src[inds[i]].  Runs the reference's
    own code when ``oracle/_ref`` is built, else its C restatement."""
    n, c = int(inds.shape[0]), int(src.shape[1])
    assert buf.dtype == np.float32 and src.dtype == np.float32 and inds.dtype == np.int32
    assert buf.flags.c_contiguous and src.flags.c_contiguous and inds.flags.c_contiguous
    r = ref_lib()
    if r is not None:
        r.ref_gather_f32(_ptr(buf), _ptr(src), _ptr(inds), n, c, int(src.shape[0]))
    else:
        _load().orc_gather_f32(_ptr(buf), _ptr(src), _ptr(inds), n, c)


def scatter_add_rows(dst: np.ndarray, buf: np.ndarray, inds: np.ndarray) -> None:
    """``GatherCPU::scatter_add`` (``gather.py:58-86``): dst[inds[i]] += buf[i]."""
    n, c = int(inds.shape[0]), int(dst.shape[1])
    assert dst.dtype == np.float32 and buf.dtype == np.float32 and inds.dtype == np.int32
    assert dst.flags.c_contiguous and buf.flags.c_contiguous and inds.flags.c_contiguous
    r = ref_lib()
    if r is not None:
        r.ref_scatter_add_f32(_ptr(dst), _ptr(buf), _ptr(inds), n, c, int(dst.shape[0]))
    else:
        _load().orc_scatter_add_f32(_ptr(dst), _ptr(buf), _ptr(inds), n, c)


def indice_conv(features: np.ndarray, filters: np.ndarray, pairs: np.ndarray, num: np.ndarray,
                num_activate_out: int, inverse: bool = False, subm: bool = False,
                bias: Optional[np.ndarray] = None, act: Optional[str] = None,
                act_alpha: float = 0.0) -> np.ndarray:
    """fp32 gather -> mm -> scatter-add forward (``convops.py:1534-1633``).

    ``filters`` is KRSC ``[K, *ksize, C]``; it is viewed ``[K, kv, C]`` and offset ``k`` uses
    ``W_k = filters[:, k, :]`` with ``out[pair_out] += x[pair_in] @ W_k^T``.
    """
    x = np.ascontiguousarray(features, dtype=np.float32)
    K, C = filters.shape[0], filters.shape[-1]
    w = np.ascontiguousarray(filters, dtype=np.float32).reshape(K, -1, C)
    kv = w.shape[1]
    cnt = _pair_counts(num, kv, x.shape[0], subm)
    if subm:
        out = np.ascontiguousarray(_mm(x, w[:, kv // 2].T))       # cppcore.py:244-246
    else:
        out = np.zeros((num_activate_out, K), dtype=np.float32)   # convops.py:1567-1568
    pin, pout = (pairs[1], pairs[0]) if inverse else (pairs[0], pairs[1])  # convops.py:1604-1605
    maxnhot = int(max([cnt[k] for k in range(kv) if not (subm and k == kv // 2)] + [0]))
    inp_buffer = np.empty((max(maxnhot, 1), C), dtype=np.float32)  # AllocKeys.InpBuffer, :1608
    out_buffer = np.empty((max(maxnhot, 1), K), dtype=np.float32)  # AllocKeys.OutBuffer, :1610
    for k in range(kv):                                            # convops.py:1612-1631
        if subm and k == kv // 2:
            continue
        n = int(cnt[k])
        if n <= 0:
            continue
        gather_rows(inp_buffer, x, np.ascontiguousarray(pin[k, :n]))
        _mm(inp_buffer[:n], w[:, k].T, out=out_buffer[:n])
        scatter_add_rows(out, out_buffer, np.ascontiguousarray(pout[k, :n]))
    if bias is not None:
        out = out + np.asarray(bias, dtype=np.float32)
    return apply_act(out, act, act_alpha)


def indice_conv_backward(features: np.ndarray, filters: np.ndarray, out_bp: np.ndarray,
                         pairs: np.ndarray, num: np.ndarray, inverse: bool = False,
                         subm: bool = False) -> Tuple[np.ndarray, np.ndarray]:
    """fp32 backward (``convops.py:1769-1860``): ``dW_k = dout[po]^T @ x[pi]``,
    ``din[pi] += dout[po] @ W_k``.  Returns ``(din [N,C], dfilters KRSC)``."""
    x = np.ascontiguousarray(features, dtype=np.float32)
    dout = np.ascontiguousarray(out_bp, dtype=np.float32)
    K, C = filters.shape[0], filters.shape[-1]
    w = np.ascontiguousarray(filters, dtype=np.float32).reshape(K, -1, C)
    kv = w.shape[1]
    cnt = _pair_counts(num, kv, x.shape[0], subm)
    dw = np.zeros_like(w)
    if subm:
        dw[:, kv // 2] = _mm(dout.T, x)               # cppcore.py:314-317
        din = np.ascontiguousarray(_mm(dout, w[:, kv // 2]))
    else:
        din = np.zeros_like(x)
    pin, pout = (pairs[1], pairs[0]) if inverse else (pairs[0], pairs[1])
    maxnhot = int(max([cnt[k] for k in range(kv) if not (subm and k == kv // 2)] + [0]))
    inp_buffer = np.empty((max(maxnhot, 1), C), dtype=np.float32)
    out_buffer = np.empty((max(maxnhot, 1), K), dtype=np.float32)
    for k in range(kv):                               # convops.py:1831-1860
        if subm and k == kv // 2:
            continue
        n = int(cnt[k])
        if n <= 0:
            continue
        gather_rows(inp_buffer, x, np.ascontiguousarray(pin[k, :n]))
        gather_rows(out_buffer, dout, np.ascontiguousarray(pout[k, :n]))
        dw[:, k] = _mm(out_buffer[:n].T, inp_buffer[:n])          # KN @ NC  (cppcore.py:341-343)
        _mm(out_buffer[:n], w[:, k], out=inp_buffer[:n])          # NK @ KC  (cppcore.py:347-348)
        scatter_add_rows(din, inp_buffer, np.ascontiguousarray(pin[k, :n]))
    return din, dw.reshape(filters.shape)


def apply_act(x: np.ndarray, act: Optional[str], alpha: float = 0.0) -> np.ndarray:
    """``inference.py:26-146`` activations: ReLU, LeakyReLU(alpha), Sigmoid."""
    if act is None or act == "none":
        return x
    if act == "relu":
        return np.maximum(x, 0)
    if act == "leaky_relu":
        return np.where(x >= 0, x, x * np.float32(alpha))
    if act == "sigmoid":
        return (1.0 / (1.0 + np.exp(-x))).astype(x.dtype)
    raise NotImplementedError(act)


def int8_conv_forward(features_i8: np.ndarray, filters_i8: np.ndarray, pairs, num,
                      num_activate_out: int, subm: bool, scales: np.ndarray, bias: np.ndarray,
                      output_add: Optional[np.ndarray] = None, output_add_scale: float = 0.0,
                      relu: bool = False, out_int8: bool = True) -> np.ndarray:
    """int8 inference formula of ``test/test_all_algo.py:222-287``: int32 accumulate, then
    ``clip(round(acc * scale[k] + bias[k] + add * add_scale))`` (numpy round-half-even)."""
    x = features_i8.astype(np.int32)
    K, C = filters_i8.shape[0], filters_i8.shape[-1]
    w = filters_i8.reshape(K, -1, C).astype(np.int32)
    kv = w.shape[1]
    cnt = _pair_counts(num, kv, x.shape[0], subm)
    acc = np.zeros((num_activate_out, K), dtype=np.int32)
    for k in range(kv):
        n = int(cnt[k])
        if n <= 0:
            continue
        # within one offset every output row occurs at most once (A.3 / A.4), so a fancy-indexed
        # "+=" is exact here (and ~50x faster than np.add.at at 100 k voxels)
        acc[pairs[1][k, :n]] += x[pairs[0][k, :n]] @ w[:, k].T
    res = acc.astype(np.float32) * scales.astype(np.float32) + bias.astype(np.float32)
    if output_add is not None:
        res = res + output_add.astype(np.float32) * np.float32(output_add_scale)
    if relu:
        res = np.maximum(res, 0)
    if out_int8:
        return np.clip(np.round(res), -128, 127).astype(np.int8)
    return res


def dense_from_sparse(features: np.ndarray, indices: np.ndarray, spatial_shape, batch_size):
    """NC(D..) dense tensor from a sparse one (``spconv/pytorch/core.py:264-275``)."""
    C = features.shape[1]
    dense = np.zeros((batch_size, *spatial_shape, C), dtype=features.dtype)
    dense[tuple(indices[:, i] for i in range(indices.shape[1]))] = features
    nd = len(spatial_shape)
    return np.ascontiguousarray(dense.transpose(0, nd + 1, *range(1, nd + 1)))


def generate_sparse_data(shape, num_points, num_channels, rng: np.random.Generator,
                         data_range=(-1, 1), dtype=np.float32):
    """Semantics of ``spconv/test_utils.py:142-195``: unique uniform-random coordinates per
    batch sample (shuffle of the full mesh grid), uniform features."""
    ndim = len(shape)
    total = int(np.prod(shape))
    inds = []
    for b, n in enumerate(num_points):
        flat = rng.permutation(total)[:n]
        coords = np.stack(np.unravel_index(flat, shape), axis=-1).astype(np.int32)
        inds.append(np.concatenate([np.full((n, 1), b, dtype=np.int32), coords], axis=1))
    indices = np.concatenate(inds, axis=0)
    feats = rng.uniform(data_range[0], data_range[1],
                        size=(indices.shape[0], num_channels)).astype(dtype)
    assert indices.shape[1] == ndim + 1
    return feats, indices


# ---------------------------------------------------------------------------- pooling
def indice_maxpool(features: np.ndarray, pairs: np.ndarray, num: np.ndarray, num_activate_out: int) -> np.ndarray:
    """ConvAlgo.Native max pooling: a ZERO-initialised output raised offset by offset
    (``spconv/pytorch/ops.py:1899-1936`` + ``IndiceMaxPoolCPU::forward``, ``maxpool.py:623-658``);
    runs the reference's own loop when ``oracle/_ref`` is built."""
    x = np.ascontiguousarray(features, dtype=np.float32)
    out = np.zeros((int(num_activate_out), x.shape[1]), dtype=np.float32)
    r = ref_lib()
    for k in range(pairs.shape[1]):
        n = int(num[k])
        if n <= 0:
            continue
        pi, po = np.ascontiguousarray(pairs[0, k, :n]), np.ascontiguousarray(pairs[1, k, :n])
        if r is not None:
            r.ref_maxpool_fwd_f32(_ptr(out), _ptr(x), _ptr(po), _ptr(pi), n, x.shape[1], out.shape[0], x.shape[0])
        else:
            np.maximum.at(out, po, x[pi])
    return out


def indice_maxpool_backward(features, out_features, out_bp, pairs, num) -> np.ndarray:
    """``din[i] += dout[o]`` where ``x[i] == y[o]`` (``ops.py:1939-1972``, ``maxpool.py:661-700``)."""
    x = np.ascontiguousarray(features, dtype=np.float32)
    y = np.ascontiguousarray(out_features, dtype=np.float32)
    dy = np.ascontiguousarray(out_bp, dtype=np.float32)
    din = np.zeros_like(x)
    r = ref_lib()
    for k in range(pairs.shape[1]):
        n = int(num[k])
        if n <= 0:
            continue
        pi, po = np.ascontiguousarray(pairs[0, k, :n]), np.ascontiguousarray(pairs[1, k, :n])
        if r is not None:
            r.ref_maxpool_bwd_f32(_ptr(y), _ptr(x), _ptr(dy), _ptr(din), _ptr(po), _ptr(pi), n, x.shape[1],
                                  y.shape[0], x.shape[0])
        else:
            np.add.at(din, pi, np.where(x[pi] == y[po], dy[po], 0))
    return din


def maxpool_implicit_gemm(features: np.ndarray, pair_fwd: np.ndarray, lowest: float) -> np.ndarray:
    """``forward_implicit_gemm_kernel`` (``maxpool.py:76-117``, CUDA only in the reference):
    max over the valid entries of ``pair_fwd[:, o]``, starting from the dtype's lowest value."""
    x = np.asarray(features, dtype=np.float32)
    kv, m = pair_fwd.shape
    out = np.full((m, x.shape[1]), lowest, dtype=np.float32)
    for k in range(kv):
        o = np.nonzero(pair_fwd[k] >= 0)[0]
        out[o] = np.maximum(out[o], x[pair_fwd[k, o]])
    return out


def maxpool_implicit_gemm_backward(features, out_features, out_bp, pair_bwd) -> np.ndarray:
    """``backward_implicit_gemm_kernel`` (``maxpool.py:159-208``)."""
    x, y, dy = (np.asarray(a, dtype=np.float32) for a in (features, out_features, out_bp))
    din = np.zeros_like(x)
    for k in range(pair_bwd.shape[0]):
        i = np.nonzero(pair_bwd[k] >= 0)[0]
        o = pair_bwd[k, i]
        din[i] += np.where(x[i] == y[o], dy[o], 0)
    return din


def avgpool_implicit_gemm(features: np.ndarray, pair_fwd: np.ndarray):
    """``forward_avgpool_implicit_gemm_kernel`` (``maxpool.py:211-259``): mean over the valid
    entries; returns ``(out, count)``."""
    x = np.asarray(features, dtype=np.float32)
    kv, m = pair_fwd.shape
    acc = np.zeros((m, x.shape[1]), dtype=np.float32)
    count = (pair_fwd >= 0).sum(axis=0).astype(np.int32)
    for k in range(kv):
        o = np.nonzero(pair_fwd[k] >= 0)[0]
        acc[o] += x[pair_fwd[k, o]]
    out = np.where(count[:, None] > 0, acc / np.maximum(count, 1)[:, None].astype(np.float32), 0).astype(np.float32)
    return out, count


def avgpool_implicit_gemm_backward(out_bp, pair_bwd, count) -> np.ndarray:
    """``backward_avgpool_implicit_gemm_kernel`` (``maxpool.py:262-300``): the reference MULTIPLIES the
    upstream gradient by the neighbour count; restated as is."""
    dy = np.asarray(out_bp, dtype=np.float32)
    din = np.zeros((pair_bwd.shape[1], dy.shape[1]), dtype=np.float32)
    for k in range(pair_bwd.shape[0]):
        i = np.nonzero(pair_bwd[k] >= 0)[0]
        o = pair_bwd[k, i]
        din[i] += dy[o] * count[o][:, None].astype(np.float32)
    return din


def global_pool_rearrange(coords: np.ndarray, batch_size: int):
    """``IndiceMaxPoolCPU::global_pool_rearrange`` (``maxpool.py:599-620``)."""
    coords = np.ascontiguousarray(coords, dtype=np.int32)
    n = coords.shape[0]
    out = np.zeros((batch_size, n), dtype=np.int32)
    counts = np.zeros((batch_size,), dtype=np.int32)
    r = ref_lib()
    if r is not None:
        r.ref_global_pool_rearrange(_ptr(out), _ptr(coords), _ptr(counts), n, coords.shape[1], batch_size)
    else:
        for i in range(n):
            b = coords[i, 0]
            if b >= 0:
                out[b, counts[b]] = i
                counts[b] += 1
    return out, counts


# ---------------------------------------------------------------------------- point -> voxel
def point2voxel_meta(vsize_xyz, coors_range_xyz):
    """``Point2VoxelCommon::calc_meta_data`` (``pointops.py:42-88``) through the reference's own code
    when ``oracle/_ref`` is built: zyx-ordered ``(vsize[3], grid[3], stride[3], range[6])``."""
    r = ref_lib()
    if r is None:
        raise RuntimeError("oracle/_ref is not built")
    vs = np.zeros(3, np.float32); grid = np.zeros(3, np.int32); stride = np.zeros(3, np.int64); rng = np.zeros(6, np.float32)
    a = np.ascontiguousarray(vsize_xyz, dtype=np.float32); b = np.ascontiguousarray(coors_range_xyz, dtype=np.float32)
    r.ref_point2voxel_meta_3d(_ptr(a), _ptr(b), _ptr(vs), _ptr(grid), _ptr(stride), _ptr(rng))
    return vs, grid, stride, rng


def point2voxel_ref(points: np.ndarray, vsize_xyz, coors_range_xyz, max_voxels: int, max_points: int,
                    empty_mean: bool = False):
    """The reference's CPU voxel generator (``Point2VoxelCPU::point_to_voxel_static``,
    ``pointops.py:589-695``) run through ``oracle/_ref``.  NOTE its ``empty_mean`` variant never resets
    the running mean between voxels (``mean_value.clear()`` on a sized-by-constructor vector,
    :676-680), so only ``empty_mean=False`` is a usable oracle."""
    r = ref_lib()
    if r is None:
        raise RuntimeError("oracle/_ref is not built")
    r.ref_point2voxel_3d.restype = ctypes.c_int
    pts = np.ascontiguousarray(points, dtype=np.float32)
    n, nf = pts.shape
    vs, grid, _, rng = point2voxel_meta(vsize_xyz, coors_range_xyz)
    voxels = np.zeros((max_voxels, max_points, nf), np.float32)
    indices = np.zeros((max_voxels, 3), np.int32)
    num = np.zeros((max_voxels,), np.int32)
    dense = np.full(tuple(int(g) for g in grid), -1, np.int32)
    ids = np.zeros((n,), np.int64)
    m = r.ref_point2voxel_3d(_ptr(pts), n, nf, _ptr(voxels), _ptr(indices), _ptr(num), _ptr(dense), _ptr(ids), _ptr(vs),
                             _ptr(grid), _ptr(rng), int(max_voxels), int(max_points), int(bool(empty_mean)), 1)
    return voxels[:m].copy(), indices[:m].copy(), num[:m].copy(), ids


def point2voxel(points: np.ndarray, vsize_xyz, coors_range_xyz, max_voxels: int, max_points: int,
                empty_mean: bool = False):
    """numpy restatement of the same algorithm (first-touch voxel ids, first ``max_points`` points per
    voxel in input order); ``empty_mean`` uses the proper per-voxel mean (what the reference's GPU
    kernel ``voxel_empty_fill_mean`` computes, ``pointops.py:252-281``)."""
    pts = np.ascontiguousarray(points, dtype=np.float32)
    n, nf = pts.shape
    nd = len(vsize_xyz)
    vs = np.array([vsize_xyz[nd - 1 - j] for j in range(nd)], np.float32)
    lo = np.array([coors_range_xyz[nd - 1 - j] for j in range(nd)], np.float32)
    hi = np.array([coors_range_xyz[2 * nd - 1 - j] for j in range(nd)], np.float32)
    grid = np.round((hi - lo) / vs).astype(np.int64)
    c = np.floor((pts[:, [nd - 1 - j for j in range(nd)]] - lo) / vs).astype(np.int64)
    ok = ((c >= 0) & (c < grid)).all(axis=1)
    key = np.zeros(n, np.int64)
    for j in range(nd):
        key = key * grid[j] + c[:, j]
    ids = np.full(n, -1, np.int64)
    voxels = np.zeros((max_voxels, max_points, nf), np.float32)
    indices = np.zeros((max_voxels, nd), np.int32)
    num = np.zeros(max_voxels, np.int32)
    seen = {}
    m = 0
    for i in range(n):
        if not ok[i]:
            continue
        v = seen.get(key[i])
        if v is None:
            if m >= max_voxels:          # voxel dropped: its points get no id (pointops.py:643-647)
                continue
            v = m
            m += 1
            seen[key[i]] = v
            indices[v] = c[i]
        ids[i] = v
        if num[v] < max_points:
            voxels[v, num[v]] = pts[i]
            num[v] += 1
    if empty_mean:
        for v in range(m):
            if 0 < num[v] < max_points:
                voxels[v, num[v]:] = voxels[v, :num[v]].mean(axis=0, dtype=np.float32)
    return voxels[:m], indices[:m], num[:m], ids
