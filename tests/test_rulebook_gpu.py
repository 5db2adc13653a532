"""GPU rulebook parity: bit-exact against the oracle (reference CPU order), through the C ABI.

Reference behaviour being pinned: ``SparseConvIndicesCPU`` (spconv/csrc/sparse/indices.py:1640-1778)
for the Native rulebook; SURVEY A.5 for the implicit-GEMM tables.
"""
import numpy as np
import pytest
import torch

from tests.util import random_cloud, surface_cloud

pytestmark = pytest.mark.gpu

CASES = [
    # (shape, pts per sample, ksize, stride, padding, dilation, subm, transpose)
    ([64, 64, 64], [5000], [3, 3, 3], [1, 1, 1], [1, 1, 1], [1, 1, 1], True, False),     # cfg1
    ([19, 18, 17], [1500, 1500], [3, 3, 3], [1, 1, 1], [0, 0, 0], [2, 2, 2], True, False),
    ([19, 18, 17], [1500, 1500], [3, 3, 3], [2, 2, 2], [1, 1, 1], [1, 1, 1], False, False),
    ([19, 18, 17], [1500], [2, 2, 2], [2, 2, 2], [0, 0, 0], [1, 1, 1], False, False),
    ([19, 18, 17], [1500], [3, 3, 3], [1, 1, 1], [0, 0, 0], [2, 2, 2], False, False),
    ([19, 18, 17], [1500], [3, 3, 3], [3, 3, 3], [2, 2, 2], [1, 1, 1], False, False),
    ([19, 18, 17], [700], [3, 3, 3], [2, 2, 2], [1, 1, 1], [1, 1, 1], False, True),
    ([40, 50], [900, 800], [3, 3], [1, 1], [1, 1], [1, 1], True, False),                   # 2-D
    ([40, 50], [900], [3, 3], [2, 2], [1, 1], [1, 1], False, False),
    ([9, 10, 11, 12], [2000], [3, 3, 3, 3], [1, 1, 1, 1], [1] * 4, [1] * 4, True, False),  # 4-D, kv=81
    ([30, 30, 30], [3000], [3, 1, 3], [1, 1, 1], [0, 0, 0], [1, 1, 1], True, False),
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"{'subm' if c[6] else 'conv'}{'T' if c[7] else ''}-{len(c[0])}d-k{c[2][0]}s{c[3][0]}p{c[4][0]}d{c[5][0]}")
def test_native_rulebook_bit_exact(case, oracle, cuda_dev):
    from spconv_b200.core import ConvAlgo
    from spconv_b200.pytorch import ops
    shape, pts, ksize, stride, padding, dilation, subm, transpose = case
    rng = np.random.default_rng(484)
    _, inds = random_cloud(rng, shape, pts, 1)
    bs = len(pts)
    ndim = len(shape)
    ref_out, ref_pairs, ref_num = oracle.get_indice_pairs(inds, bs, shape, ksize, stride, padding,
                                                          dilation, [0] * ndim, subm, transpose)
    out, pairs, num = ops.get_indice_pairs(torch.from_numpy(inds).to(cuda_dev), bs, shape,
                                           ConvAlgo.Native, ksize, stride, padding, dilation,
                                           [0] * ndim, subm, transpose)
    torch.cuda.synchronize()
    assert np.array_equal(num.cpu().numpy(), ref_num), (num.cpu().numpy(), ref_num)
    assert np.array_equal(out.cpu().numpy(), ref_out)
    assert np.array_equal(pairs.cpu().numpy(), ref_pairs)


@pytest.mark.parametrize("case", CASES[:10], ids=lambda c: f"{'subm' if c[6] else 'conv'}{'T' if c[7] else ''}-{len(c[0])}d-k{c[2][0]}s{c[3][0]}p{c[4][0]}d{c[5][0]}")
@pytest.mark.parametrize("do_sort", [True, False])
def test_implicit_gemm_rulebook_bit_exact(case, do_sort, oracle, cuda_dev):
    from spconv_b200.core import ConvAlgo
    from spconv_b200.pytorch import ops
    shape, pts, ksize, stride, padding, dilation, subm, transpose = case
    rng = np.random.default_rng(50051)
    _, inds = random_cloud(rng, shape, pts, 1)
    bs, ndim = len(pts), len(shape)
    ref_out, ref_pairs, ref_num = oracle.get_indice_pairs(inds, bs, shape, ksize, stride, padding,
                                                          dilation, [0] * ndim, subm, transpose)
    tab = oracle.implicit_gemm_tables(ref_pairs, ref_num, inds.shape[0], ref_out.shape[0], subm,
                                      do_sort)
    res = ops.get_indice_pairs_implicit_gemm(torch.from_numpy(inds).to(cuda_dev), bs, shape,
                                             ConvAlgo.MaskImplicitGemm, ksize, stride, padding,
                                             dilation, [0] * ndim, subm, transpose, is_train=True,
                                             do_sort=do_sort)
    torch.cuda.synchronize()
    out_inds, _, pair_fwd, pair_bwd, mask_fwd, mask_bwd, sort_fwd, sort_bwd, masks = res
    assert np.array_equal(out_inds.cpu().numpy(), ref_out)
    assert np.array_equal(pair_fwd.cpu().numpy(), tab["pair_fwd"])
    assert np.array_equal(pair_bwd.cpu().numpy(), tab["pair_bwd"])
    # masks come back SORTED (thrust::sort_by_key sorts keys in place, all.py:935-1000)
    assert np.array_equal(mask_fwd[0].cpu().numpy().view(np.uint32), tab["mask_fwd"])
    assert np.array_equal(sort_fwd[0].cpu().numpy(), tab["argsort_fwd"])
    if not subm:
        assert np.array_equal(mask_bwd[0].cpu().numpy().view(np.uint32), tab["mask_bwd"])
        assert np.array_equal(sort_bwd[0].cpu().numpy(), tab["argsort_bwd"])
    assert masks[0][0] == 0xffffffff


def test_subm_inference_tables(oracle, cuda_dev):
    """is_train=False: SubM returns pair [1, kv, N] only (ops.py:469-479)."""
    from spconv_b200.core import ConvAlgo
    from spconv_b200.pytorch import ops
    rng = np.random.default_rng(1)
    _, inds = random_cloud(rng, [20, 20, 20], [2000], 1)
    res = ops.get_indice_pairs_implicit_gemm(torch.from_numpy(inds).to(cuda_dev), 1, [20, 20, 20],
                                             ConvAlgo.MaskImplicitGemm, [3] * 3, [1] * 3, [1] * 3,
                                             [1] * 3, [0] * 3, True, False, is_train=False)
    assert res[3].numel() == 0
    ref_out, ref_pairs, ref_num = oracle.get_indice_pairs(inds, 1, [20] * 3, [3] * 3, [1] * 3,
                                                          [1] * 3, [1] * 3, [0] * 3, True)
    tab = oracle.implicit_gemm_tables(ref_pairs, ref_num, 2000, 2000, True)
    assert np.array_equal(res[2].cpu().numpy(), tab["pair_fwd"])


def test_large_clustered_cloud_properties(oracle, cuda_dev):
    """KITTI-shaped grid at BASELINE size: size-independent properties + oracle equality
    (the C oracle handles 100k voxels in well under a second)."""
    from spconv_b200.core import ConvAlgo
    from spconv_b200.pytorch import ops
    rng = np.random.default_rng(50051)
    shape = [41, 1600, 1408]
    inds = surface_cloud(rng, shape, 100_000)
    n = inds.shape[0]
    dev_inds = torch.from_numpy(inds).to(cuda_dev)
    res = ops.get_indice_pairs_implicit_gemm(dev_inds, 1, shape, ConvAlgo.MaskImplicitGemm,
                                             [3] * 3, [1] * 3, [1] * 3, [1] * 3, [0] * 3, True)
    pair_fwd = res[2].cpu().numpy()
    pair_bwd = res[3].cpu().numpy()
    mask = res[4][0].cpu().numpy().view(np.uint32)[:, 0]
    argsort = res[6][0].cpu().numpy()
    kv = 27
    # symmetry of SubM pairs: pair_bwd[k] == pair_fwd[kv-1-k]
    assert np.array_equal(pair_bwd, pair_fwd[::-1])
    # centre is the identity
    assert np.array_equal(pair_fwd[kv // 2], np.arange(n))
    # involution: j = pair_fwd[k][o] >= 0  =>  pair_fwd[kv-1-k][j] == o
    for k in range(kv):
        o = np.nonzero(pair_fwd[k] >= 0)[0]
        assert np.array_equal(pair_fwd[kv - 1 - k][pair_fwd[k][o]], o)
    # masks sorted ascending, argsort a permutation, mask bits == table occupancy
    assert np.all(np.diff(mask.astype(np.int64)) >= 0)
    assert np.array_equal(np.sort(argsort), np.arange(n))
    occ = np.zeros(n, np.uint32)
    for k in range(kv):
        occ |= (pair_fwd[k] >= 0).astype(np.uint32) << np.uint32(k)
    assert np.array_equal(mask, occ[argsort])
    # and the full thing equals the oracle
    ref_out, ref_pairs, ref_num = oracle.get_indice_pairs(inds, 1, shape, [3] * 3, [1] * 3,
                                                          [1] * 3, [1] * 3, [0] * 3, True)
    tab = oracle.implicit_gemm_tables(ref_pairs, ref_num, n, n, True)
    assert np.array_equal(pair_fwd, tab["pair_fwd"])
    assert np.array_equal(argsort, tab["argsort_fwd"])
    # native compact pairs too
    out, pairs, num = ops.get_indice_pairs(dev_inds, 1, shape, ConvAlgo.Native, [3] * 3, [1] * 3,
                                           [1] * 3, [1] * 3, [0] * 3, True)
    assert np.array_equal(num.cpu().numpy(), ref_num)
    assert np.array_equal(pairs.cpu().numpy(), ref_pairs)
    pairs_per_voxel = (2 * int(ref_num.sum()) + n) / n
    assert 3.0 < pairs_per_voxel < 12.0, pairs_per_voxel   # clustered like LiDAR (fixture: 6.28)


def test_strided_large_and_int64_keys(oracle, cuda_dev):
    """nuScenes-like stride-2 rulebook + the int64-key path (volume >= 2^31)."""
    from spconv_b200.core import ConvAlgo
    from spconv_b200.pytorch import ops
    rng = np.random.default_rng(7)
    shape = [41, 1440, 1440]
    inds = surface_cloud(rng, shape, 60_000)
    k, s, p, d = [3] * 3, [2] * 3, [1] * 3, [1] * 3
    ref_out, ref_pairs, ref_num = oracle.get_indice_pairs(inds, 1, shape, k, s, p, d, [0] * 3, False)
    out, pairs, num = ops.get_indice_pairs(torch.from_numpy(inds).to(cuda_dev), 1, shape,
                                           ConvAlgo.Native, k, s, p, d, [0] * 3, False)
    assert np.array_equal(out.cpu().numpy(), ref_out)
    assert np.array_equal(num.cpu().numpy(), ref_num)
    assert np.array_equal(pairs.cpu().numpy(), ref_pairs)
    # huge virtual grid -> int64 linear keys (ops.py:188-190)
    big = [2000, 2000, 2000]
    _, inds2 = random_cloud(rng, [50, 50, 50], [4000], 1)
    inds2[:, 1:] += 1900
    ref = oracle.get_indice_pairs(inds2, 1, big, [3] * 3, [1] * 3, [1] * 3, [1] * 3, [0] * 3, True)
    got = ops.get_indice_pairs(torch.from_numpy(inds2).to(cuda_dev), 1, big, ConvAlgo.Native,
                               [3] * 3, [1] * 3, [1] * 3, [1] * 3, [0] * 3, True)
    assert np.array_equal(got[2].cpu().numpy(), ref[2])
    assert np.array_equal(got[1].cpu().numpy(), ref[1])
    ref = oracle.get_indice_pairs(inds2, 1, big, k, s, p, d, [0] * 3, False)
    got = ops.get_indice_pairs(torch.from_numpy(inds2).to(cuda_dev), 1, big, ConvAlgo.Native,
                               k, s, p, d, [0] * 3, False)
    assert np.array_equal(got[0].cpu().numpy(), ref[0])
    assert np.array_equal(got[1].cpu().numpy(), ref[1])


def test_vanished_points_error(cuda_dev):
    from spconv_b200.core import ConvAlgo
    from spconv_b200.pytorch import ops
    # k=2, s=3, p=0 covers input coordinates {0,1,3,4,6,7}: a point at 2 has no output
    inds = torch.tensor([[0, 2, 2, 2]], dtype=torch.int32, device=cuda_dev)
    with pytest.raises(ValueError, match="vanished"):
        ops.get_indice_pairs(inds, 1, [8, 8, 8], ConvAlgo.Native, [2] * 3, [3] * 3, [0] * 3,
                             [1] * 3, [0] * 3, False)
    with pytest.raises(RuntimeError, match="odd ksize"):
        ops.get_indice_pairs(inds, 1, [8, 8, 8], ConvAlgo.Native, [2] * 3, [1] * 3, [0] * 3,
                             [1] * 3, [0] * 3, True)


def test_tile_table_from_row_table_matches_column_path(cuda_dev):
    """spx_build_tile_table has two sources for the same table: the column-major pair table and
    the row-major by-product of the 3x3x3 probe kernel.  Both must give identical blocks and
    tile masks (== per-tile OR of the sorted masks, the reference's mask_output_fwd with
    mask_width 128, spconv/csrc/sparse/convops.py:2180-2189)."""
    from spconv_b200.core import ConvAlgo
    from spconv_b200.pytorch import ops
    rng = np.random.default_rng(5)
    shape = [24, 200, 176]
    inds = torch.from_numpy(surface_cloud(rng, shape, 9000 + 77)).to(cuda_dev)
    n = inds.shape[0]
    # the fused native call (default) builds the table from the row-major by-product inside C; its cached
    # result must equal what the separate calls produce from either source
    fused = ops.get_indice_pairs_implicit_gemm(inds, 1, shape, ConvAlgo.MaskImplicitGemm, [3] * 3, [1] * 3,
                                               [1] * 3, [1] * 3, [0] * 3, True, False, is_train=True)
    fused_cache = getattr(fused[6][0], "_spx_tile_cache", None)
    assert fused_cache is not None, "the fused rulebook call must leave the tile table cached on the argsort"
    old = ops.SPCONV_B200_FUSED_RULEBOOK
    ops.SPCONV_B200_FUSED_RULEBOOK = False
    try:
        res = ops.get_indice_pairs_implicit_gemm(inds, 1, shape, ConvAlgo.MaskImplicitGemm, [3] * 3, [1] * 3,
                                                 [1] * 3, [1] * 3, [0] * 3, True, False, is_train=True)
    finally:
        ops.SPCONV_B200_FUSED_RULEBOOK = old
    for a, b in ((fused[2], res[2]), (fused[3], res[3]), (fused[4][0], res[4][0]), (fused[6][0], res[6][0])):
        assert torch.equal(a, b)
    pair_fwd, mask, argsort = res[2], res[4][0], res[6][0]
    hint = getattr(argsort, "_spx_row_table", None)
    assert hint is not None, "3x3x3 SubM rulebook must hand the row-major table to the tile builder"
    rows = hint[1].cpu().numpy()
    pf = pair_fwd.cpu().numpy()
    assert np.array_equal(rows[:, :27], pf.T) and (rows[:, 27:] == -1).all()
    t_rows, m_rows = ops._tile_tables(pair_fwd, mask, argsort, n, 27, owner=argsort)
    assert argsort._spx_row_table is None                       # consumed
    t_cols, m_cols = ops._tile_tables(pair_fwd, mask, argsort, n, 27, owner=None)
    assert torch.equal(t_rows, t_cols) and torch.equal(m_rows, m_cols)
    assert torch.equal(fused_cache[1], t_cols) and torch.equal(fused_cache[2], m_cols)
    tiles = (n + 127) // 128
    flat = t_cols.cpu().numpy()
    tab = flat[:tiles * 28 * 128].reshape(tiles, 28, 128)
    order = argsort.cpu().numpy()
    padded = np.full(tiles * 128, -1, np.int64)
    padded[:n] = order
    assert np.array_equal(tab[:, 27, :].reshape(-1), padded)
    want = np.where(padded[None, :] >= 0, pf[:, np.maximum(padded, 0)], -1)            # [27, tiles*128]
    assert np.array_equal(tab[:, :27, :].transpose(1, 0, 2).reshape(27, -1), want)
    sm = np.zeros(tiles * 128, np.uint32)
    sm[:n] = mask.cpu().numpy().reshape(-1).view(np.uint32)
    assert np.array_equal(m_cols.cpu().numpy().reshape(-1).view(np.uint32),
                          np.bitwise_or.reduce(sm.reshape(tiles, 128), axis=1))
    # schedule records behind the blocks: every tile once, heaviest (most offsets) first, ties in
    # ascending tile order; scheduler scratch zeroed
    rec = flat[tiles * 28 * 128: tiles * 28 * 128 + tiles * 8].reshape(tiles, 8)
    tmask = m_cols.cpu().numpy().reshape(-1).view(np.uint32)
    cost = np.array([bin(int(v)).count("1") for v in np.where(tmask == 0, 1, tmask)])
    want_order = np.argsort(-cost, kind="stable")
    assert np.array_equal(rec[:, 0], want_order)
    assert np.array_equal(rec[:, 1].view(np.uint32), np.where(tmask == 0, 1, tmask)[want_order])
    assert (rec[:, 2:] == 0).all()
    assert (flat[tiles * 28 * 128 + tiles * 8:] == 0).all() and flat.shape[0] == tiles * 28 * 128 + tiles * 8 + 64
