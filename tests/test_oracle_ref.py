"""Pins the oracle to the REFERENCE ITSELF: ``oracle/_ref`` is the reference's own CPU rulebook and
gather/scatter C++ (``spconv/csrc/sparse/indices.py:77-269,1621-1778``, ``gather.py:30-86``),
extracted from /root/reference and compiled by ``oracle/make_ref.py``.  The C restatement
(``oracle/spconv_oracle.c``) that every GPU parity test compares against must agree with it
BIT FOR BIT -- pair order, first-touch output order, counts -- on config 1, on the rulebook
geometries of ``tests/test_rulebook_gpu.py`` and on the reference's LiDAR fixture coordinates.

Skipped only when neither /root/reference nor a prebuilt ``oracle/_ref`` library exists.
"""
import os

import numpy as np
import pytest

from tests.util import random_cloud, surface_cloud

GOLD = os.path.join(os.path.dirname(__file__), "golden")

CASES = [
    # (shape, pts per sample, ksize, stride, padding, dilation, subm, transpose) -- same list as
    # tests/test_rulebook_gpu.py (the GPU rulebooks are compared with the port on these)
    ([64, 64, 64], [5000], [3, 3, 3], [1, 1, 1], [1, 1, 1], [1, 1, 1], True, False),     # BASELINE configs[0]
    ([19, 18, 17], [1500, 1500], [3, 3, 3], [1, 1, 1], [0, 0, 0], [2, 2, 2], True, False),
    ([19, 18, 17], [1500, 1500], [3, 3, 3], [2, 2, 2], [1, 1, 1], [1, 1, 1], False, False),
    ([19, 18, 17], [1500], [2, 2, 2], [2, 2, 2], [0, 0, 0], [1, 1, 1], False, False),
    ([19, 18, 17], [1500], [3, 3, 3], [1, 1, 1], [0, 0, 0], [2, 2, 2], False, False),
    ([19, 18, 17], [1500], [3, 3, 3], [3, 3, 3], [2, 2, 2], [1, 1, 1], False, False),
    ([19, 18, 17], [700], [3, 3, 3], [2, 2, 2], [1, 1, 1], [1, 1, 1], False, True),
    ([40, 50], [900, 800], [3, 3], [1, 1], [1, 1], [1, 1], True, False),
    ([40, 50], [900], [3, 3], [2, 2], [1, 1], [1, 1], False, False),
    ([9, 10, 11, 12], [2000], [3, 3, 3, 3], [1, 1, 1, 1], [1] * 4, [1] * 4, True, False),
    ([30, 30, 30], [3000], [3, 1, 3], [1, 1, 1], [0, 0, 0], [1, 1, 1], True, False),
    ([200], [120], [5], [2], [2], [1], False, False),                                     # 1-D
]


@pytest.fixture(scope="module")
def ref(oracle):
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built and /root/reference not present")
    return oracle


def _both(orc, inds, bs, shape, k, s, p, d, subm, transpose):
    nd = len(shape)
    a = orc.get_indice_pairs(inds, bs, shape, k, s, p, d, [0] * nd, subm, transpose, impl="port")
    b = orc.get_indice_pairs(inds, bs, shape, k, s, p, d, [0] * nd, subm, transpose, impl="ref")
    return a, b


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"{'subm' if c[6] else 'conv'}{'T' if c[7] else ''}-{len(c[0])}d-k{c[2][0]}s{c[3][0]}p{c[4][0]}d{c[5][0]}")
@pytest.mark.parametrize("seed", [484, 50051])
def test_port_equals_reference_cpu_rulebook(case, seed, ref):
    shape, pts, ksize, stride, padding, dilation, subm, transpose = case
    rng = np.random.default_rng(seed)
    _, inds = random_cloud(rng, shape, pts, 1)
    (o_a, p_a, n_a), (o_b, p_b, n_b) = _both(ref, inds, len(pts), shape, ksize, stride, padding, dilation,
                                             subm, transpose)
    assert np.array_equal(n_a, n_b)
    assert np.array_equal(o_a, o_b)          # first-touch output order
    assert np.array_equal(p_a, p_b)          # pair ORDER, -1 padding included


def test_port_equals_reference_on_duplicates_and_out_of_range_batch(ref):
    """duplicate coordinates (first index wins, hash.insert) and rows whose batch index is outside
    [0, batch_size) -- the two input irregularities the reference code has explicit behaviour for"""
    rng = np.random.default_rng(9)
    _, inds = random_cloud(rng, [16, 16, 16], [900], 1)
    inds = np.concatenate([inds, inds[:50], inds[100:130]]).astype(np.int32)
    inds[7, 0] = 3
    inds[11, 0] = -1
    for subm in (True, False):
        s = [1] * 3 if subm else [2] * 3
        a, b = _both(ref, inds, 1, [16, 16, 16], [3] * 3, s, [1] * 3, [1] * 3, subm, False)
        for x, y in zip(a, b):
            assert np.array_equal(x, y)


def test_port_equals_reference_on_lidar_fixture(ref):
    """the coordinates of the reference's own test fixture (test/data/test_spconv.pkl): SubM 3^3 and
    the stride-2 downsample; totals are the committed fixture facts"""
    import json
    g = np.load(os.path.join(GOLD, "fixture_coords.npz"))
    facts = json.load(open(os.path.join(GOLD, "fixture_facts.json")))
    inds, shape = np.ascontiguousarray(g["coors"]), [int(v) for v in g["shape"]]
    a, b = _both(ref, inds, 1, shape, [3] * 3, [1] * 3, [1] * 3, [1] * 3, True, False)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    assert 2 * int(b[2].sum()) + inds.shape[0] == facts["subm_k3_pairs_total"]
    a, b = _both(ref, inds, 1, shape, [3] * 3, [2] * 3, [1] * 3, [1] * 3, False, False)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    assert int(b[2].sum()) == facts["conv_k3s2p1_pairs"] and b[0].shape[0] == facts["conv_k3s2p1_outputs"]


def test_port_equals_reference_on_kitti_shaped_surface_cloud(ref):
    rng = np.random.default_rng(50051)
    shape = [41, 1600, 1408]
    inds = surface_cloud(rng, shape, 30_000, batch=2)
    for subm, s in ((True, [1] * 3), (False, [2] * 3)):
        a, b = _both(ref, inds, 2, shape, [3] * 3, s, [1] * 3, [1] * 3, subm, False)
        for x, y in zip(a, b):
            assert np.array_equal(x, y)


def test_reference_gather_scatter_equal_the_restatement(ref):
    """GatherCPU::gather / scatter_add (gather.py:30-86) vs the C restatement and numpy"""
    rng = np.random.default_rng(1)
    src = rng.standard_normal((5000, 24)).astype(np.float32)
    inds = rng.permutation(5000)[:3000].astype(np.int32)
    lib, rlib = ref._load(), ref.ref_lib()
    buf_a = np.empty((3000, 24), np.float32)
    buf_b = np.empty_like(buf_a)
    lib.orc_gather_f32(ref._ptr(buf_a), ref._ptr(src), ref._ptr(inds), 3000, 24)
    rlib.ref_gather_f32(ref._ptr(buf_b), ref._ptr(src), ref._ptr(inds), 3000, 24, 5000)
    assert np.array_equal(buf_a, buf_b) and np.array_equal(buf_a, src[inds])
    dst_a = rng.standard_normal((5000, 24)).astype(np.float32)
    dst_b = dst_a.copy()
    want = dst_a.copy()
    np.add.at(want, inds, buf_a)
    lib.orc_scatter_add_f32(ref._ptr(dst_a), ref._ptr(buf_a), ref._ptr(inds), 3000, 24)
    rlib.ref_scatter_add_f32(ref._ptr(dst_b), ref._ptr(buf_a), ref._ptr(inds), 3000, 24, 5000)
    assert np.array_equal(dst_a, dst_b) and np.array_equal(dst_a, want)


def test_subm_even_ksize_error_matches(ref):
    _, inds = random_cloud(np.random.default_rng(0), [8, 8, 8], [50], 1)
    for impl in ("port", "ref"):
        with pytest.raises(RuntimeError, match="odd ksize"):
            ref.get_indice_pairs(inds, 1, [8, 8, 8], [2] * 3, [1] * 3, [0] * 3, [1] * 3, [0] * 3, True, impl=impl)


def test_pooling_oracle_against_reference_cpu_loop(ref):
    """IndiceMaxPoolCPU::forward / backward / global_pool_rearrange (maxpool.py:590-700, compiled into
    oracle/_ref) vs the numpy restatements the GPU pooling tests use"""
    rng = np.random.default_rng(0)
    feats, inds = random_cloud(rng, [18, 20, 22], [1200, 900], 16)
    o, pairs, num = ref.get_indice_pairs(inds, 2, [18, 20, 22], [3] * 3, [2] * 3, [1] * 3, [1] * 3, [0] * 3, False)
    got = ref.indice_maxpool(feats, pairs, num, o.shape[0])              # runs the reference's loop
    want = np.zeros_like(got)
    for k in range(27):
        np.maximum.at(want, pairs[1, k, :num[k]], feats[pairs[0, k, :num[k]]])
    assert np.array_equal(got, want)
    tabs = ref.implicit_gemm_tables(pairs, num, inds.shape[0], o.shape[0], False)
    dense = ref.maxpool_implicit_gemm(feats, tabs["pair_fwd"], -3e38)
    assert np.array_equal(np.maximum(dense, 0), got)                     # Native = zero floor
    g = rng.standard_normal(got.shape).astype(np.float32)
    d_ref = ref.indice_maxpool_backward(feats, dense, g, pairs, num)     # reference loop
    d_np = ref.maxpool_implicit_gemm_backward(feats, dense, g, tabs["pair_bwd"])
    assert np.abs(d_ref - d_np).max() < 1e-6
    oi, cnt = ref.global_pool_rearrange(inds, 2)
    assert cnt.tolist() == [1200, 900] and np.array_equal(oi[1, :900], np.arange(1200, 2100))


def test_point2voxel_restatement_against_reference_cpu_generator(ref):
    """Point2VoxelCPU::point_to_voxel_static (pointops.py:589-695, compiled into oracle/_ref) vs the numpy
    restatement, incl. the max_num_voxels cap and out-of-range points"""
    rng = np.random.default_rng(0)
    pts = rng.uniform([-1, -41, -4, 0], [71, 41, 2, 1], size=(20000, 4)).astype(np.float32)
    vs, cr = [0.4, 0.4, 0.5], [0, -40, -3, 70.4, 40, 1]
    for max_voxels, max_points in ((3000, 5), (50000, 2)):
        a = ref.point2voxel_ref(pts, vs, cr, max_voxels, max_points)
        b = ref.point2voxel(pts, vs, cr, max_voxels, max_points)
        for x, y in zip(a, b):
            assert np.array_equal(x, y)
    _, grid, stride, rng6 = ref.point2voxel_meta(vs, cr)
    assert grid.tolist() == [8, 200, 176] and stride.tolist() == [35200, 176, 1]
