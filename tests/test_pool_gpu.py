"""Sparse max / average pooling against the oracle: the ConvAlgo.Native pool runs the
reference's own CPU loop (``IndiceMaxPoolCPU``, extracted into oracle/_ref); the implicit-GEMM
variants are compared with numpy restatements of the reference CUDA kernels
(``spconv/csrc/sparse/maxpool.py:76-300``) and with torch dense pooling."""
import numpy as np
import pytest
import torch

from tests.util import random_cloud, rel_l2

pytestmark = pytest.mark.gpu


def _cloud(seed, shape, pts, c, positive=False):
    rng = np.random.default_rng(seed)
    feats, inds = random_cloud(rng, shape, pts, c)
    if positive:
        feats = np.abs(feats) + 0.1
    return feats, inds


@pytest.mark.parametrize("k,s,p", [(3, 2, 1), (2, 2, 0), (3, 1, 1)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
def test_maxpool_implicit_gemm_fwd_bwd(k, s, p, dtype, oracle, cuda_dev):
    import spconv_b200.pytorch as spconv
    shape = [18, 20, 22]
    feats, inds = _cloud(3, shape, [1400, 1300], 32)
    feats = torch.from_numpy(feats).to(dtype).float().numpy()          # exactly representable
    pool = spconv.SparseMaxPool3d(k, s, p, indice_key="p")
    xf = torch.from_numpy(feats).to(cuda_dev).to(dtype).requires_grad_(True)
    x = spconv.SparseConvTensor(xf, torch.from_numpy(inds).to(cuda_dev), shape, 2)
    y = pool(x)
    o, pairs, num = oracle.get_indice_pairs(inds, 2, shape, [k] * 3, [s] * 3, [p] * 3, [1] * 3, [0] * 3, False)
    assert np.array_equal(y.indices.cpu().numpy(), o)
    tabs = oracle.implicit_gemm_tables(pairs, num, inds.shape[0], o.shape[0], False)
    lowest = float(torch.finfo(dtype).min)
    ref = oracle.maxpool_implicit_gemm(feats, tabs["pair_fwd"], lowest)
    assert np.array_equal(y.features.detach().float().cpu().numpy(), ref)          # max is exact
    g = np.random.default_rng(1).uniform(-1, 1, ref.shape).astype(np.float32)
    g = torch.from_numpy(g).to(dtype)
    y.features.backward(g.to(cuda_dev))
    ref_din = oracle.maxpool_implicit_gemm_backward(feats, ref, g.float().numpy(), tabs["pair_bwd"])
    tol = 1e-6 if dtype == torch.float32 else 2e-3
    assert np.abs(xf.grad.float().cpu().numpy() - ref_din).max() <= tol * max(1.0, np.abs(ref_din).max())
    # the pooled tensor keeps the rulebook for a paired inverse conv
    assert "p" in y.indice_dict and y.spatial_shape == oracle.get_conv_output_size(shape, [k] * 3, [s] * 3, [p] * 3, [1] * 3)


def test_maxpool_equals_dense_maxpool_on_positive_features(cuda_dev):
    """with positive features (empty cells = 0 never win) sparse max pooling == dense max_pool3d"""
    import spconv_b200.pytorch as spconv
    shape = [16, 16, 16]
    feats, inds = _cloud(5, shape, [1500], 16, positive=True)
    x = spconv.SparseConvTensor(torch.from_numpy(feats).to(cuda_dev), torch.from_numpy(inds).to(cuda_dev), shape, 1)
    y = spconv.SparseMaxPool3d(2, 2)(x)
    dense = torch.nn.functional.max_pool3d(x.dense(), 2, 2)
    assert torch.equal(y.dense(), dense)


@pytest.mark.parametrize("k,s,p", [(3, 2, 1), (2, 2, 0)])
def test_maxpool_native_matches_reference_cpu_loop(k, s, p, oracle, cuda_dev):
    """ConvAlgo.Native: zero-initialised output raised per offset (negative maxima clamp to 0, as in
    the reference) -- checked against IndiceMaxPoolCPU::forward / backward"""
    import spconv_b200.pytorch as spconv
    from spconv_b200.core import ConvAlgo
    shape = [18, 20, 22]
    feats, inds = _cloud(7, shape, [1200, 900], 16)
    pool = spconv.SparseMaxPool3d(k, s, p, algo=ConvAlgo.Native)
    xf = torch.from_numpy(feats).to(cuda_dev).requires_grad_(True)
    y = pool(spconv.SparseConvTensor(xf, torch.from_numpy(inds).to(cuda_dev), shape, 2))
    o, pairs, num = oracle.get_indice_pairs(inds, 2, shape, [k] * 3, [s] * 3, [p] * 3, [1] * 3, [0] * 3, False)
    ref = oracle.indice_maxpool(feats, pairs, num, o.shape[0])
    assert np.array_equal(y.indices.cpu().numpy(), o)
    assert np.array_equal(y.features.detach().cpu().numpy(), ref)
    assert (ref == 0).any() and (ref > 0).any()
    g = np.random.default_rng(2).uniform(-1, 1, ref.shape).astype(np.float32)
    y.features.backward(torch.from_numpy(g).to(cuda_dev))
    ref_din = oracle.indice_maxpool_backward(feats, ref, g, pairs, num)
    assert np.abs(xf.grad.cpu().numpy() - ref_din).max() < 1e-5


@pytest.mark.parametrize("train", [True, False])
def test_avgpool_fwd_bwd(train, oracle, cuda_dev):
    import spconv_b200.pytorch as spconv
    shape = [20, 24]
    rng = np.random.default_rng(11)
    feats, inds = random_cloud(rng, shape, [260, 240], 8)
    pool = spconv.SparseAvgPool2d(3, 2, 1)
    pool.train(train)
    xf = torch.from_numpy(feats).to(cuda_dev).requires_grad_(True)
    y = pool(spconv.SparseConvTensor(xf, torch.from_numpy(inds).to(cuda_dev), shape, 2))
    o, pairs, num = oracle.get_indice_pairs(inds, 2, shape, [3] * 2, [2] * 2, [1] * 2, [1] * 2, [0] * 2, False)
    tabs = oracle.implicit_gemm_tables(pairs, num, inds.shape[0], o.shape[0], False)
    ref, count = oracle.avgpool_implicit_gemm(feats, tabs["pair_fwd"])
    assert np.array_equal(y.indices.cpu().numpy(), o)
    assert np.abs(y.features.detach().cpu().numpy() - ref).max() < 1e-6
    if train:
        g = rng.uniform(-1, 1, ref.shape).astype(np.float32)
        y.features.backward(torch.from_numpy(g).to(cuda_dev))
        ref_din = oracle.avgpool_implicit_gemm_backward(g, tabs["pair_bwd"], count)
        assert rel_l2(xf.grad.cpu().numpy(), ref_din) < 1e-6


def test_global_pools(oracle, cuda_dev):
    import spconv_b200.pytorch as spconv
    from spconv_b200.pytorch import ops
    shape = [12, 12, 12]
    feats, inds = _cloud(13, shape, [300, 500, 200], 8)
    perm = np.random.default_rng(0).permutation(inds.shape[0])          # samples interleaved
    feats, inds = feats[perm], inds[perm]
    d_inds = torch.from_numpy(inds).to(cuda_dev)
    oi, cnt = ops.global_pool_rearrange(d_inds, 3)
    r_oi, r_cnt = oracle.global_pool_rearrange(inds, 3)
    assert np.array_equal(cnt.cpu().numpy(), r_cnt)
    for b in range(3):
        assert np.array_equal(oi[b, :r_cnt[b]].cpu().numpy(), r_oi[b, :r_cnt[b]])
    x = spconv.SparseConvTensor(torch.from_numpy(feats).to(cuda_dev), d_inds, shape, 3)
    mx, av = spconv.SparseGlobalMaxPool()(x), spconv.SparseGlobalAvgPool()(x)
    for b in range(3):
        sel = feats[inds[:, 0] == b]
        assert np.allclose(mx[b].cpu().numpy(), sel.max(0)) and np.allclose(av[b].cpu().numpy(), sel.mean(0), atol=1e-6)
