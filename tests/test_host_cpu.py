"""Host-side logic of the drop-in surface (no GPU): tensor container, module construction,
algo defaults, cache / error behaviour, containers.  Mirrors what a spconv.pytorch user relies on
(SURVEY Appendix B)."""
import numpy as np
import pytest
import torch

import spconv_b200.pytorch as spconv
from spconv_b200.core import Activation, ConvAlgo


def _tensor(n=6, c=4, shape=(8, 8, 8)):
    rng = np.random.default_rng(0)
    flat = rng.permutation(int(np.prod(shape)))[:n]
    coords = np.stack(np.unravel_index(flat, shape), -1).astype(np.int32)
    inds = np.concatenate([np.zeros((n, 1), np.int32), coords], 1)
    return spconv.SparseConvTensor(torch.randn(n, c), torch.from_numpy(inds), list(shape), 1)


def test_sparse_conv_tensor_contract():
    x = _tensor()
    with pytest.raises(ValueError, match="replace_feature"):
        x.features = torch.zeros(1)
    with pytest.raises(AssertionError):
        spconv.SparseConvTensor(torch.randn(3, 4), torch.zeros((3, 4), dtype=torch.int64), [8] * 3, 1)
    with pytest.raises(AssertionError):
        spconv.SparseConvTensor(torch.randn(3, 4), torch.zeros((3, 4), dtype=torch.int32), [8] * 2, 1)
    y = x.replace_feature(x.features * 2)
    assert y.indice_dict is x.indice_dict and y.indices is x.indices
    assert torch.equal((x + y).features, x.features * 3)
    assert torch.equal((x.features + x).features, x.features * 2)
    z = x.shadow_copy()
    assert z.features is x.features and z.indice_dict is x.indice_dict
    x.indice_dict["k"] = 1
    assert x.find_indice_pair("k") == 1 and x.find_indice_pair(None) is None
    sel = x.select_by_index(torch.tensor([0, 2]))
    assert sel.features.shape[0] == 2 and sel.indice_dict == {}
    assert x.spatial_size == 512


def test_dense_roundtrip():
    x = _tensor(n=20, c=3, shape=(5, 6, 7))
    d = x.dense()
    assert d.shape == (1, 3, 5, 6, 7)
    i = x.indices.long()
    assert torch.equal(d[i[:, 0], :, i[:, 1], i[:, 2], i[:, 3]], x.features)
    assert x.dense(channels_first=False).shape == (1, 5, 6, 7, 3)
    back = spconv.SparseConvTensor.from_dense(x.dense(channels_first=False))
    assert back.features.shape[0] == 20 and back.spatial_shape == [5, 6, 7]


def test_module_construction_and_defaults():
    m = spconv.SubMConv3d(16, 32, 3, indice_key="a")
    assert tuple(m.weight.shape) == (32, 3, 3, 3, 16)            # KRSC
    assert m.algo == ConvAlgo.MaskImplicitGemm and m.subm and m.bias.shape == (32,)
    assert spconv.SparseConv3d(4, 4, 5).algo == ConvAlgo.Native   # kv=125 > 32
    assert spconv.SparseConv3d(4, 4, 5, large_kernel_fast_algo=True).algo == ConvAlgo.MaskImplicitGemm
    assert spconv.SparseConv3d(4, 4, 3, algo=ConvAlgo.Native).algo == ConvAlgo.Native
    bound = np.sqrt(6.0 / ((1 + 5.0) * 16 * 27))
    assert float(m.weight.abs().max()) <= bound + 1e-7
    assert float(m.bias.abs().max()) <= 1 / np.sqrt(16 * 27) + 1e-7
    with pytest.raises(AssertionError, match="groups"):
        spconv.SubMConv3d(4, 4, 3, groups=2)
    inv = spconv.SparseInverseConv3d(8, 4, 3, indice_key="d")
    assert inv.inverse and not inv.subm
    assert "kernel_size=[3, 3, 3]" in repr(m)
    sd = m.state_dict()
    assert set(sd) == {"weight", "bias"}
    assert spconv.SparseConv2d(3, 5, 3).ndim == 2 and spconv.SubMConv1d(3, 5, 3).ndim == 1


def test_conv1x1_is_a_matmul_on_cpu():
    x = _tensor(n=10, c=4)
    m = spconv.SubMConv3d(4, 6, 1)
    y = m(x)
    ref = x.features @ m.weight.view(6, 4).t() + m.bias
    assert torch.allclose(y.features, ref, atol=1e-6)


def test_no_cpu_fallback():
    x = _tensor(n=10, c=4)
    for algo in (ConvAlgo.Native, ConvAlgo.MaskImplicitGemm):
        with pytest.raises(RuntimeError, match="no CPU path"):
            spconv.SubMConv3d(4, 4, 3, algo=algo)(x)


def test_training_rejects_fused_activation():
    x = _tensor(n=10, c=4)
    m = spconv.SubMConv3d(4, 4, 3, act_type=Activation.ReLU)
    m.train()
    with pytest.raises(AssertionError, match="act don't support backward"):
        m(x)


def test_sequential_container():
    seq = spconv.SparseSequential(spconv.SubMConv3d(4, 4, 1), torch.nn.ReLU(),
                                  head=spconv.SubMConv3d(4, 2, 1))
    assert len(seq) == 3 and isinstance(seq[1], torch.nn.ReLU) and isinstance(seq[-1], spconv.SubMConv3d)
    y = seq(_tensor(n=9, c=4))
    assert y.features.shape == (9, 2)
    spconv.assign_name_for_sparse_modules(seq)
    assert seq[0]._sparse_unique_name == "0" and seq[-1]._sparse_unique_name == "head"
    # dense layers are skipped on an empty tensor
    empty = spconv.SparseConvTensor(torch.zeros(0, 4), torch.zeros((0, 4), dtype=torch.int32), [8] * 3, 1)
    out = spconv.SparseSequential(torch.nn.BatchNorm1d(4))(empty)
    assert out.features.shape[0] == 0
    from collections import OrderedDict
    od = spconv.SparseSequential(OrderedDict(a=spconv.SubMConv3d(4, 4, 1)))
    assert list(od._modules) == ["a"]


def test_output_size_helpers():
    from spconv_b200.pytorch import ops
    assert ops.get_conv_output_size([41, 1600, 1408], [3] * 3, [2] * 3, [1] * 3, [1] * 3) == [21, 800, 704]
    assert ops.get_conv_output_size([8, 8], [-1, 3], [1, 1], [0, 0], [1, 1]) == [1, 6]
    assert ops.get_deconv_output_size([4, 4], [3, 3], [2, 2], [1, 1], [1, 1], [1, 1]) == [8, 8]
    with pytest.raises(ValueError):
        ops.get_deconv_output_size([4], [-1], [1], [0], [1], [0])


def test_saved_weight_layout_hook_converts_legacy_checkpoints():
    """SPCONV_SAVED_WEIGHT_LAYOUT = RSKC / RSCK checkpoints load as KRSC (spconv/pytorch/conv.py:648-683)"""
    import torch
    import spconv_b200.pytorch as spconv
    from spconv_b200 import constants
    src = spconv.SparseConv3d(4, 6, [3, 2, 1], bias=False, record_voxel_count=True)
    w = src.weight.detach().clone()                                   # [6, 3, 2, 1, 4]
    try:
        for layout, perm in (("RSKC", (1, 2, 3, 0, 4)), ("RSCK", (1, 2, 3, 4, 0)), ("KRSC", (0, 1, 2, 3, 4)), ("", (0, 1, 2, 3, 4))):
            constants.SAVED_WEIGHT_LAYOUT = layout
            dst = spconv.SparseConv3d(4, 6, [3, 2, 1], bias=False, record_voxel_count=True)
            dst.load_state_dict({"weight": w.permute(*perm).contiguous()})      # no voxel-count buffer: supplied
            assert torch.equal(dst.weight, w), layout
            assert dst.get_max_num_voxels() is not None
    finally:
        constants.SAVED_WEIGHT_LAYOUT = ""


def test_fuse_bn_weights_equals_dense_conv_then_bn():
    """eval-time BN folding (example/fuse_bn_act.py:36-56) on the KRSC filter"""
    import torch
    import spconv_b200.pytorch as spconv
    torch.manual_seed(0)
    conv = torch.nn.Conv3d(4, 6, 3, bias=True).eval()
    bn = torch.nn.BatchNorm3d(6).eval()
    with torch.no_grad():
        bn.running_mean.uniform_(-1, 1)
        bn.running_var.uniform_(0.5, 2)
        bn.weight.uniform_(0.5, 1.5)
        bn.bias.uniform_(-1, 1)
    krsc = conv.weight.detach().permute(0, 2, 3, 4, 1).contiguous()
    wf, bf = spconv.fuse_bn_weights(krsc, conv.bias.detach(), bn.running_mean, bn.running_var, bn.eps,
                                    bn.weight.detach(), bn.bias.detach())
    x = torch.randn(2, 4, 8, 8, 8)
    got = torch.nn.functional.conv3d(x, wf.detach().permute(0, 4, 1, 2, 3), bf.detach())
    assert (bn(conv(x)) - got).abs().max() < 1e-5
    # no conv bias / no affine
    wf2, bf2 = spconv.fuse_bn_weights(krsc, None, bn.running_mean, bn.running_var, bn.eps, None, None)
    ref2 = (torch.nn.functional.conv3d(x, conv.weight, None) - bn.running_mean.view(1, -1, 1, 1, 1)) \
        * torch.rsqrt(bn.running_var + bn.eps).view(1, -1, 1, 1, 1)
    got2 = torch.nn.functional.conv3d(x, wf2.detach().permute(0, 4, 1, 2, 3), bf2.detach())
    assert (ref2 - got2).abs().max() < 1e-5


def test_prefetchable_chain_follows_module_order_and_stops_where_it_cannot():
    """RulebookPrefetcher's layer discovery (host logic, no device): keyed layers in application order, strided ones
    included; a layer without indice_key, with the Native algo, or an inverse conv ends the chain"""
    import spconv_b200.pytorch as spconv
    from spconv_b200.core import ConvAlgo
    from spconv_b200.pytorch.prefetch import input_level_subm_layers, prefetchable_chain
    net = spconv.SparseSequential(
        spconv.SubMConv3d(4, 8, 3, indice_key="a"), spconv.SubMConv3d(8, 8, 3, indice_key="a"),
        spconv.SparseConv3d(8, 16, 3, 2, 1, indice_key="d1"), spconv.SubMConv3d(16, 16, 3, indice_key="b"),
        spconv.SubMConv3d(16, 16, 1, indice_key="b"),               # 1x1: no rulebook, skipped
        spconv.SparseConv3d(16, 16, 3, 2, 1),                        # no key: the chain ends here
        spconv.SubMConv3d(16, 16, 3, indice_key="c"))
    assert [m.indice_key for m in prefetchable_chain(net)] == ["a", "a", "d1", "b"]
    assert [m.indice_key for m in input_level_subm_layers(net)] == ["a"]
    native = [spconv.SubMConv3d(4, 4, 3, indice_key="n", algo=ConvAlgo.Native), spconv.SubMConv3d(4, 4, 3, indice_key="m")]
    assert prefetchable_chain(native) == []
    inv = [spconv.SparseConv3d(4, 4, 3, 2, 1, indice_key="d"), spconv.SparseInverseConv3d(4, 4, 3, indice_key="d")]
    assert [m.indice_key for m in prefetchable_chain(inv)] == ["d"]
    pre = spconv.RulebookPrefetcher(net)                              # construction needs no device
    assert len(pre.layers) == 4 and pre.stream is None
