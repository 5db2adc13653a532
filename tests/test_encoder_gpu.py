"""BASELINE.json configs[2]: the SECOND-style 6-layer sparse encoder
(SubM16 x2 -> SparseConv 16->32 s2 -> SubM32 -> SparseConv 32->64 s2 -> SparseConv 64->128 s2),
fp16 forward + backward, checked LAYER BY LAYER against the CPU oracle:

* every layer's output ``indices`` and masked-implicit-GEMM tables are bit-exact;
* every layer's output features, input gradient and weight gradient are within the reference's
  fp16 tolerance (``test/test_all_algo.py:325-329``: ``||err||_2 < 10 * max(C, K) / 16`` ... we use
  the tighter relative form, rel-L2 <= 1e-2) of the fp32 oracle evaluated on THIS engine's
  fp16 inputs of that layer (so errors do not compound across layers);
* the two SubM(16) layers share one ``indice_key`` and therefore one rulebook
  (``spconv/pytorch/conv.py:247-319,345-444``).

Reference pattern for the net: ``test/fake_train.py:52-99``.
"""
import numpy as np
import pytest
import torch

from bench_utils import ENCODER6_LAYERS, make_encoder6, surface_cloud
from tests.util import rel_l2

pytestmark = pytest.mark.gpu


def _cloud(rng, shape, n_per_sample, batch):
    inds = surface_cloud(rng, shape, n_per_sample, batch=batch)
    feats = rng.uniform(-1, 1, size=(inds.shape[0], 16)).astype(np.float32)
    return inds, feats


@pytest.mark.parametrize("algo_name", ["MaskImplicitGemm", "Native"])
def test_encoder6_layerwise_vs_oracle(algo_name, oracle, cuda_dev):
    import spconv_b200.pytorch as spconv
    from spconv_b200.core import ConvAlgo
    from spconv_b200.pytorch import ops
    algo = ConvAlgo[algo_name]
    rng = np.random.default_rng(2024)
    shape, batch = [24, 200, 176], 2
    inds, feats = _cloud(rng, shape, 6000, batch)
    torch.manual_seed(7)
    layers = [m.to(cuda_dev).half() for m in make_encoder6(spconv, algo=algo)]
    with torch.no_grad():
        for m in layers:                     # keep fp16 activations O(1) through six linear layers
            m.weight.mul_(3.0)
    x0 = torch.from_numpy(feats).to(cuda_dev).half().requires_grad_(True)
    x = spconv.SparseConvTensor(x0, torch.from_numpy(inds).to(cuda_dev), shape, batch)
    ops.launch_count(reset=True)
    acts = [x]
    for m in layers:
        y = m(acts[-1])
        y.features.retain_grad()
        acts.append(y)
    assert ops.last_kernel_family() == 2, "tcgen05 kernels did not serve the encoder"
    assert set(acts[-1].indice_dict) == {"subm1", "down1", "subm2", "down2", "down3"}
    g = torch.Generator(device=cuda_dev).manual_seed(11)
    G = (torch.rand(acts[-1].features.shape, device=cuda_dev, generator=g) * 2 - 1).half()
    (acts[-1].features.float() * G.float()).sum().backward()
    torch.cuda.synchronize()

    cur_inds, cur_shape = inds, list(shape)
    seen_keys = {}
    for li, ((kind, c_in, c_out, key), m) in enumerate(zip(ENCODER6_LAYERS, layers)):
        subm = kind == "subm"
        st = [1] * 3 if subm else [2] * 3
        o_inds, pairs, num = oracle.get_indice_pairs(cur_inds, batch, cur_shape, [3] * 3, st, [1] * 3, [1] * 3,
                                                     [0] * 3, subm)
        got = acts[li + 1]
        # ---- rulebook: bit-exact
        assert np.array_equal(got.indices.cpu().numpy(), o_inds), f"layer {li}: out indices differ"
        out_shape = cur_shape if subm else oracle.get_conv_output_size(cur_shape, [3] * 3, st, [1] * 3, [1] * 3)
        assert got.spatial_shape == out_shape
        datas = got.indice_dict[key]
        if algo == ConvAlgo.MaskImplicitGemm:
            ref = oracle.implicit_gemm_tables(pairs, num, cur_inds.shape[0], o_inds.shape[0], subm)
            assert np.array_equal(datas.pair_fwd.cpu().numpy(), ref["pair_fwd"]), f"layer {li}: pair_fwd"
            assert np.array_equal(datas.pair_bwd.cpu().numpy(), ref["pair_bwd"]), f"layer {li}: pair_bwd"
            assert np.array_equal(datas.mask_argsort_fwd_splits[0].cpu().numpy(), ref["argsort_fwd"])
            assert np.array_equal(datas.pair_mask_fwd_splits[0].cpu().numpy().view(np.uint32).reshape(-1),
                                  ref["mask_fwd"].reshape(-1))
            if not subm:
                assert np.array_equal(datas.mask_argsort_bwd_splits[0].cpu().numpy(), ref["argsort_bwd"])
        else:
            assert np.array_equal(datas.indice_pairs.cpu().numpy(), pairs), f"layer {li}: native pairs"
            assert np.array_equal(datas.indice_pair_num.cpu().numpy(), num)
        if key in seen_keys:                 # indice_key reuse: the very same rulebook object
            assert seen_keys[key] is datas
        seen_keys[key] = datas
        # ---- arithmetic, on this engine's own fp16 inputs of the layer
        x_in = acts[li].features.detach().float().cpu().numpy()
        w = m.weight.detach().float().cpu().numpy()
        ref_out = oracle.indice_conv(x_in, w, pairs, num, o_inds.shape[0], False, subm)
        out = got.features.detach().float().cpu().numpy()
        assert rel_l2(out, ref_out) < 1e-2, f"layer {li} fwd rel-l2 {rel_l2(out, ref_out):.3e}"
        assert np.linalg.norm(out - ref_out) < 10 * max(c_in, c_out) / 16 * max(1.0, np.abs(ref_out).max())
        dout = got.features.grad.detach().float().cpu().numpy()
        ref_din, ref_dw = oracle.indice_conv_backward(x_in, w, dout, pairs, num, False, subm)
        din = acts[li].features.grad.detach().float().cpu().numpy()
        dw = m.weight.grad.detach().float().cpu().numpy()
        assert rel_l2(din, ref_din) < 1e-2, f"layer {li} dgrad rel-l2 {rel_l2(din, ref_din):.3e}"
        assert rel_l2(dw, ref_dw) < 1e-2, f"layer {li} wgrad rel-l2 {rel_l2(dw, ref_dw):.3e}"
        cur_inds, cur_shape = o_inds, out_shape
    # two SubM(16) layers, one rulebook: 5 rulebooks for 6 layers
    assert len(seen_keys) == 5


def test_encoder6_sequential_with_relu_trains(cuda_dev):
    """The same net through SparseSequential with ReLU in between: one optimizer step lowers the loss
    (the reference's smoke criterion, test/fake_train.py:run)."""
    import spconv_b200.pytorch as spconv
    rng = np.random.default_rng(5)
    shape, batch = [24, 200, 176], 2
    inds, feats = _cloud(rng, shape, 5000, batch)
    torch.manual_seed(3)
    net = spconv.SparseSequential(*make_encoder6(spconv, relu=True, bias=True)).to(cuda_dev)
    opt = torch.optim.SGD(net.parameters(), lr=1e-2)
    d_inds = torch.from_numpy(inds).to(cuda_dev)
    d_feats = torch.from_numpy(feats).to(cuda_dev)
    losses = []
    for _ in range(3):
        with torch.autocast("cuda", dtype=torch.float16):
            y = net(spconv.SparseConvTensor(d_feats, d_inds, shape, batch))
        loss = (y.features.float() - 1.0).square().mean()
        opt.zero_grad()
        loss.backward()
        opt.step()
        losses.append(float(loss))
    assert all(np.isfinite(losses)) and losses[-1] < losses[0], losses
    assert y.spatial_shape == [3, 25, 22] and y.features.shape[1] == 128
