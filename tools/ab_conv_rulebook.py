"""A/B of the regular-conv rulebook (3x3x3, stride 2) through the fused one-call path:
debug bit 128 = round-1 path (full table + collect + CUB sort), 2048 = bitmap ranking but the two mask
sorts / tile tables one after the other, 0 = default (bitmap ranking, paired sorts and tile tables)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench_utils import surface_cloud
from spconv_b200 import _cabi
from spconv_b200.core import ConvAlgo
from spconv_b200.pytorch import ops

dev = torch.device("cuda:0")
flush = torch.empty(64 << 20, device=dev)
lib = _cabi.load()
for shape, n in (([41, 1600, 1408], 100_000), ([41, 1440, 1440], 300_000)):
    rng = np.random.default_rng(50051)
    inds = torch.from_numpy(surface_cloud(rng, shape, n)).to(dev)
    keep = None
    for dbg, name in ((128, "round-1"), (2048, "rank"), (0, "rank+paired")):
        _cabi.check(lib.spx_debug_configure(-1, 0, dbg, None, 0), "cfg")
        ts = []
        for rep in range(12):
            flush.zero_(); torch.cuda._sleep(400_000)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            res = ops.get_indice_pairs_implicit_gemm(inds, 1, shape, ConvAlgo.MaskImplicitGemm, [3] * 3, [2] * 3, [1] * 3,
                                                     [1] * 3, [0] * 3, False, False, is_train=True)
            e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        got = [t.clone() for t in (res[0], res[2], res[3], res[4][0], res[5][0], res[6][0], res[7][0])]
        tabs = [res[6][0]._spx_tile_cache[1].clone(), res[7][0]._spx_tile_cache[1].clone()] if hasattr(res[6][0], "_spx_tile_cache") else []
        same = "" if keep is None else f"identical to round-1: {all(bool(torch.equal(a, b)) for a, b in zip(keep, got + tabs))}"
        if keep is None:
            keep = got + tabs
        print(f"conv k3s2 N={n} M={res[0].shape[0]} {name:12s} median {np.median(ts):7.1f} us  min {min(ts):7.1f} us  {same}")
_cabi.check(lib.spx_debug_configure(-1, 0, 0, None, 0), "cfg")
