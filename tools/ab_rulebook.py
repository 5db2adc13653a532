"""A/B timing of the SubM rulebook pieces on the headline cloud: legacy two-kernel radix passes vs the
onesweep path (debug bit 64), and the per-region event times of one step."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench_utils import surface_cloud
from spconv_b200 import _cabi
from spconv_b200.core import ConvAlgo
from spconv_b200.pytorch import ops
from spconv_b200.pytorch.core import CUDAKernelTimer

dev = torch.device("cuda:0")
shape = [41, 1600, 1408]
flush = torch.empty(64 << 20, device=dev)
for n in (100_000, 800_000):
    rng = np.random.default_rng(50051)
    inds = torch.from_numpy(surface_cloud(rng, shape, n if n <= 100_000 else 100_000, batch=max(1, n // 100_000))).to(dev)
    bs = max(1, n // 100_000)
    for dbg, name in ((0, "two-kernel"), (512, "cooperative")):
        _cabi.check(_cabi.load().spx_debug_configure(-1, 0, dbg, None, 0), "cfg")
        acc = {}
        ref = None
        for rep in range(8):
            timer = CUDAKernelTimer(True)
            flush.zero_(); torch.cuda._sleep(800_000)
            res = ops.get_indice_pairs_implicit_gemm(inds, bs, shape, ConvAlgo.MaskImplicitGemm, [3] * 3, [1] * 3, [1] * 3,
                                                     [1] * 3, [0] * 3, True, False, is_train=True, timer=timer)
            for k, v in timer.get_all_pair_time().items():
                acc.setdefault(k, []).append(v)
        print(f"N={inds.shape[0]:7d} {name:9s}", {k: round(float(np.median(v)) * 1e3, 1) for k, v in acc.items()}, "us")
        srt = res[6][0].clone()
        if dbg == 0:
            keep = srt
        else:
            print("   argsort identical to legacy:", bool(torch.equal(keep, srt)))
_cabi.check(_cabi.load().spx_debug_configure(-1, 0, 0, None, 0), "cfg")

# ---- regular conv rulebook (configs[3] shape): legacy (full table + scan-collect + CUB) vs default
shape4 = [41, 1440, 1440]
rng = np.random.default_rng(50051)
inds4 = torch.from_numpy(surface_cloud(rng, shape4, 300_000)).to(dev)
keep = None
for dbg, name in ((128, "legacy"), (0, "default"), (512, "default+coop")):
    _cabi.check(_cabi.load().spx_debug_configure(-1, 0, dbg, None, 0), "cfg")
    acc = {}
    for rep in range(6):
        timer = CUDAKernelTimer(True)
        flush.zero_(); torch.cuda._sleep(800_000)
        res = ops.get_indice_pairs_implicit_gemm(inds4, 1, shape4, ConvAlgo.MaskImplicitGemm, [3] * 3, [2] * 3, [1] * 3,
                                                 [1] * 3, [0] * 3, False, False, is_train=True, timer=timer)
        for k, v in timer.get_all_pair_time().items():
            acc.setdefault(k, []).append(v)
    print(f"conv k3s2 N=300000 M={res[0].shape[0]} {name:8s}", {k: round(float(np.median(v)) * 1e3, 1) for k, v in acc.items()}, "us")
    if name == "legacy":
        keep = [t.clone() for t in (res[0], res[2], res[3], res[6][0], res[7][0])]
    else:
        print("   identical to legacy:", all(bool(torch.equal(a, b)) for a, b in zip(keep, (res[0], res[2], res[3], res[6][0], res[7][0]))))
_cabi.check(_cabi.load().spx_debug_configure(-1, 0, 0, None, 0), "cfg")
