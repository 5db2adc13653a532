#!/usr/bin/env python
"""Benchmark of the rulebook -> implicit-GEMM hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W              # this engine, one rank per GPU
    python bench.py --impl reference --gpus N --steps K ...    # reference CPU algorithm (oracle port)

A "step" is one pass of the hot path over one point cloud per GPU: SubM rulebook generation
(hash + probe + mask sort) -> SubMConv3d 3x3x3 forward -> backward (input grad + weight grad);
at N > 1 every rank processes its own cloud (weak scaling) and the weight gradient is
all-reduced once per step over NCCL.  ``value`` = active voxels of all ranks / step time with the
inputs resident in HBM; ``e2e`` = the same metric through the public module API starting from
pinned HOST buffers (H2D of coordinates + features and D2H of loss + weight gradient inside the
timed region).  See DESIGN.md section "Measurement".
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
from typing import Dict, List

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from bench_utils import (ClockSampler, algorithmic_bytes, conv_flops, load_peaks,  # noqa: E402
                         surface_cloud)

WORKLOADS = {
    # BASELINE.json configs[1]: the configuration the metric is quoted on
    "submconv3d_k3_c64_fp16_100k_kitti": dict(shape=[41, 1600, 1408], n=100_000, c_in=64, c_out=64,
                                              dtype="fp16", subm=True, ksize=3, stride=1, padding=1),
    # BASELINE.json configs[3]
    "sparseconv3d_k3s2_c64_128_bf16_300k": dict(shape=[41, 1440, 1440], n=300_000, c_in=64, c_out=128,
                                                dtype="bf16", subm=False, ksize=3, stride=2, padding=1),
}
DEFAULT_WORKLOAD = "submconv3d_k3_c64_fp16_100k_kitti"
METRIC = "active-voxels/sec fwd+bwd SubMConv3d 3^3 C=64"


def metric_name(workload: str) -> str:
    """BASELINE.json's metric for the default workload; other workloads are labelled as what they are."""
    return METRIC if workload == DEFAULT_WORKLOAD else f"active-voxels/sec fwd+bwd {workload}"
NUM_CLOUDS = 4          # distinct clouds per rank, rotated so consecutive steps never share inputs
L2_FLUSH_BYTES = 256 << 20


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--graph", type=int, default=1, help="replay the device-resident step from CUDA graphs")
    ap.add_argument("--cpu-sample", type=int, default=100_000, help="voxels in the CPU-baseline sample")
    return ap.parse_args()


# ============================================================================ CPU reference arm
def cpu_reference_step(orc, inds, feats, w, dout, wl):
    """The reference's CPU path for one cloud: rulebook (single-threaded hash map, as
    spconv/csrc/sparse/indices.py:1640-1778) + gather/mm/scatter forward and backward with the
    host BLAS on all cores (convops.py:1606-1633, :1831-1860)."""
    nd = 3
    ks, st, pd = [wl["ksize"]] * nd, [wl["stride"]] * nd, [wl["padding"]] * nd
    t0 = time.perf_counter()
    out_inds, pairs, num = orc.get_indice_pairs(inds, 1, wl["shape"], ks, st, pd, [1] * nd, [0] * nd,
                                                wl["subm"])
    t1 = time.perf_counter()
    m = out_inds.shape[0]
    orc.indice_conv(feats, w, pairs, num, m, False, wl["subm"])
    t2 = time.perf_counter()
    orc.indice_conv_backward(feats, w, dout[:m], pairs, num, False, wl["subm"])
    t3 = time.perf_counter()
    return {"rulebook_s": t1 - t0, "fwd_s": t2 - t1, "bwd_s": t3 - t2, "total_s": t3 - t0, "n": inds.shape[0]}


CPU_THREAD_CAP = 16     # the small per-offset GEMMs get SLOWER with more BLAS threads (measured: 128 -> 3.2 s/step)


def cpu_threads() -> int:
    return min(os.cpu_count() or 1, CPU_THREAD_CAP)


def limit_blas_threads():
    """The reference's CPU path runs its per-offset mm on the host BLAS; give it the thread count
    where it is fastest on this box instead of oversubscribing every core."""
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=cpu_threads())
    except Exception:
        pass


def make_cpu_sample(wl, n, seed):
    rng = np.random.default_rng(seed)
    scale = max(n / wl["n"], 1e-3) ** 0.5
    shape = [wl["shape"][0], max(64, int(wl["shape"][1] * scale)), max(64, int(wl["shape"][2] * scale))]
    inds = surface_cloud(rng, shape, n)
    feats = rng.uniform(-1, 1, size=(inds.shape[0], wl["c_in"])).astype(np.float32)
    w = rng.uniform(-1, 1, size=(wl["c_out"], 3, 3, 3, wl["c_in"])).astype(np.float32)
    dout = rng.uniform(-0.2, 0.2, size=(inds.shape[0] * 2, wl["c_out"])).astype(np.float32)
    wl_s = dict(wl)
    wl_s["shape"] = shape
    return inds, feats, w, dout, wl_s


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as orc
    orc.build()
    limit_blas_threads()
    wl = WORKLOADS[args.workload]
    inds, feats, w, dout, wl_s = make_cpu_sample(wl, args.cpu_sample, 1234)
    for _ in range(min(args.warmup, 2)):
        cpu_reference_step(orc, inds, feats, w, dout, wl_s)
    recs = []
    t_begin = time.perf_counter()
    for _ in range(args.steps):
        recs.append(cpu_reference_step(orc, inds, feats, w, dout, wl_s))
        if time.perf_counter() - t_begin > 150:          # keep the arm within a few minutes
            break
    tot = sum(r["total_s"] for r in recs)
    steps = len(recs)
    value = inds.shape[0] * steps / tot
    cores = cpu_threads()
    sample = (f"{inds.shape[0]} voxels of the same generator in a {wl_s['shape']} grid, fp32, "
              f"single-threaded rulebook + numpy/BLAS gather-mm-scatter fwd+bwd")
    line = {
        "impl": "reference", "metric": metric_name(args.workload), "value": value, "unit": "voxels/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": min(args.warmup, 2), "ms_per_step": 1e3 * tot / steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32",
        "data": "synthetic",
        "config": {"workload": args.workload, "sample": sample, "impl_note":
                   "reference cannot be built here (pccm/cumm absent): oracle port of its CPU algorithm"},
        "cpu_baseline": {"value": value, "unit": "voxels/s", "cores": cores, "kind": "port", "sample": sample,
                         "rulebook_ms": 1e3 * sum(r["rulebook_s"] for r in recs) / steps,
                         "fwd_ms": 1e3 * sum(r["fwd_s"] for r in recs) / steps,
                         "bwd_ms": 1e3 * sum(r["bwd_s"] for r in recs) / steps},
        "e2e": {"value": value, "unit": "voxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ============================================================================ GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist

    import spconv_b200.pytorch as spconv
    from spconv_b200.core import ConvAlgo
    from spconv_b200.pytorch import ops
    from spconv_b200.pytorch.core import CUDAKernelTimer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py (impl=ours) needs a CUDA device; there is no CPU path"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    wl = WORKLOADS[args.workload]
    tdt = {"fp16": torch.float16, "bf16": torch.bfloat16, "fp32": torch.float32}[wl["dtype"]]
    elem = 2 if wl["dtype"] != "fp32" else 4
    nd = 3
    ks, st, pd, dl = [wl["ksize"]] * nd, [wl["stride"]] * nd, [wl["padding"]] * nd, [1] * nd
    kv = wl["ksize"] ** nd
    C, K = wl["c_in"], wl["c_out"]

    # ---------------- synthetic inputs: NUM_CLOUDS clouds per rank, pinned on the host + resident in HBM
    clouds = []
    for i in range(NUM_CLOUDS):
        rng = np.random.default_rng(50051 + 1000 * rank + i)
        inds_np = surface_cloud(rng, wl["shape"], wl["n"])
        feats_np = rng.uniform(-1, 1, size=(inds_np.shape[0], C)).astype(np.float32)
        h_inds = torch.from_numpy(inds_np).pin_memory()
        h_feats = torch.from_numpy(feats_np).to(tdt).pin_memory()
        clouds.append(dict(h_inds=h_inds, h_feats=h_feats, d_inds=h_inds.to(dev), d_feats=h_feats.to(dev),
                           n=inds_np.shape[0]))
    torch.manual_seed(48848)
    conv_cls = spconv.SubMConv3d if wl["subm"] else spconv.SparseConv3d
    layer = conv_cls(C, K, wl["ksize"], wl["stride"], wl["padding"], bias=False,
                     algo=ConvAlgo.MaskImplicitGemm).to(dev).to(tdt)
    layer.train()
    weight = layer.weight.detach()
    n_per_step = sum(c["n"] for c in clouds) / NUM_CLOUDS

    # per-cloud upstream gradient (device resident); output count known after one rulebook build
    for c in clouds:
        res = ops.get_indice_pairs_implicit_gemm(c["d_inds"], 1, wl["shape"], ConvAlgo.MaskImplicitGemm, ks, st,
                                                 pd, dl, [0] * nd, wl["subm"], False, is_train=True)
        c["m"] = res[0].shape[0]
        c["pairs_total"] = int((res[2] >= 0).sum().item())
        g = torch.Generator(device=dev).manual_seed(7)
        c["dout"] = (torch.rand((c["m"], K), device=dev, generator=g) * 0.4 - 0.2).to(tdt)
    flush = torch.empty(L2_FLUSH_BYTES // 4, dtype=torch.float32, device=dev)

    def device_step(c, timer=None):
        """rulebook -> forward -> backward on device-resident inputs (operator layer)."""
        kw = {} if timer is None else {"timer": timer}
        res = ops.get_indice_pairs_implicit_gemm(c["d_inds"], 1, wl["shape"], ConvAlgo.MaskImplicitGemm, ks, st,
                                                 pd, dl, [0] * nd, wl["subm"], False, is_train=True, **kw)
        out_inds, _, pair_fwd, pair_bwd, mask_fwd, mask_bwd, sort_fwd, sort_bwd, masks = res
        out, mask_out, mw = ops.implicit_gemm(c["d_feats"], weight, pair_fwd, mask_fwd, sort_fwd,
                                              out_inds.shape[0], masks, True, wl["subm"], **kw)
        din, dw = ops.implicit_gemm_backward(c["d_feats"], weight, c["dout"], pair_fwd, pair_bwd, mask_fwd,
                                             mask_bwd, sort_fwd, sort_bwd, mask_out, masks, mw, wl["subm"], **kw)
        return out, din, dw

    def allreduce(t):
        if world > 1:
            dist.all_reduce(t)

    # ---------------- warm-up (also configures kernels / NCCL before any graph capture)
    for i in range(max(args.warmup, 3)):
        out, din, dw = device_step(clouds[i % NUM_CLOUDS])
        allreduce(dw)
    torch.cuda.synchronize()

    # ---------------- optional CUDA graphs of the device-resident step (one per cloud)
    graphs, graph_out = None, None
    use_graph = bool(args.graph) and wl["subm"]     # regular conv has a host sync (output count)
    if use_graph:
        try:
            graphs, graph_out = [], []
            for c in clouds:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    o = device_step(c)
                graphs.append(g)
                graph_out.append(o)
            torch.cuda.synchronize()
        except Exception as e:                       # capture is an optimisation, never a requirement
            print(f"[bench] CUDA-graph capture failed ({type(e).__name__}: {e}); running eagerly", file=sys.stderr)
            graphs, graph_out, use_graph = None, None, False
            torch.cuda.synchronize()

    def timed_loop(step_fn, steps):
        """K steps; CUDA events on the launching stream around every step, L2 flushed in between
        (outside the events); returns per-step ms."""
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        for i in range(steps):
            flush.zero_()
            evs[i][0].record()
            step_fn(i)
            evs[i][1].record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        return [a.elapsed_time(b) for a, b in evs]

    def value_step(i):
        j = i % NUM_CLOUDS
        if use_graph:
            graphs[j].replay()
            dw = graph_out[j][2]
        else:
            dw = device_step(clouds[j])[2]
        allreduce(dw)

    # ---------------- e2e: public module API from pinned host buffers
    # Eager: every step copies this step's coordinates + features H2D, runs
    # SparseConvTensor -> layer -> loss -> backward, and reads loss + dW back.
    h_loss = torch.zeros((), dtype=torch.float32).pin_memory()
    h_dw = torch.zeros_like(weight, device="cpu").pin_memory()

    def e2e_body(d_inds, d_feats):
        xf = d_feats.detach().requires_grad_(True)
        x = spconv.SparseConvTensor(xf, d_inds, wl["shape"], 1)
        layer.weight.grad = None
        y = layer(x)
        loss = y.features.square().mean(dtype=torch.float32)
        loss.backward()
        return loss

    def e2e_step_eager(i):
        c = clouds[i % NUM_CLOUDS]
        d_inds = c["h_inds"].to(dev, non_blocking=True)
        d_feats = c["h_feats"].to(dev, non_blocking=True)
        loss = e2e_body(d_inds, d_feats)
        allreduce(layer.weight.grad)
        h_loss.copy_(loss.detach(), non_blocking=True)
        h_dw.copy_(layer.weight.grad, non_blocking=True)

    for i in range(3):
        e2e_step_eager(i)
    torch.cuda.synchronize()

    # Graph-captured e2e step with double buffering: the replay of step i computes on device
    # buffer i%2 and, on a forked stream inside the same graph, copies the NEXT cloud H2D into
    # buffer (i+1)%2 -- so every timed step still contains one full H2D of a step's inputs and
    # the D2H of its results, but the copy overlaps the kernels.
    e2e_graphs = None
    if use_graph:
        try:
            n_max = max(c["n"] for c in clouds)
            bufs = [dict(inds=torch.empty((n_max, 4), dtype=torch.int32, device=dev),
                         feats=torch.empty((n_max, C), dtype=tdt, device=dev)) for _ in range(2)]
            side = torch.cuda.Stream()
            e2e_graphs = []
            for j, c in enumerate(clouds):
                cur, nxt = bufs[j % 2], bufs[(j + 1) % 2]
                cn = clouds[(j + 1) % NUM_CLOUDS]
                cur["inds"][:c["n"]].copy_(c["h_inds"])
                cur["feats"][:c["n"]].copy_(c["h_feats"])
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    main = torch.cuda.current_stream()
                    side.wait_stream(main)
                    with torch.cuda.stream(side):
                        nxt["inds"][:cn["n"]].copy_(cn["h_inds"], non_blocking=True)
                        nxt["feats"][:cn["n"]].copy_(cn["h_feats"], non_blocking=True)
                    loss = e2e_body(cur["inds"][:c["n"]], cur["feats"][:c["n"]])
                    if world == 1:
                        h_loss.copy_(loss.detach(), non_blocking=True)
                        h_dw.copy_(layer.weight.grad, non_blocking=True)
                    main.wait_stream(side)
                e2e_graphs.append((g, loss, layer.weight.grad))
            torch.cuda.synchronize()
            # prologue: cloud 0 must be resident in buffer 0 before the first replay
            bufs[0]["inds"][:clouds[0]["n"]].copy_(clouds[0]["h_inds"])
            bufs[0]["feats"][:clouds[0]["n"]].copy_(clouds[0]["h_feats"])
            torch.cuda.synchronize()
        except Exception as e:
            print(f"[bench] e2e CUDA-graph capture failed ({type(e).__name__}: {e}); e2e runs eagerly",
                  file=sys.stderr)
            e2e_graphs = None
            torch.cuda.synchronize()

    def e2e_step(i):
        if e2e_graphs is None:
            return e2e_step_eager(i)
        g, loss, grad = e2e_graphs[i % NUM_CLOUDS]
        g.replay()
        if world > 1:
            allreduce(grad)
            h_loss.copy_(loss.detach(), non_blocking=True)
            h_dw.copy_(grad, non_blocking=True)

    # kernels of THIS library per step (graph replays re-issue exactly the captured launches)
    ops.launch_count(reset=True)
    device_step(clouds[0])
    launches_per_step = ops.launch_count(reset=True)
    torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_value = timed_loop(value_step, args.steps)
    launches = launches_per_step * args.steps
    ms_e2e = timed_loop(e2e_step, args.steps)
    ms_e2e_eager = timed_loop(e2e_step_eager, args.steps)
    clocks = sampler.stop() if rank == 0 else {}

    # ---------------- per-kernel timing for the roofline (events around every C-ABI region)
    reps = max(5, min(args.steps, 20))
    samples: Dict[str, List[float]] = {}
    for i in range(reps):
        timer = CUDAKernelTimer(True)
        flush.zero_()
        # a ~0.5 ms device-side spin lets the host queue the whole step ahead of the GPU, so the
        # events bracket back-to-back kernels and not the host's launch latency
        torch.cuda._sleep(1_000_000)
        device_step(clouds[i % NUM_CLOUDS], timer)
        for k, v in timer.get_all_pair_time().items():
            samples.setdefault(k, []).append(v)
    # median over the repetitions: one host hiccup inside a region must not become the kernel's time
    regions = {k: float(np.median(v)) for k, v in samples.items()}

    # ---------------- reduce over ranks (max time, sum voxels)
    t_value = torch.tensor([float(np.mean(ms_value)), float(np.mean(ms_e2e)), float(np.mean(ms_e2e_eager))],
                           device=dev, dtype=torch.float64)
    n_total = torch.tensor([n_per_step], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_value, op=dist.ReduceOp.MAX)
        dist.all_reduce(n_total, op=dist.ReduceOp.SUM)
    ms_step, ms_step_e2e, ms_step_e2e_eager = float(t_value[0]), float(t_value[1]), float(t_value[2])
    voxels = float(n_total[0])

    if rank == 0:
        peaks = load_peaks()
        c0 = clouds[0]
        name_map = {"implicit_gemm": "fwd", "implicit_gemm_dgrad": "dgrad", "implicit_gemm_wgrad": "wgrad"}
        dom_region = max((k for k in regions if k in name_map), key=lambda k: regions[k])
        kind = name_map[dom_region]
        abytes = algorithmic_bytes(kind, c0["n"], c0["m"], C, K, kv, elem)
        achieved = abytes / (regions[dom_region] * 1e-3) / 1e9
        flops = conv_flops(c0["pairs_total"], C, K)
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "r01_traffic.json")
        if os.path.exists(tpath):
            try:
                traffic = json.load(open(tpath)).get(kind)
            except Exception:
                traffic = None
        # bounded CPU-baseline sample on this host
        from oracle import oracle as orc
        orc.build()
        s_inds, s_feats, s_w, s_dout, wl_s = make_cpu_sample(wl, args.cpu_sample, 1234)
        limit_blas_threads()
        cpu_reference_step(orc, s_inds[:2000], s_feats[:2000], s_w, s_dout, wl_s)     # warm BLAS
        recs, t0 = [], time.perf_counter()
        while len(recs) < 3 or (time.perf_counter() - t0 < 10 and len(recs) < 20):
            recs.append(cpu_reference_step(orc, s_inds, s_feats, s_w, s_dout, wl_s))
        cpu_tot = sum(r["total_s"] for r in recs)
        cpu_value = s_inds.shape[0] * len(recs) / cpu_tot
        sample = (f"{s_inds.shape[0]} voxels, same generator, {wl_s['shape']} grid, fp32, {len(recs)} reps "
                  f"({cpu_tot:.1f} s): single-threaded rulebook + numpy/BLAS gather-mm-scatter fwd+bwd")
        line = {
            "metric": metric_name(args.workload), "value": voxels / (ms_step * 1e-3), "unit": "voxels/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": wl["dtype"],
            "data": "synthetic",
            "config": {"workload": args.workload, "grid": wl["shape"], "active_voxels_per_gpu": int(n_per_step),
                       "pairs_per_voxel": round(c0["pairs_total"] / c0["n"], 2), "c_in": C, "c_out": K,
                       "step": "subm rulebook (hash+probe+mask sort) + fwd + dgrad + wgrad"
                               + (" + NCCL all-reduce(dW)" if world > 1 else ""),
                       "parallelism": f"dp{world} (one cloud per GPU)", "cuda_graph": use_graph,
                       "l2": f"{L2_FLUSH_BYTES >> 20} MiB buffer written between timed steps; "
                             f"{NUM_CLOUDS} rotating clouds"},
            "e2e": {"value": voxels / (ms_step_e2e * 1e-3), "unit": "voxels/s", "ms_per_step": ms_step_e2e,
                    "h2d_bytes_per_step": int(c0["h_inds"].numel() * 4 + c0["h_feats"].numel() * elem),
                    "d2h_bytes_per_step": int(weight.numel() * elem + 4),
                    "api": "SparseConvTensor -> SubMConv3d.forward -> loss.backward (pinned host in, loss+dW out)",
                    "cuda_graph": e2e_graphs is not None,
                    "overlap": "H2D of the next cloud on a forked stream inside the step's graph"
                               if e2e_graphs is not None else "none",
                    "eager_value": voxels / (ms_step_e2e_eager * 1e-3), "eager_ms_per_step": ms_step_e2e_eager},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "kernel_ms": {k: round(v, 4) for k, v in sorted(regions.items())},
            "roofline": {"bound": "hbm", "kernel": f"tc_gather_gemm/{kind}" if kind != "wgrad" else "tc_wgrad",
                         "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": achieved / peaks["hbm_gbs"], "traffic": traffic,
                         "peak_source": peaks["source"] + " (MEASURED_PEAKS.json hbm_gbs)"
                         if peaks["source"] == "measured" else "fallback (B200_PROFILING.md)",
                         "algorithmic_bytes": abytes, "launch_ms": regions[dom_region],
                         "tensor_tflops_fwd": flops / (regions.get("implicit_gemm", float("nan")) * 1e-3) / 1e12,
                         "tensor_frac_fwd": flops / (regions.get("implicit_gemm", float("nan")) * 1e-3) / 1e12
                         / peaks["bf16_tflops"]},
            "cpu_baseline": {"value": cpu_value, "unit": "voxels/s", "cores": cpu_threads(), "kind": "port",
                             "sample": sample,
                             "rulebook_ms": 1e3 * sum(r["rulebook_s"] for r in recs) / len(recs),
                             "fwd_ms": 1e3 * sum(r["fwd_s"] for r in recs) / len(recs),
                             "bwd_ms": 1e3 * sum(r["bwd_s"] for r in recs) / len(recs)},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
