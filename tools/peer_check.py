"""Two (or more) processes, one GPU each: the fused weight-gradient exchange over CUDA IPC peer memory
against NCCL.  Run:  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/peer_check.py
Prints "peer_check OK" on rank 0; exit code != 0 on any mismatch."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from spconv_b200.core import ConvAlgo  # noqa: E402
from spconv_b200.pytorch import ops  # noqa: E402
from spconv_b200.pytorch.dist import PeerGroup  # noqa: E402
from tests.util import random_cloud  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank)))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    peers = PeerGroup(capacity_bytes=4 << 20, average=True)
    shape = [19, 18, 17]
    worst = 0.0
    for it, (C, K, subm, dt) in enumerate([(64, 64, True, torch.float16), (64, 128, False, torch.bfloat16),
                                           (32, 32, True, torch.float16), (48, 24, True, torch.float16)] * 3):
        rng = np.random.default_rng(7000 + 31 * it + rank)              # every rank its own shard
        feats, inds = random_cloud(rng, shape, [1000 + 100 * rank, 800], C)
        st = [1] * 3 if subm else [2] * 3
        res = ops.get_indice_pairs_implicit_gemm(torch.from_numpy(inds).to(dev), 2, shape, ConvAlgo.MaskImplicitGemm, [3] * 3,
                                                 st, [1] * 3, [1] * 3, [0] * 3, subm, False, is_train=True)
        x = torch.from_numpy(feats).to(dev, dt)
        dout = torch.from_numpy(rng.uniform(-0.2, 0.2, size=(res[0].shape[0], K)).astype(np.float32)).to(dev, dt)
        w = torch.from_numpy(np.random.default_rng(it).uniform(-0.5, 0.5, size=(K, 3, 3, 3, C)).astype(np.float32)).to(dev, dt)  # replicated
        args = (x, w, dout, res[2], res[3], res[4], res[5], res[6], res[7], None, res[8], 128, subm)
        ops.set_peer_group(None)
        local = ops.implicit_gemm_backward(*args)[1].float()
        want = local.clone()
        dist.all_reduce(want)                                             # fp32 sum of the rounded per-rank results
        mag = local.abs()
        dist.all_reduce(mag)
        ops.set_peer_group(peers)
        got = ops.implicit_gemm_backward(*args)[1]
        ulp = 2.0 ** -7 if dt == torch.bfloat16 else 2.0 ** -10
        tol = ulp * (mag / world + (want / world).abs()) + 1e-6
        err = (got.float() - want / world).abs()
        assert bool((err <= tol).all()), (rank, it, float(err.max()))
        worst = max(worst, float((err / tol).max()))
        # bit-identical replicas
        mine = got.view(torch.int16).to(torch.int32)
        lo, hi = mine.clone(), mine.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        assert torch.equal(lo, hi), (rank, it, "replicas differ")
    # small tensors, in place, fp32, sum
    ops.set_peer_group(None)
    peers.scale = 1.0
    ops.set_peer_group(peers)
    for n in (3, 1001, 40000):
        t = torch.arange(n, device=dev, dtype=torch.float32) * (rank + 1)
        ops.peer_allreduce_(t)
        want = torch.arange(n, device=dev, dtype=torch.float32) * (world * (world + 1) // 2)
        assert torch.equal(t, want), (rank, n)
    assert peers.error() == 0
    # timing: 200 fused exchanges of a C = K = 64 gradient vs NCCL on the same bytes
    x, w, dout = args[0], args[1], args[2]
    t = torch.zeros(27 * 64 * 64, device=dev, dtype=torch.float16)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    for _ in range(20):
        ops.peer_allreduce_(t)
        dist.all_reduce(t)
    dist.barrier()
    torch.cuda.synchronize()
    ev[0].record()
    for _ in range(200):
        ops.peer_allreduce_(t)
    ev[1].record()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    ev[2].record()
    for _ in range(200):
        dist.all_reduce(t)
    ev[3].record()
    torch.cuda.synchronize()
    us_peer, us_nccl = ev[0].elapsed_time(ev[1]) * 5, ev[2].elapsed_time(ev[3]) * 5
    ops.set_peer_group(None)
    dist.barrier()
    peers.close()
    if rank == 0:
        print(f"peer_check OK world {world} worst err/tol {worst:.3f}; 221 KB fp16 gradient: peer exchange {us_peer:.1f} us, "
              f"NCCL all_reduce {us_nccl:.1f} us per call (back-to-back launches)")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
