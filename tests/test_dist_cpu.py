"""world_size-2 gloo tests of the data-parallel plumbing (host logic; runs on CPU)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, tmp):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from spconv_b200.pytorch.dist import GradBucket, allreduce_gradients, shard_batch
    torch.manual_seed(0)
    lin = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.Linear(3, 2))
    bucket = GradBucket(lin.parameters())
    x = torch.full((5, 4), float(rank + 1))
    lin(x).sum().backward()
    local = bucket.flat.clone()
    bucket.all_reduce(average=True)
    gathered = [torch.zeros_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    expect = sum(gathered) / world
    ok1 = torch.allclose(bucket.flat, expect, atol=1e-6)
    ok1 = ok1 and all(p.grad.data_ptr() >= bucket.flat.data_ptr() for p in lin.parameters())
    # optimizer.zero_grad() defaults to set_to_none=True and drops the aliases: the bucket must
    # pull the freshly allocated gradients back in instead of reducing stale zeros
    opt = torch.optim.SGD(lin.parameters(), lr=0.1)
    opt.zero_grad()
    assert all(p.grad is None for p in lin.parameters())
    lin(x).sum().backward()
    fresh = torch.cat([p.grad.reshape(-1) for p in lin.parameters()]).clone()
    bucket.all_reduce(average=True)
    gathered = [torch.zeros_like(fresh) for _ in range(world)]
    dist.all_gather(gathered, fresh)
    ok1 = ok1 and torch.allclose(bucket.flat, sum(gathered) / world, atol=1e-6)
    ok1 = ok1 and all(p.grad.data_ptr() >= bucket.flat.data_ptr() for p in lin.parameters())
    # stateless variant
    lin2 = torch.nn.Linear(4, 2)
    lin2(torch.full((3, 4), float(rank))).sum().backward()
    g_local = lin2.weight.grad.clone()
    allreduce_gradients(lin2, average=False)
    gl = [torch.zeros_like(g_local) for _ in range(world)]
    dist.all_gather(gl, g_local)
    ok2 = torch.allclose(lin2.weight.grad, sum(gl), atol=1e-6)
    # batch sharding: 4 samples over 2 ranks
    inds = torch.tensor([[0, 1, 1, 1], [1, 2, 2, 2], [2, 3, 3, 3], [3, 4, 4, 4], [2, 5, 5, 5]], dtype=torch.int32)
    feats = torch.arange(5).float().unsqueeze(1)
    li, lf, lbs = shard_batch(inds, feats, 4, rank, world)
    ok3 = lbs == 2 and li[:, 0].tolist() == ([0, 1, 1] if rank == 0 else [0, 1]) \
        and lf[:, 0].tolist() == ([0., 2., 4.] if rank == 0 else [1., 3.])
    open(os.path.join(tmp, f"r{rank}.txt"), "w").write(f"{int(ok1)}{int(ok2)}{int(ok3)}")
    dist.barrier()
    dist.destroy_process_group()


def test_grad_bucket_and_sharding_world2(tmp_path):
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        assert open(tmp_path / f"r{r}.txt").read() == "111", f"rank {r} failed"


def test_bench_reference_arm_is_rank0_only(tmp_path):
    """--impl reference under torchrun: rank 0 prints one JSON line, other ranks exit silently."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--steps", "1", "--warmup", "0", "--cpu-sample", "3000"], env=env,
                         capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == ""
    env["RANK"] = "0"
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--steps", "1", "--warmup", "0", "--cpu-sample", "3000"], env=env,
                         capture_output=True, text=True, timeout=120)
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["kind"] in ("reference", "port")
    assert line["e2e"]["h2d_bytes_per_step"] == 0
