"""Multi-process (N > 1) host logic on CPU: world_size-2 gloo.  The batch path shards independent
transforms across ranks with no data-path collective; what the ranks share is only the
plumbing bench.py uses -- rendezvous, barrier, max-over-ranks timing -- and the shard arithmetic."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def shard(batch_total, world, rank):
    """Contiguous block partition [g*B/G, (g+1)*B/G) used by bench.py (each rank owns batch_per_gpu rows)."""
    per = batch_total // world
    return rank * per, (rank + 1) * per


def _worker(rank, world, port, n, per_rank, out_q):
    sys.path.insert(0, ROOT)
    from oracle import oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard(per_rank * world, world, rank)
    # every rank generates ITS rows of the global synthetic batch and transforms them independently
    x = O.fill_input(hi - lo, n, np.complex64, first_transform=lo)
    y = O.transform(x, O.FFT)
    dist.barrier()
    # max-over-ranks reduction of the per-rank time, as bench.py does with the device timings
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    # a checksum of checksums: gathered to rank 0 to compare with the unsharded computation
    cs = torch.tensor([float(np.abs(y).sum())], dtype=torch.float64)
    dist.all_reduce(cs, op=dist.ReduceOp.SUM)
    if rank == 0:
        out_q.put((float(t.item()), float(cs.item())))
    dist.destroy_process_group()


def test_batch_sharding_world_size_2():
    from oracle import oracle as O
    n, per_rank, world = 96, 6, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, per_rank, q)) for r in range(world)]
    for p in procs:
        p.start()
    tmax, checksum = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert tmax == float(world)
    whole = O.transform(O.fill_input(per_rank * world, n, np.complex64), O.FFT)
    assert abs(checksum - float(np.abs(whole).sum())) < 1e-3 * abs(checksum)


def test_shards_are_disjoint_and_cover_the_batch():
    for world in (1, 2, 4, 8):
        total = 4096 * world
        seen = []
        for r in range(world):
            lo, hi = shard(total, world, r)
            seen.extend(range(lo, hi))
        assert seen == list(range(total))


def test_bench_reference_arm_runs_on_cpu():
    """`bench.py --impl reference` (the reference's CPU algorithm via the oracle port) needs no GPU."""
    import json
    import subprocess
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c1",
                        "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["kind"] == "port"
    assert line["e2e"]["h2d_bytes_per_step"] == 0
