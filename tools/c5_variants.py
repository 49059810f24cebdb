"""BASELINE configs[4] in one launch (torchrun, one rank per GPU): checks the peer-memory exchange path against
the NCCL path and a forward+inverse round trip at full size, then times the variants.
    PYTHONPATH=. python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/c5_variants.py [log2 N]"""
import json
import os
import sys

import torch
import torch.distributed as dist

import fourier_b200 as fb
from fourier_b200.distributed import CudaBackend, DistributedFft

k = int(sys.argv[1]) if len(sys.argv) > 1 else 30
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
fb.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n1, n2 = 1 << (k // 2), 1 << (k - k // 2)
n = n1 * n2
blk = n // world
be = CudaBackend("f32")
peer = DistributedFft(n1, n2, rank, world, be, exchange="peer")
nccl = DistributedFft(n1, n2, rank, world, be, exchange="nccl")
a, b = peer.buffers()
c, d = nccl.buffers()
fb.fill_input(a.view(1, blk), first_transform=rank)
c.copy_(a)
x0 = a.clone()


def allmax(v):
    t = torch.tensor([float(v)], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


res = {"log2n": k, "gpus": world}
for chunks in (1, 8):
    peer.chunks = chunks
    a.copy_(x0)
    c.copy_(x0)
    yp = peer.transform(a, b)
    yn = nccl.transform(c, d)
    scale = allmax(yn.abs().max())
    res[f"peer(chunks={chunks})_vs_nccl_max_rel_err"] = allmax((yp - yn).abs().max()) / scale
    back = peer.transform(yp, a if yp is b else b, forward=False)
    res[f"peer(chunks={chunks})_forward_inverse_max_abs_err"] = allmax((back / n - x0).abs().max())


def timed(fn, steps=10, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return allmax(e0.elapsed_time(e1) / steps)


state = {"p": (a, b), "n": (c, d)}


def step(plan, key):
    x, s = state[key]
    out = plan.transform(x, s)
    state[key] = (out, x if out is s else s)


for chunks in (1, 4, 8, 16):
    peer.chunks = chunks
    res[f"ms_peer_chunks{chunks}"] = timed(lambda: step(peer, "p"))
nccl.chunks = 8
res["ms_nccl_chunks8"] = timed(lambda: step(nccl, "n"))
r1 = n1 // world
for name, fn in (("ms_exchange_only_peer", lambda: peer._fft_then_exchange(a, b, r1, n2, 0, True, None)),
                 ("ms_row_ffts_only", lambda: be.fft_rows(a, n1, True))):
    res[name] = timed(fn)
if rank == 0:
    print("C5VARIANTS " + json.dumps(res), flush=True)
dist.destroy_process_group()
