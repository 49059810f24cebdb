"""Where the time of the distributed transform goes (run under torchrun, one rank per GPU):
    PYTHONPATH=. python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/c5_diag.py [log2 N]"""
import os
import sys

import torch
import torch.distributed as dist

import fourier_b200 as fb
from fourier_b200.distributed import CudaBackend, DistributedFft

k = int(sys.argv[1]) if len(sys.argv) > 1 else 28
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
fb.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n1, n2 = 1 << (k // 2), 1 << (k - k // 2)
blk = n1 * n2 // world
x = torch.empty(blk, dtype=torch.complex64, device="cuda")
s = torch.empty_like(x)
fb.fill_input(x.view(1, blk), first_transform=rank)
be = CudaBackend("f32")


def timed(name, fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    ms = torch.tensor([a.elapsed_time(b) / reps], device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"{name:64s} {ms.item():8.3f} ms", flush=True)


def a2a_pieces(K):
    piece = blk // K
    w = [be.all_to_all(s[i * piece:(i + 1) * piece], x[i * piece:(i + 1) * piece], None) for i in range(K)]
    for h in w:
        h.wait()


for K in (1, 4, 8):
    timed(f"all_to_all of the rank's block in {K} pieces (no compute)", lambda: a2a_pieces(K))
for K in (1, 2, 4, 8, 16):
    plan = DistributedFft(n1, n2, rank, world, be, chunks=K)
    r1, r2 = n1 // world, n2 // world
    timed(f"chunks={K}: exchange only (pack, all_to_all, unpack)", lambda: plan._exchange(x, s, r1, n2))
    timed(f"chunks={K}: exchange + local FFTs", lambda: plan._exchange(x, s, r1, n2, then=lambda rows, f: be.fft_rows(rows, n1, True)))
    timed(f"chunks={K}: whole transform", lambda: plan.transform(x, s))
dist.destroy_process_group()
