"""Worker of tests/test_gpu_multirank.py: one process per GPU (launched by torch.distributed.run).

Checks, with REAL peer memory (CUDA IPC over NVLink) and real NCCL collectives between the ranks:
  * DistributedFft(exchange="peer") and ("nccl"): natural-order and transposed output against the single-GPU plan
    and against the oracle (N = 2^20);
  * the forward + inverse round trip;
  * peer mode re-entrancy: the result of call 1 is consumed while a deliberately delayed rank is still
    busy and the other ranks have already entered call 2 -- the entry barrier of transform() must keep the fast
    ranks from overwriting the slow rank's result buffer.
Prints "MULTIRANK OK" on rank 0 when everything passed on every rank.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fourier_b200 as fb  # noqa: E402
from fourier_b200.distributed import CudaBackend, DistributedFft  # noqa: E402
from oracle import oracle as O  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
fb.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))

K1, K2 = 10, 10
n1, n2 = 1 << K1, 1 << K2
n = n1 * n2
blk = n // world
failures = []


def check(name, cond, detail=""):
    if not cond:
        failures.append(f"rank {rank}: {name} {detail}")


def gather(t):
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather([torch.view_as_real(p) for p in parts], torch.view_as_real(t).contiguous())
    return torch.cat(parts)


# single-GPU references (every rank computes them: cheap at 2^20)
full = torch.empty(n, dtype=torch.complex64, device="cuda")
fb.fill_input(full.view(1, n))
ref = torch.empty_like(full)
fb.create_fft_f32(n).transform(full.view(1, n), ref.view(1, n), fb.Transform.Fft)
want = torch.from_numpy(O.transform(O.fill_input(1, n, np.complex64)[0], O.FFT)).cuda()
scale = float(ref.abs().max())
check("single-GPU plan vs oracle", float((ref - want).abs().max()) / scale < 1e-5)
ref_t = ref.view(n2, n1).t().contiguous().view(-1)       # Y[k1][k2] = X[k1 + n1*k2]

for mode in ("peer", "nccl"):
    plan = DistributedFft(n1, n2, rank, world, CudaBackend("f32"), exchange=mode)
    x, scratch = plan.buffers()
    mine = full[rank * blk:(rank + 1) * blk]
    for natural in (True, False):
        x.copy_(mine)
        out = plan.transform(x, scratch, natural_order=natural)
        got = gather(out)
        target = ref if natural else ref_t
        err = float((got - target).abs().max()) / scale
        check(f"{mode} natural={natural} vs single-GPU plan", err < 1e-5, f"rel err {err:.3e}")
        err_o = float((got - (want if natural else want.view(n2, n1).t().contiguous().view(-1))).abs().max()) / scale
        check(f"{mode} natural={natural} vs oracle", err_o < 1e-5, f"rel err {err_o:.3e}")
    # forward + inverse round trip
    x.copy_(mine)
    out = plan.transform(x, scratch)
    back = plan.transform(out, x if out is scratch else scratch, forward=False) / n
    rt = float((back - mine).abs().max())
    check(f"{mode} round trip", rt < 1e-4, f"abs err {rt:.3e}")
    if mode == "peer":
        # re-entrancy: rank 0 is slow to consume its result; the others rush into the next call
        x.copy_(mine)
        out1 = plan.transform(x, scratch)
        if rank == 0:
            torch.cuda._sleep(int(2e8))                      # ~0.1 s of GPU time on rank 0's stream
        kept = out1.clone()                                  # the consumer of call 1 (stream-ordered)
        other = x if out1 is scratch else scratch
        other.copy_(mine)                                    # next input (same signal), then call 2 right away
        out2 = plan.transform(other, out1)
        torch.cuda.synchronize()
        got1 = gather(kept)
        err = float((got1 - ref).abs().max()) / scale
        check("peer re-entrancy: result of call 1 intact on a delayed rank", err < 1e-5, f"rel err {err:.3e}")
        err2 = float((gather(out2) - ref).abs().max()) / scale
        check("peer re-entrancy: call 2", err2 < 1e-5, f"rel err {err2:.3e}")
    plan.close()

# the grid-limited persistent exchange kernel with real peers
os.environ["FOURIER_B200_EXCHANGE_BLOCKS"] = "64"
plan = DistributedFft(n1, n2, rank, world, CudaBackend("f32"), exchange="peer", chunks=4)
x, scratch = plan.buffers()
x.copy_(full[rank * blk:(rank + 1) * blk])
err = float((gather(plan.transform(x, scratch)) - ref).abs().max()) / scale
check("peer, 64 exchange blocks, 4 pipelined chunks", err < 1e-5, f"rel err {err:.3e}")
plan.close()
del os.environ["FOURIER_B200_EXCHANGE_BLOCKS"]

bad = torch.tensor([len(failures)], device="cuda")
dist.all_reduce(bad)
for f in failures:
    print("FAILED:", f, flush=True)
if rank == 0:
    print("MULTIRANK OK" if int(bad.item()) == 0 else f"MULTIRANK FAILED ({int(bad.item())} checks)", flush=True)
dist.destroy_process_group()
sys.exit(0 if not failures else 1)
