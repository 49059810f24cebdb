"""Multi-GPU check of the distributed six-step transform; run under torchrun, one rank per GPU:
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/dist_check.py [log2 n1] [log2 n2] [peer|nccl]
Every rank generates its block of the hash-generated input; rank 0 gathers the distributed result and
compares it with the single-GPU plan and, for N <= 2^22, with the oracle."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fourier_b200 as fb  # noqa: E402
from fourier_b200.distributed import CudaBackend, DistributedFft  # noqa: E402

k1 = int(sys.argv[1]) if len(sys.argv) > 1 else 11
k2 = int(sys.argv[2]) if len(sys.argv) > 2 else 11
mode = sys.argv[3] if len(sys.argv) > 3 else "peer"
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
fb.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n1, n2 = 1 << k1, 1 << k2
n = n1 * n2
blk = n // world
plan = DistributedFft(n1, n2, rank, world, CudaBackend("f32"), exchange=mode)
x, scratch = plan.buffers()
fb.fill_input(x.view(1, blk), first_transform=rank)        # rows of a (world x blk) batch == blocks of one signal
out = plan.transform(x, scratch)
out = plan.transform(out, x if out is scratch else scratch, forward=False) / n      # and back again
back = torch.empty_like(out)
fb.fill_input(back.view(1, blk), first_transform=rank)
rt = float((out - back).abs().max())
out = plan.transform(back.clone() if mode == "nccl" else x.copy_(back), scratch)
torch.cuda.synchronize()
parts = [torch.empty_like(out) for _ in range(world)] if rank == 0 else None
dist.gather(torch.view_as_real(out).contiguous(), [torch.view_as_real(p) for p in parts] if rank == 0 else None, dst=0)
if rank == 0:
    got = torch.cat(parts)
    full = torch.empty(n, dtype=torch.complex64, device="cuda")
    fb.fill_input(full.view(1, n))
    ref = torch.empty_like(full)
    fb.create_fft_f32(n).transform(full.view(1, n), ref.view(1, n), fb.Transform.Fft)
    err = float((got - ref).abs().max() / ref.abs().max())
    print(f"distributed N=2^{k1 + k2} over {world} GPUs ({mode}) vs single-GPU plan: max rel err {err:.3e}; "
          f"forward+inverse round trip max abs err {rt:.3e}")
    ok = err < 1e-5 and rt < 1e-4
    if n <= 1 << 22:
        from oracle import oracle as O
        want = O.transform(O.fill_input(1, n, np.complex64)[0], O.FFT)
        e2 = float(np.abs(got.cpu().numpy() - want).max() / np.abs(want).max())
        print(f"  vs oracle: max rel err {e2:.3e}")
        ok = ok and e2 < 1e-5
    print("DIST CHECK", "OK" if ok else "FAILED")
dist.destroy_process_group()
