"""Times the local kernels of the distributed transform at the shapes of BASELINE configs[4]
(N = 2^30 over 8 ranks: 2^27 samples per rank, 4096 x 32768 local matrix) on ONE GPU.
    PYTHONPATH=. python tools/c5_kernels.py [chunks]"""
import sys

import torch

import fourier_b200 as fb
from fourier_b200.distributed import CudaBackend

K = int(sys.argv[1]) if len(sys.argv) > 1 else 4
P, n1 = 8, 1 << 15
rows_loc, cols = n1 // P, 1 << 15
cb = cols // P
cbk = cb // K
be = CudaBackend("f32")
x = torch.empty(rows_loc * cols, dtype=torch.complex64, device="cuda")
fb.fill_input(x.view(1, -1))
y = torch.empty_like(x)
piece = P * cbk * rows_loc


def timed(name, fn, reps=10, bytes_moved=None):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    extra = f"  {bytes_moved / ms / 1e6:8.0f} GB/s" if bytes_moved else ""
    print(f"{name:58s} {ms:8.3f} ms{extra}", flush=True)
    return ms


full = x.numel() * 8 * 2
timed("transpose, whole local matrix (4096 x 32768)", lambda: be.transpose(x, y, rows_loc, cols), bytes_moved=full)
timed(f"pack, all {K} pieces", lambda: [be.pack(x, y[k * piece:(k + 1) * piece], P, rows_loc, cbk, cols, k * cbk, None)
                                      for k in range(K)], bytes_moved=full)
timed(f"pack + twiddle, all {K} pieces",
      lambda: [be.pack(x, y[k * piece:(k + 1) * piece], P, rows_loc, cbk, cols, k * cbk, (True, 4096, 1 << 30))
               for k in range(K)], bytes_moved=full)
timed(f"swap_leading, all {K} pieces",
      lambda: [be.swap_leading(x[k * piece:(k + 1) * piece], y[k * piece:(k + 1) * piece], P, cbk, rows_loc)
               for k in range(K)], bytes_moved=full)
timed("twiddle_rows, whole local matrix", lambda: be.twiddle_rows(y, cb, n1, 4096, 1 << 30, True), bytes_moved=full)
timed(f"fft_rows 2^15, all {K} pieces", lambda: [be.fft_rows(y[k * piece:(k + 1) * piece], n1, True) for k in range(K)],
      bytes_moved=full)
timed("fft_rows 2^15, whole local matrix (4096 rows)", lambda: be.fft_rows(y, n1, True), bytes_moved=full)
timed("device copy of the local block", lambda: y.copy_(x), bytes_moved=full)
