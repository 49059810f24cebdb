"""Runs one batched transform with FOURIER_B200_TRACE set (phase timestamps of CTA 0 of the fused kernel).
usage: PYTHONPATH=. FOURIER_B200_TRACE=out.txt python tools/trace_run.py [f32|f64] [log2 N] [batch]"""
import sys
import torch
import fourier_b200 as fb

real = sys.argv[1] if len(sys.argv) > 1 else "f32"
k = int(sys.argv[2]) if len(sys.argv) > 2 else (20 if real == "f32" else 16)
batch = int(sys.argv[3]) if len(sys.argv) > 3 else (256 if real == "f32" else 4096)
p = fb.create_fft_f32(1 << k) if real == "f32" else fb.create_fft_f64(1 << k)
dt = torch.complex64 if real == "f32" else torch.complex128
x = torch.empty((batch, 1 << k), dtype=dt, device="cuda")
fb.fill_input(x)
y = torch.empty_like(x)
p.transform(x, y, fb.Transform.Fft)
torch.cuda.synchronize()
print(p.info())
