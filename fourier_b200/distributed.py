"""One huge 1-D complex FFT distributed over the GPUs of a box (BASELINE.json configs[4]: N = 2^30 over 8
B200s): the six-step algorithm with all-to-all transposes between local batched FFTs.

The reference has no counterpart (single-threaded CPU code); this is the only place of the
engine where the path has a real exchange step, hence the only place a collective is used.

Layout.  N = N1*N2.  The input is block-distributed in natural order: rank r of P holds samples
[r*N/P, (r+1)*N/P), i.e. rows n1 in [r*N1/P, (r+1)*N1/P) of x[n1][n2] (n = n1*N2 + n2).  The output has the
same block distribution of natural order X[k], k = k1 + N1*k2: rank r holds rows k2 of X[k2][k1].

    1. exchange:  [n1_loc][n2]  ->  [n2_loc][n1]
    2. N2/P local FFTs of length N1 over n1, then  *= w_N^{n2*k1}
    3. exchange:  [n2_loc][k1]  ->  [k1_loc][n2]
    4. N1/P local FFTs of length N2 over n2
    5. exchange:  [k1_loc][k2]  ->  [k2_loc][k1]      (natural order)

One exchange = the transpose of a row-distributed matrix: local transpose (the block for rank q becomes
contiguous), one `all_to_all_single` of N/P samples per rank (N/P * (P-1)/P cross NVLink), local swap of
the two leading axes of the received [source rank][my rows][their rows].  The local pieces are this
library's kernels (batched FFT plans, fourier_b200_transpose_*, _swap_leading_*, _twiddle_rows_*);
torch provides memory, streams and the collective.

`backend` abstracts the local compute so that the exchange logic is tested on CPU with gloo + numpy
(tests/test_distributed_host_logic.py); the product backend is `CudaBackend` (no CPU fallback).
"""
import numpy as np

from . import _lib


class CudaBackend:
    """Local compute on the rank's GPU through the C ABI."""

    def __init__(self, real):
        import torch
        self.torch = torch
        self.real = real
        self.t = "float" if real == "f32" else "double"
        self.dtype = torch.complex64 if real == "f32" else torch.complex128
        self._plans = {}

    def _stream(self, x):
        return self.torch.cuda.current_stream(x.device).cuda_stream

    def _call(self, name, *args):
        rc = getattr(_lib.load(), f"fourier_b200_{name}_{self.t}")(*args)
        if rc != 0:
            raise RuntimeError(f"fourier_b200_{name}_{self.t} failed: {_lib.last_error()}")

    def fft_rows(self, x, n, forward):
        """In-place batched FFT over rows of length n (unscaled in both directions)."""
        from . import Fft, Transform
        plan = self._plans.get(n)
        if plan is None:
            plan = self._plans[n] = Fft(n, self.real)
        plan.transform_in_place(x.view(-1, n), Transform.Fft if forward else Transform.UnscaledIfft)

    def transpose(self, src, dst, rows, cols):
        self._call("transpose", src.data_ptr(), dst.data_ptr(), 1, rows, cols, self._stream(src))

    def swap_leading(self, src, dst, a, b, inner):
        self._call("swap_leading", src.data_ptr(), dst.data_ptr(), a, b, inner, self._stream(src))

    def twiddle_rows(self, x, rows, cols, row0, n_total, forward):
        self._call("twiddle_rows", x.data_ptr(), rows, cols, row0, n_total, int(forward), self._stream(x))

    def all_to_all(self, dst, src, group):
        import torch.distributed as dist
        # NCCL has no complex dtype: exchange the raw (re, im) scalars
        dist.all_to_all_single(self.torch.view_as_real(dst), self.torch.view_as_real(src), group=group)


class NumpyBackend:
    """CPU stand-in used ONLY by the host-logic tests (gloo): the same interface on torch CPU tensors."""

    def __init__(self):
        import torch
        self.torch = torch

    def fft_rows(self, x, n, forward):
        a = x.view(-1, n).numpy()
        a[...] = np.fft.fft(a, axis=-1) if forward else np.fft.ifft(a, axis=-1) * n

    def transpose(self, src, dst, rows, cols):
        dst.view(cols, rows).copy_(src.view(rows, cols).t())

    def swap_leading(self, src, dst, a, b, inner):
        dst.view(b, a, inner).copy_(src.view(a, b, inner).transpose(0, 1))

    def twiddle_rows(self, x, rows, cols, row0, n_total, forward):
        idx = (np.arange(rows, dtype=np.int64)[:, None] + row0) * np.arange(cols, dtype=np.int64)[None, :] % n_total
        w = np.exp((-2j if forward else 2j) * np.pi * idx / n_total)
        a = x.view(rows, cols).numpy()
        a *= w.astype(a.dtype)

    def all_to_all(self, dst, src, group):
        import torch.distributed as dist
        dist.all_to_all_single(self.torch.view_as_real(dst), self.torch.view_as_real(src), group=group)


class DistributedFft:
    """Plan for one length-N transform over `world` ranks (N = n1 * n2, both divisible by world)."""

    def __init__(self, n1, n2, rank, world, backend, group=None):
        if n1 % world or n2 % world:
            raise ValueError("n1 and n2 must be divisible by the number of ranks")
        self.n1, self.n2, self.n = n1, n2, n1 * n2
        self.rank, self.world, self.backend, self.group = rank, world, backend, group

    def local_samples(self):
        return self.n // self.world

    def wire_bytes_per_exchange(self, itemsize):
        """Bytes each rank sends over NVLink per exchange."""
        return self.local_samples() * itemsize * (self.world - 1) // self.world

    def _exchange(self, src, dst, rows_loc, cols):
        """Transpose of the row-distributed global matrix [rows_loc * P][cols]: on return `dst` holds this
        rank's [cols / P] rows of the transposed matrix, each rows_loc * P long.  `src` is clobbered."""
        P, be = self.world, self.backend
        cb = cols // P
        be.transpose(src, dst, rows_loc, cols)              # dst = [cols][rows_loc] = [P][cb][rows_loc]
        if P == 1:
            return dst
        be.all_to_all(src, dst, self.group)                 # src = [P (source rank)][cb][rows_loc]
        be.swap_leading(src, dst, P, cb, rows_loc)          # dst = [cb][P][rows_loc] = [cb][rows_loc * P]
        return dst

    def transform(self, x, scratch, forward=True):
        """x: this rank's N/P samples (its block of the natural order), scratch: same size.  Both are
        clobbered; returns the one holding this rank's block of the result (unscaled in both directions:
        Transform.Fft / Transform.UnscaledIfft)."""
        P, be = self.world, self.backend
        n1, n2 = self.n1, self.n2
        r1, r2 = n1 // P, n2 // P
        a = self._exchange(x, scratch, r1, n2)              # [n2_loc][n1]
        b = x if a is scratch else scratch
        be.fft_rows(a, n1, forward)
        be.twiddle_rows(a, r2, n1, self.rank * r2, self.n, forward)
        b = self._exchange(a, b, r2, n1)                    # [k1_loc][n2]
        a = x if b is scratch else scratch
        be.fft_rows(b, n2, forward)
        return self._exchange(b, a, r1, n2)                 # [k2_loc][k1]
