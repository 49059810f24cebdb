"""fourier_b200 -- Python mirror of the `fourier` crate's public surface over libfourier.so.

The product is the C-ABI library (include/fourier.h, include/fourier_b200.h; hand-written sm_100a
CUDA).  This module is the thin host-side binding used by the tests and the benchmark; it mirrors the
reference's operator interface name for name:

    reference (Rust)                                      here
    fourier::Transform {Fft, Ifft, UnscaledIfft,          Transform (IntEnum, same C codes,
      SqrtScaledFft, SqrtScaledIfft}  fft.rs:5-16           is_forward / inverse: fft.rs:20-36)
    fourier::Fft trait  fft.rs:40-82                      class Fft: size, transform_in_place,
                                                            transform, fft_in_place, ifft_in_place, fft, ifft
    fourier::create_fft_f32 / create_fft_f64              create_fft_f32 / create_fft_f64
      fourier/src/lib.rs:31-60

Buffers are numpy arrays (host memory: staged over PCIe by the library) or torch CUDA tensors
(device memory: transformed in HBM on torch's current stream).  A leading batch dimension is
allowed: shape (..., size) means prod(...) independent transforms.  There is no CPU fallback.
"""
import ctypes
import enum

import numpy as np

from . import _lib

__all__ = ["Transform", "Fft", "create_fft_f32", "create_fft_f64", "set_device", "fill_input", "lib_path"]


class Transform(enum.IntEnum):
    """fourier-algorithms/src/fft.rs:5-16; integer values are the C codes of fourier.h:30-36."""
    Fft = 0
    Ifft = 1
    UnscaledIfft = 2
    SqrtScaledFft = 3
    SqrtScaledIfft = 4

    def is_forward(self):
        return self in (Transform.Fft, Transform.SqrtScaledFft)

    def inverse(self):
        return {Transform.Fft: Transform.Ifft, Transform.Ifft: Transform.Fft,
                Transform.SqrtScaledFft: Transform.SqrtScaledIfft,
                Transform.SqrtScaledIfft: Transform.SqrtScaledFft,
                Transform.UnscaledIfft: None}[self]


def _is_torch(x):
    return type(x).__module__.startswith("torch")


class Fft:
    """A plan for one transform size: the `Box<dyn Fft<Real = T> + Send>` of the reference."""

    def __init__(self, size, real, general=False):
        if real not in ("f32", "f64"):
            raise ValueError("real must be 'f32' or 'f64'")
        self.real = real
        self._t = "float" if real == "f32" else "double"
        self._np_dtype = np.dtype(np.complex64 if real == "f32" else np.complex128)
        L = _lib.load()
        ctor = getattr(L, f"fourier_b200_create_general_{self._t}" if general else f"fourier_create_{self._t}")
        self._plan = ctor(int(size))
        if not self._plan:
            raise RuntimeError(f"fourier_create_{self._t}({size}) returned NULL: {_lib.last_error()}")
        self._size = int(size)
        self._device = self.info()["device"]

    # -- lifetime -------------------------------------------------------------------------------
    def close(self):
        p, self._plan = getattr(self, "_plan", None), None
        if p:
            getattr(_lib.load(), f"fourier_destroy_{self._t}")(p)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- Fft trait --------------------------------------------------------------------------------
    def size(self):
        return self._size

    def info(self):
        i = _lib.PlanInfo()
        getattr(_lib.load(), f"fourier_b200_plan_info_{self._t}")(self._plan, ctypes.byref(i))
        d = {f: getattr(i, f) for f, _ in i._fields_}
        d["path_name"] = _lib.load().fourier_b200_path_name(i.path).decode()
        d["inner_path_name"] = _lib.load().fourier_b200_path_name(i.inner_path).decode()
        return d

    def kernel_name(self):
        """The kernel that moves (nearly) all of this plan's bytes, as a profiler lists it."""
        return getattr(_lib.load(), f"fourier_b200_plan_kernel_{self._t}")(self._plan).decode()

    def _describe(self, x):
        """-> (pointer, batch, is_device, stream)"""
        if _is_torch(x):
            import torch
            want = torch.complex64 if self.real == "f32" else torch.complex128
            if x.dtype != want:
                raise TypeError(f"expected {want}, got {x.dtype}")
            if not x.is_contiguous():
                raise ValueError("buffers must be contiguous")
            if x.shape[-1] != self._size:
                raise ValueError(f"last dimension must be {self._size}")  # assert_eq!, fft.rs:57-58
            batch = x.numel() // self._size
            if x.is_cuda:
                if x.device.index != self._device:
                    raise ValueError(f"plan lives on cuda:{self._device}, buffer on {x.device}")
                return x.data_ptr(), batch, True, torch.cuda.current_stream(x.device).cuda_stream
            return x.data_ptr(), batch, False, None
        if not isinstance(x, np.ndarray):
            raise TypeError("buffers must be numpy arrays or torch tensors")
        if x.dtype != self._np_dtype:
            raise TypeError(f"expected {self._np_dtype}, got {x.dtype}")
        if not x.flags.c_contiguous:
            raise ValueError("buffers must be C-contiguous")
        if x.shape[-1] != self._size:
            raise ValueError(f"last dimension must be {self._size}")
        return x.ctypes.data, x.size // self._size, False, None

    def _run(self, src, dst, transform):
        code = int(Transform(transform))
        pi, bi, di, si = self._describe(src)
        po, bo, do_, so = self._describe(dst)
        if bi != bo:
            raise ValueError("input and output hold a different number of transforms")
        if di != do_:
            raise ValueError("input and output must both be host or both be device buffers")
        L = _lib.load()
        if di:
            rc = getattr(L, f"fourier_b200_transform_batch_async_{self._t}")(self._plan, pi, po, bi, code, si)
        else:
            rc = getattr(L, f"fourier_b200_transform_batch_{self._t}")(self._plan, pi, po, bi, code)
        if rc != 0:
            raise RuntimeError(f"transform failed (cuda error {rc}): {_lib.last_error()}")

    def transform_in_place(self, input, transform):
        """Fft::transform_in_place (fft.rs:48)."""
        self._run(input, input, transform)

    def transform(self, input, output, transform):
        """Fft::transform (fft.rs:51-61): out of place, input untouched."""
        self._run(input, output, transform)

    def fft_in_place(self, input):
        self.transform_in_place(input, Transform.Fft)

    def ifft_in_place(self, input):
        self.transform_in_place(input, Transform.Ifft)

    def fft(self, input, output):
        self.transform(input, output, Transform.Fft)

    def ifft(self, input, output):
        self.transform(input, output, Transform.Ifft)

    # -- distributed six-step transform: row FFTs + exchange in one pass (csrc/dist_fft.cu) -----------------
    def fft_rows_exchange(self, rows, outs, out_ld, out_off, forward=True, twiddle=None):
        """FFT of the contiguous rows of the CUDA tensor `rows` (unscaled) whose last register stage stores the result
        transposed into the buffers `outs` (a ctypes array of device pointers, one per destination rank):
        outs[q][c*out_ld + out_off + r] = FFT(rows[r])[q*cb + c] * w_Ntot^{(row0 + r)*(q*cb + c)}, cb = size / len(outs);
        twiddle = None or (row0, Ntot).  Two-pass plans only (raises NotImplementedError otherwise)."""
        ptr, count, on_device, stream = self._describe(rows)
        if not on_device:
            raise ValueError("fft_rows_exchange needs CUDA tensors")
        mode, row0, n_total = (0, 0, 0) if twiddle is None else ((1 if forward else 2), int(twiddle[0]), int(twiddle[1]))
        rc = getattr(_lib.load(), f"fourier_b200_fft_rows_exchange_{self._t}")(
            self._plan, ptr, count, int(bool(forward)), outs, len(outs), int(out_ld), int(out_off), mode, row0,
            n_total, stream)
        if rc == 801:
            raise NotImplementedError(_lib.last_error())
        if rc != 0:
            raise RuntimeError(f"fft_rows_exchange failed (cuda error {rc}): {_lib.last_error()}")

    # -- the raw single-transform reference ABI (for the boundary tests) ------------------------------
    def c_transform(self, input, output, transform):
        pi, _, _, _ = self._describe(input)
        po, _, _, _ = self._describe(output)
        getattr(_lib.load(), f"fourier_transform_{self._t}")(self._plan, pi, po, int(transform))

    def c_transform_in_place(self, input, transform):
        pi, _, _, _ = self._describe(input)
        getattr(_lib.load(), f"fourier_transform_in_place_{self._t}")(self._plan, pi, int(transform))


def create_fft_f32(size, general=False):
    """fourier::create_fft_f32 (fourier/src/lib.rs:31-43)."""
    return Fft(size, "f32", general=general)


def create_fft_f64(size, general=False):
    """fourier::create_fft_f64 (fourier/src/lib.rs:49-60)."""
    return Fft(size, "f64", general=general)


def set_device(device):
    rc = _lib.load().fourier_b200_set_device(int(device))
    if rc != 0:
        raise RuntimeError(f"fourier_b200_set_device({device}) failed: {_lib.last_error()}")


def fill_input(tensor, first_transform=0, seed=0xDEADBEEF):
    """Fill a torch CUDA complex tensor of shape (..., size) with the synthetic benchmark input
    (same counter-hash stream as the oracle's fill_input), on torch's current stream."""
    import torch
    if not tensor.is_cuda or not tensor.is_contiguous():
        raise ValueError("fill_input needs a contiguous CUDA tensor")
    t = {torch.complex64: "float", torch.complex128: "double"}[tensor.dtype]
    n = tensor.shape[-1]
    count = 2 * tensor.numel()
    stream = torch.cuda.current_stream(tensor.device).cuda_stream
    rc = getattr(_lib.load(), f"fourier_b200_fill_input_{t}")(
        tensor.data_ptr(), 2 * n * int(first_transform), count, seed, stream)
    if rc != 0:
        raise RuntimeError(f"fill_input failed: {_lib.last_error()}")
    return tensor


def lib_path():
    return _lib.LIB_PATH
