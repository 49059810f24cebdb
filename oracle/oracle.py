"""ctypes binding of the CPU oracle (oracle/libfourier_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the cpu_baseline /
``--impl reference`` legs of bench.py.  Nothing under fourier_b200/ may import this module.
The algorithm it wraps is the plain-C restatement of the reference (see fourier_oracle.h).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libfourier_oracle.so")
_AVX2_PATH = os.path.join(_HERE, "libfourier_oracle_avx2.so")   # -O3 -mavx2 build, used for TIMING only

FFT, IFFT, UNSCALED_IFFT, SQRT_SCALED_FFT, SQRT_SCALED_IFFT = range(5)
SEED = 0xDEADBEEF  # echoes fourier/tests/integrity.rs:159


def build(force=False):
    """Compile the C restatement with the committed Makefile (gcc, -ffp-contract=off)."""
    src = [os.path.join(_HERE, f) for f in ("fourier_oracle.c", "fourier_oracle_impl.inc", "fourier_oracle.h")]
    newest = max(os.path.getmtime(s) for s in src)
    if not force and all(os.path.exists(p) and os.path.getmtime(p) >= newest for p in (_LIB_PATH, _AVX2_PATH)):
        return _LIB_PATH
    subprocess.run(["make", "-C", _HERE, "-B"], check=True, capture_output=True)
    return _LIB_PATH


_lib = None
_timing_lib = None


def _cpu_has_avx2():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    return " avx2" in line
    except OSError:
        pass
    return False


def timing_build():
    """Which build the timing legs run: 'O3 avx2' when the host CPU has AVX2, else the plain 'O2 scalar' one."""
    return "gcc -O3 -mavx2 -ffp-contract=off" if _cpu_has_avx2() else "gcc -O2 -ffp-contract=off (no AVX2 on this host)"


def _bind(path):
    if True:
        L = ctypes.CDLL(path)
        for sfx, real in (("f32", ctypes.c_float), ("f64", ctypes.c_double)):
            rp = ctypes.POINTER(real)
            g = lambda name: getattr(L, f"fo_{name}_{sfx}")
            g("create").restype = ctypes.c_void_p
            g("create").argtypes = [ctypes.c_size_t]
            g("destroy").argtypes = [ctypes.c_void_p]
            g("size").restype = ctypes.c_size_t
            g("size").argtypes = [ctypes.c_void_p]
            g("is_bluestein").argtypes = [ctypes.c_void_p]
            g("inner_size").restype = ctypes.c_size_t
            g("inner_size").argtypes = [ctypes.c_void_p]
            g("counts").argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_size_t)]
            g("num_twiddles").restype = ctypes.c_size_t
            g("num_twiddles").argtypes = [ctypes.c_void_p]
            g("copy_twiddles").argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
            g("transform_in_place").argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
            g("transform").argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
            g("naive_dft").argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
            g("fill_input").argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_size_t, ctypes.c_uint64]
            g("transform_batch").argtypes = [ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t,
                                             ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_double)]
        L.fo_hash64.restype = ctypes.c_uint64
        L.fo_hash64.argtypes = [ctypes.c_uint64, ctypes.c_uint64]
    return L


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = _bind(_LIB_PATH)
    return _lib


def timing_lib():
    """The AVX2 -O3 build of the same source (bench.py's CPU arm); falls back to lib() on a host without AVX2."""
    global _timing_lib
    if _timing_lib is None:
        build()
        _timing_lib = _bind(_AVX2_PATH) if _cpu_has_avx2() else lib()
    return _timing_lib


def _sfx(dtype):
    dtype = np.dtype(dtype)
    if dtype == np.complex64:
        return "f32"
    if dtype == np.complex128:
        return "f64"
    raise TypeError(f"oracle handles complex64/complex128, got {dtype}")


class Plan:
    """Mirror of `Box<dyn Fft>` as built by create_fft_f32/f64 (fourier/src/lib.rs:31-60)."""

    def __init__(self, size, dtype):
        self.dtype = np.dtype(dtype)
        self.sfx = _sfx(dtype)
        self._f = lambda name: getattr(lib(), f"fo_{name}_{self.sfx}")
        self._p = self._f("create")(size)
        if not self._p:
            raise ValueError(f"oracle cannot plan size {size}")
        self.size = size

    def close(self):
        if getattr(self, "_p", None):
            self._f("destroy")(self._p)
            self._p = None

    __del__ = close

    @property
    def is_bluestein(self):
        return bool(self._f("is_bluestein")(self._p))

    @property
    def inner_size(self):
        return int(self._f("inner_size")(self._p))

    @property
    def counts(self):
        out = (ctypes.c_size_t * 5)()
        self._f("counts")(self._p, out)
        return list(out)

    def twiddles(self, forward=True):
        n = int(self._f("num_twiddles")(self._p))
        out = np.empty(n, dtype=self.dtype)
        self._f("copy_twiddles")(self._p, int(forward), out.ctypes.data)
        return out

    def transform(self, x, transform=FFT):
        x = np.ascontiguousarray(x, dtype=self.dtype)
        assert x.shape[-1] == self.size
        flat = x.reshape(-1, self.size)
        out = np.empty_like(flat)
        for b in range(flat.shape[0]):
            self._f("transform")(self._p, flat[b].ctypes.data, out[b].ctypes.data, int(transform))
        return out.reshape(x.shape)


def transform(x, transform=FFT):
    """One-shot: plan for x.shape[-1] in x.dtype, transform every row."""
    x = np.asarray(x)
    p = Plan(x.shape[-1], x.dtype)
    try:
        return p.transform(x, transform)
    finally:
        p.close()


def transform_batch(x, transform=FFT, threads=1, timing=False):
    """Multi-threaded batch (one plan per thread). Returns (out, seconds of the transform loops).
    timing=True runs the AVX2 -O3 build (timing_lib) instead of the plain one the parity tests use."""
    x = np.ascontiguousarray(x)
    n = x.shape[-1]
    batch = x.size // n
    out = np.empty_like(x)
    sec = ctypes.c_double(0.0)
    rc = getattr(timing_lib() if timing else lib(), f"fo_transform_batch_{_sfx(x.dtype)}")(
        n, x.ctypes.data, out.ctypes.data, batch, int(transform), int(threads), ctypes.byref(sec))
    if rc != 0:
        raise RuntimeError("oracle batch transform failed")
    return out, sec.value


def naive_dft(x, inverse=False):
    """The reference TEST oracle (fourier/tests/integrity.rs:6-40), accumulating in x.dtype."""
    x = np.ascontiguousarray(x)
    out = np.empty_like(x)
    getattr(lib(), f"fo_naive_dft_{_sfx(x.dtype)}")(x.ctypes.data, out.ctypes.data, x.shape[-1], int(inverse))
    return out


def fill_input(batch, n, dtype, first_transform=0, seed=SEED):
    """Synthetic input rows [first_transform, first_transform+batch): counter-hash U[-1,1)."""
    dtype = np.dtype(dtype)
    out = np.empty((batch, n), dtype=dtype)
    getattr(lib(), f"fo_fill_input_{_sfx(dtype)}")(
        out.ctypes.data, 2 * n * first_transform, 2 * n * batch, seed)
    return out


def fill_input_numpy(batch, n, dtype, first_transform=0, seed=SEED):
    """Same generator in pure numpy (cross-checks the C and CUDA implementations)."""
    dtype = np.dtype(dtype)
    idx = np.arange(2 * n * batch, dtype=np.uint64) + np.uint64(2 * n * first_transform)
    with np.errstate(over="ignore"):
        z = np.uint64(seed) + (idx + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    if dtype == np.complex64:
        u = (z >> np.uint64(40)).astype(np.float32) * np.float32(1.0 / 8388608.0) - np.float32(1.0)
    else:
        u = (z >> np.uint64(11)).astype(np.float64) * (1.0 / 4503599627370496.0) - 1.0
    return u.view(dtype).reshape(batch, n)
