"""Shared test helpers: the reference test-suite's comparison rule and input procedure."""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def golden_10pt():
    with open(os.path.join(HERE, "golden", "integrity_10pt.json")) as f:
        g = json.load(f)
    x = np.array([complex(a, b) for a, b in g["x"]], dtype=np.complex128)
    y = np.array([complex(a, b) for a, b in g["y"]], dtype=np.complex128)
    return x, y


def _ulp_diff(a, b):
    """|a-b| in units in the last place, float_cmp style (ordered-integer distance)."""
    if a.dtype == np.float32:
        ia, ib = a.view(np.int32).astype(np.int64), b.view(np.int32).astype(np.int64)
        ia = np.where(ia < 0, np.int64(-(2**31)) - ia, ia)
        ib = np.where(ib < 0, np.int64(-(2**31)) - ib, ib)
        return np.abs(ia - ib)
    ia, ib = a.view(np.int64), b.view(np.int64)
    fa = np.where(ia < 0, -(ia.astype(np.float64) + 2.0**63), ia.astype(np.float64))
    fb = np.where(ib < 0, -(ib.astype(np.float64) + 2.0**63), ib.astype(np.float64))
    return np.abs(fa - fb)


def assert_near_reference_rule(actual, expected):
    """fourier/tests/integrity.rs:89-143: per component approx_eq with epsilon 1e-4 | 8 ulps (f32),
    1e-11 | 8 ulps (f64)."""
    actual = np.ascontiguousarray(actual)
    expected = np.ascontiguousarray(expected).astype(actual.dtype)
    real = np.float32 if actual.dtype == np.complex64 else np.float64
    eps = 1e-4 if real == np.float32 else 1e-11
    a = actual.view(real)
    e = expected.view(real)
    ok = (np.abs(a.astype(np.float64) - e.astype(np.float64)) <= eps) | (_ulp_diff(a, e) <= 8)
    assert ok.all(), f"{(~ok).sum()} components differ beyond {eps}|8ulp; worst {np.abs(a - e).max()}"


def rel_err(got, ref):
    """North-star parity metric: max|got-ref| / max|ref|."""
    got = np.asarray(got).astype(np.complex128)
    ref = np.asarray(ref).astype(np.complex128)
    return float(np.abs(got - ref).max() / max(np.abs(ref).max(), 1e-300))


def sweep_input(max_size, dtype, forward, seed=0xDEADBEEF):
    """Procedure of integrity.rs:152-165: N(0, sigma) components, sigma = 1 forward, MAX_SIZE inverse.
    (The reference's RNG streams are ChaCha + thread_rng and half non-deterministic; the procedure is
    reproduced, not the bits.)"""
    rng = np.random.default_rng(seed)
    sigma = 1.0 if forward else float(max_size)
    z = rng.normal(0.0, sigma, max_size) + 1j * rng.normal(0.0, sigma, max_size)
    return z.astype(dtype)


def truth_f64(x, code):
    """Independent reference: numpy pocketfft in f64, with the Transform scaling definitions
    (fourier-algorithms/src/fft.rs:5-16, autosort/mod.rs:381-385)."""
    x = np.asarray(x).astype(np.complex128)
    n = x.shape[-1]
    if code in (0, 3):
        y = np.fft.fft(x, axis=-1)
    else:
        y = np.fft.ifft(x, axis=-1) * n
    scale = {0: 1.0, 1: 1.0 / n, 2: 1.0, 3: 1.0 / np.sqrt(n), 4: 1.0 / np.sqrt(n)}[code]
    return y * scale
