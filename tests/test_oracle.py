"""Pins the CPU oracle (oracle/) against every golden vector / known-answer test the reference's own
tests hold for this path, plus an independent f64 pocketfft check.  CPU only."""
import numpy as np
import pytest

from oracle import oracle as O
from helpers import assert_near_reference_rule, golden_10pt, rel_err, sweep_input, truth_f64

DTYPES = [np.complex64, np.complex128]


@pytest.mark.parametrize("dtype", DTYPES)
def test_golden_10pt_naive_dft(dtype):
    # integrity.rs:42-87 verbatim: the naive DFT/IDFT must reproduce the golden pair
    x, y = golden_10pt()
    assert_near_reference_rule(O.naive_dft(x.astype(dtype)), y)
    assert_near_reference_rule(O.naive_dft(y.astype(dtype), inverse=True), x)


@pytest.mark.parametrize("dtype", DTYPES)
def test_golden_10pt_fft(dtype):
    # N=10 is not {2,3}-smooth -> exercises the Bluestein path of the oracle on the golden pair
    x, y = golden_10pt()
    p = O.Plan(10, dtype)
    assert p.is_bluestein and p.inner_size == 32
    assert_near_reference_rule(p.transform(x.astype(dtype), O.FFT), y)
    assert_near_reference_rule(p.transform(y.astype(dtype), O.IFFT), x)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("forward", [True, False])
def test_sweep_1_to_255(dtype, forward):
    # integrity.rs:145-192: every size 1..=255 vs the naive DFT, reference tolerances
    data = sweep_input(256, dtype, forward)
    code = O.FFT if forward else O.IFFT
    for size in range(1, 256):
        x = data[:size]
        got = O.transform(x, code)
        want = O.naive_dft(x, inverse=not forward)
        assert_near_reference_rule(got, want)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("size", [64, 73, 128])
def test_static_sizes(dtype, size):
    # integrity.rs:234-254 (static 64 Autosort / 73 Bluestein), fourier-macros doc-test size 128
    for forward in (True, False):
        x = sweep_input(256, dtype, forward)[:size]
        got = O.transform(x, O.FFT if forward else O.IFFT)
        assert_near_reference_rule(got, O.naive_dft(x, inverse=not forward))
    p = O.Plan(size, dtype)
    assert p.is_bluestein == (size == 73)


@pytest.mark.parametrize("dtype", DTYPES)
def test_ffi_impulse_round_trip(dtype):
    # fourier-ffi/test.c:8-21: {1,0,0,0} -> FFT -> IFFT in place returns the input within 1e-10
    x = np.array([1, 0, 0, 0], dtype=dtype)
    p = O.Plan(4, dtype)
    y = p.transform(x, O.FFT)
    assert np.allclose(y, np.ones(4))
    z = p.transform(y, O.IFFT)
    assert np.abs(z - x).max() <= 1e-10


def test_plan_selection_and_factorisation():
    # fourier/src/lib.rs:38-42 + autosort/mod.rs:104-134
    cases = {1024: [1, 2, 1, 0, 0], 1 << 16: [1, 4, 1, 0, 0], 1 << 20: [1, 6, 0, 0, 0], 2048: [1, 3, 0, 0, 0],
             2: [0, 0, 0, 0, 1], 3: [0, 0, 0, 1, 0], 6: [0, 0, 0, 1, 1], 12: [1, 0, 0, 1, 0], 1: [0, 0, 0, 0, 0]}
    for n, counts in cases.items():
        p = O.Plan(n, np.complex64)
        assert not p.is_bluestein and p.counts == counts, (n, p.counts)
    assert len(O.Plan(1 << 20, np.complex64).twiddles()) == 1348168
    assert len(O.Plan(1 << 16, np.complex128).twiddles()) == 84260
    assert len(O.Plan(1024, np.complex64).twiddles()) == 1316
    for n in (5, 7, 10, 73, 191, 1009):
        p = O.Plan(n, np.complex64)
        assert p.is_bluestein
        m = p.inner_size
        assert m >= 2 * n - 1 and m & (m - 1) == 0 and m // 2 < 2 * n - 1
    assert O.Plan(1009, np.complex64).inner_size == 2048
    with pytest.raises(ValueError):
        O.Plan(0, np.complex64)  # the reference hangs on 0 (mod.rs:112); the oracle refuses


def test_twiddle_table_layout():
    # autosort/mod.rs:24-46: stage of size S, radix R: row i = [1, w_S^i, ..., w_S^{(R-1)i}]
    p = O.Plan(1024, np.complex128)
    tw = p.twiddles(True)
    s, pos = 1024, 0
    for radix in (4, 8, 8, 4):
        m = s // radix
        i = np.arange(m)[:, None]
        j = np.arange(radix)[None, :]
        want = np.exp(-2j * np.pi * (i * j) / s).reshape(-1)
        assert np.abs(tw[pos:pos + s] - want).max() < 1e-15
        pos += s
        s //= radix
    assert pos == len(tw)
    assert np.array_equal(p.twiddles(False), np.conj(tw))


@pytest.mark.parametrize("dtype,tol", [(np.complex64, 1e-5), (np.complex128, 1e-12)])
@pytest.mark.parametrize("code", [0, 1, 2, 3, 4])
def test_all_transform_codes_vs_pocketfft(dtype, tol, code):
    # the three Transform variants the reference never tests, checked against their definitions
    rng = np.random.default_rng(code)
    for n in (1, 2, 3, 4, 6, 8, 9, 16, 27, 30, 96, 100, 243, 256, 1009, 1024, 2048, 4096):
        x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(dtype)
        assert rel_err(O.transform(x, code), truth_f64(x, code)) < tol, (n, code)


@pytest.mark.parametrize("dtype,tol,n", [(np.complex64, 1e-5, 1 << 20), (np.complex128, 1e-12, 1 << 16)])
def test_baseline_sizes_vs_pocketfft(dtype, tol, n):
    x = O.fill_input(2, n, dtype)
    assert rel_err(O.transform(x, O.FFT), truth_f64(x, O.FFT)) < tol


@pytest.mark.parametrize("dtype", DTYPES)
def test_input_generator_c_equals_numpy(dtype):
    a = O.fill_input(3, 1009, dtype, first_transform=7)
    b = O.fill_input_numpy(3, 1009, dtype, first_transform=7)
    assert np.array_equal(a.view(np.uint8), b.view(np.uint8))
    assert a.real.min() >= -1 and a.real.max() < 1 and abs(a.real.mean()) < 0.05


def test_batch_driver_matches_single():
    x = O.fill_input(5, 96, np.complex64)
    out, sec = O.transform_batch(x, O.FFT, threads=2)
    assert sec >= 0
    assert np.array_equal(out, O.transform(x, O.FFT))
