/*
 * fourier_oracle.h -- CPU oracle for the batched 1-D complex FFT hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / `--impl reference` legs of bench.py may load it.  The product library
 * (libfourier.so, include/fourier.h) never links, calls or falls back to anything in oracle/.
 *
 * What it is: a plain-C restatement of the reference's CPU algorithm (calebzulawski/fourier @
 * dc49696): Stockham autosort with radices [4,8,4,3,2], the reference's twiddle-table layout
 * and f64->T rounding, its butterfly operation order, its scale epilogue, its Bluestein wrapper
 * and its plan-selection rule.  The functions are in fourier_oracle_impl.inc, each citing the
 * reference file:line it follows.
 *
 * Pinning status.  The reference is Rust and neither rustc nor cargo exist in this image, so
 * the reference itself cannot be run here to produce bit-exact fixtures; its only third-party
 * arithmetic is num-complex 0.2 `Complex<T>` + - * (textbook (ac-bd, ad+bc), restated in
 * fo_mul_*).  The oracle is pinned against every golden vector / known-answer test the
 * reference's own tests hold for this path (tests/test_oracle.py):
 *   - the 10-point x -> y golden pair, fourier/tests/integrity.rs:48-72 (forward, and y -> x
 *     inverse with 1/N), in f32 and f64;
 *   - the sweep procedure of integrity.rs:145-192: every size 1..=255, f32/f64, Fft and Ifft,
 *     against the reference test's own naive DFT (restated as fo_naive_dft_*), with the
 *     reference tolerances (1e-4 | 8 ulp for f32, 1e-11 | 8 ulp for f64, integrity.rs:89-143);
 *   - static sizes 64 and 73 (integrity.rs:234-254), doc-test size 128;
 *   - the 4-point impulse FFT -> IFFT round trip of fourier-ffi/test.c:8-21.
 * The reference holds NO bit-exact FFT output vectors (its tests are tolerance based), so
 * parity with the reference is tolerance-pinned, not bit-pinned.  A second, independent check
 * (numpy pocketfft in f64) bounds oracle-vs-truth in the same tests.
 */
#ifndef FOURIER_ORACLE_H_
#define FOURIER_ORACLE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* transform codes: fourier-ffi/src/lib.rs:3-12 == fourier-ffi/include/fourier.h:30-36 */
enum { FO_FFT = 0, FO_IFFT = 1, FO_UNSCALED_IFFT = 2, FO_SQRT_SCALED_FFT = 3, FO_SQRT_SCALED_IFFT = 4 };

/* Transform::is_forward, fourier-algorithms/src/fft.rs:20-25 */
int fo_is_forward(int transform);

/* splitmix64 finaliser of (seed, counter): the synthetic-input generator shared with the GPU */
uint64_t fo_hash64(uint64_t seed, uint64_t counter);

struct fo_plan_f32;
struct fo_plan_f64;

#define FO_DECLARE(SFX, REAL)                                                                      \
  struct fo_plan_##SFX *fo_create_##SFX(size_t size);                                              \
  void fo_destroy_##SFX(struct fo_plan_##SFX *);                                                   \
  size_t fo_size_##SFX(const struct fo_plan_##SFX *);                                              \
  int fo_is_bluestein_##SFX(const struct fo_plan_##SFX *);                                         \
  size_t fo_inner_size_##SFX(const struct fo_plan_##SFX *);                                        \
  void fo_counts_##SFX(const struct fo_plan_##SFX *, size_t *out5);                                \
  size_t fo_num_twiddles_##SFX(const struct fo_plan_##SFX *);                                      \
  void fo_copy_twiddles_##SFX(const struct fo_plan_##SFX *, int forward, REAL *out);               \
  void fo_transform_in_place_##SFX(struct fo_plan_##SFX *, REAL *data, int transform);             \
  void fo_transform_##SFX(struct fo_plan_##SFX *, const REAL *in, REAL *out, int transform);       \
  void fo_naive_dft_##SFX(const REAL *in, REAL *out, size_t n, int inverse);                       \
  void fo_fill_input_##SFX(REAL *out, uint64_t first_scalar, size_t count, uint64_t seed);         \
  int fo_transform_batch_##SFX(size_t n, const REAL *in, REAL *out, size_t batch, int transform,   \
                               int threads, double *seconds);

FO_DECLARE(f32, float)
FO_DECLARE(f64, double)

#ifdef __cplusplus
}
#endif
#endif /* FOURIER_ORACLE_H_ */
