/* Drop-in acceptance program for the C ABI (include/fourier.h), written for this repo in the spirit
 * of the reference's fourier-ffi/test.c:7-39 (4-point impulse FFT -> in-place IFFT round trip within
 * 1e-10) and extended with a 1024-point known-answer case (BASELINE.json config 1) and the
 * error-convention checks of SURVEY.md 8b.  Exit code 0 = pass. */
#include "fourier.h"
#include <complex.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

static int fail(const char *what) { fprintf(stderr, "FAIL: %s\n", what); return 1; }

static int impulse_float(void) {
  float complex in[4] = {1, 0, 0, 0}, out[4];
  struct fourier_fft_float *p = fourier_create_float(4);
  if (!p) return fail("fourier_create_float(4) returned NULL");
  fourier_transform_float(p, in, out, FOURIER_TRANSFORM_FFT);
  for (int i = 0; i < 4; i++)
    if (cabsf(out[i] - 1.0f) > 1e-6f) return fail("float impulse spectrum is not flat");
  fourier_transform_in_place_float(p, out, FOURIER_TRANSFORM_IFFT);
  fourier_destroy_float(p);
  for (int i = 0; i < 4; i++)
    if (cabsf(in[i] - out[i]) > 1e-10f) return fail("float round trip");
  return 0;
}

static int impulse_double(void) {
  double complex in[4] = {1, 0, 0, 0}, out[4];
  struct fourier_fft_double *p = fourier_create_double(4);
  if (!p) return fail("fourier_create_double(4) returned NULL");
  fourier_transform_double(p, in, out, FOURIER_TRANSFORM_FFT);
  fourier_transform_in_place_double(p, out, FOURIER_TRANSFORM_IFFT);
  fourier_destroy_double(p);
  for (int i = 0; i < 4; i++)
    if (cabs(in[i] - out[i]) > 1e-10) return fail("double round trip");
  return 0;
}

/* x[n] = exp(2*pi*i*5n/N) -> X[k] = N at k = 5, 0 elsewhere */
static int tone_1024(void) {
  enum { N = 1024 };
  static float complex in[N], out[N];
  const double two_pi = 6.283185307179586476925286766559;
  for (int n = 0; n < N; n++) in[n] = (float)cos(two_pi * 5 * n / N) + I * (float)sin(two_pi * 5 * n / N);
  struct fourier_fft_float *p = fourier_create_float(N);
  if (!p) return fail("fourier_create_float(1024) returned NULL");
  fourier_transform_float(p, in, out, FOURIER_TRANSFORM_FFT);
  fourier_destroy_float(p);
  for (int k = 0; k < N; k++) {
    float complex want = (k == 5) ? (float)N : 0.0f;
    if (cabsf(out[k] - want) > 1e-5f * N) return fail("1024-point tone spectrum");
  }
  return 0;
}

static int conventions(void) {
  if (fourier_create_float(0) != NULL) return fail("size 0 must be refused with NULL");
  float complex x[4] = {1, 2, 3, 4}, y[4] = {1, 2, 3, 4};
  struct fourier_fft_float *p = fourier_create_float(4);
  if (!p) return fail("create");
  fourier_transform_in_place_float(p, x, 99); /* unknown code: silent no-op (ffi lib.rs:10,37) */
  for (int i = 0; i < 4; i++)
    if (x[i] != y[i]) return fail("unknown transform code must leave the buffer untouched");
  fourier_transform_in_place_float(NULL, x, 0); /* must not crash */
  fourier_destroy_float(p);
  fourier_destroy_float(NULL); /* tolerated */
  return 0;
}

int main(void) {
  int bad = impulse_float() + impulse_double() + tone_1024() + conventions();
  if (bad) return 1;
  printf("drop-in C ABI test passed\n");
  return 0;
}
