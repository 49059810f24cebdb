/*
 * fourier_oracle.c -- instantiates the oracle for f32 and f64.
 * TEST INFRASTRUCTURE ONLY; see fourier_oracle.h for scope and pinning status.
 * Build: make -C oracle   (gcc -O2 -ffp-contract=off, no -ffast-math).
 */
#define _POSIX_C_SOURCE 200809L
#include "fourier_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define FO_PI 3.14159265358979323846264338327950288 /* core::f64::consts::PI */

/* RADICES, fourier-algorithms/src/autosort/mod.rs:20-21 */
#define FO_NUM_RADICES 5
static const size_t FO_RADICES[FO_NUM_RADICES] = {4, 8, 4, 3, 2};

int fo_is_forward(int transform) { return transform == FO_FFT || transform == FO_SQRT_SCALED_FFT; }

uint64_t fo_hash64(uint64_t seed, uint64_t counter) {
  uint64_t z = seed + (counter + 1) * 0x9E3779B97F4A7C15ULL;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}

static double fo_now(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* ---- f32 ---- */
#define REAL float
#define SFX f32
#define FO_SQRT sqrtf
/* top 24 bits -> [-1, 1) on a 2^-23 grid (exact in f32) */
#define FO_UNIT(h) ((float)((h) >> 40) * (1.0f / 8388608.0f) - 1.0f)
#include "fourier_oracle_impl.inc"
#undef REAL
#undef SFX
#undef FO_SQRT
#undef FO_UNIT

/* ---- f64 ---- */
#define REAL double
#define SFX f64
#define FO_SQRT sqrt
/* top 53 bits -> [-1, 1) on a 2^-52 grid (exact in f64) */
#define FO_UNIT(h) ((double)((h) >> 11) * (1.0 / 4503599627370496.0) - 1.0)
#include "fourier_oracle_impl.inc"
#undef REAL
#undef SFX
#undef FO_SQRT
#undef FO_UNIT
