// Exercises include/fourier_fft.hpp: the C++ mirror of the reference's Rust operator interface
// (Transform, Fft trait, create_fft_f32/f64 -- fourier-algorithms/src/fft.rs:5-82, fourier/src/lib.rs:31-60).
#include "fourier_fft.hpp"

#include <cmath>
#include <cstdio>

template <typename Real> static bool check(std::unique_ptr<fourier::Fft<Real>> fft, std::size_t n) {
  if (fft->size() != n) return false;
  std::vector<std::complex<Real>> x(n), y(n), z(n);
  for (std::size_t i = 0; i < n; ++i) x[i] = {Real(std::sin(0.37 * i)), Real(std::cos(1.3 * i))};
  fft->fft(x, y);                 // Fft::fft
  fft->ifft(y, z);                // Fft::ifft (scaled by 1/N)
  Real err = 0, mag = 0;
  for (std::size_t i = 0; i < n; ++i) { err = std::max(err, std::abs(z[i] - x[i])); mag = std::max(mag, std::abs(x[i])); }
  if (err > (sizeof(Real) == 4 ? 2e-5 : 1e-12) * mag) return false;
  z = x;
  fft->fft_in_place(z);           // Fft::fft_in_place == out-of-place result
  for (std::size_t i = 0; i < n; ++i) if (z[i] != y[i]) return false;
  // unitary pair: SqrtScaledFft then SqrtScaledIfft returns the input
  fft->transform(x.data(), y.data(), n, fourier::Transform::SqrtScaledFft);
  fft->transform_in_place(y.data(), n, fourier::Transform::SqrtScaledIfft);
  for (std::size_t i = 0; i < n; ++i) err = std::max(err, std::abs(y[i] - x[i]));
  return err <= (sizeof(Real) == 4 ? 2e-5 : 1e-12) * mag && fourier::is_forward(fourier::Transform::SqrtScaledFft);
}

int main() {
  const std::size_t sizes[] = {1, 6, 73, 1024, 1009, 4096};
  for (std::size_t n : sizes) {
    if (!check<float>(fourier::create_fft_f32(n), n)) { std::fprintf(stderr, "FAIL f32 %zu\n", n); return 1; }
    if (!check<double>(fourier::create_fft_f64(n), n)) { std::fprintf(stderr, "FAIL f64 %zu\n", n); return 1; }
  }
  bool threw = false;
  try { fourier::create_fft_f32(0); } catch (const std::runtime_error&) { threw = true; }
  if (!threw) { std::fprintf(stderr, "FAIL: size 0 must fail\n"); return 1; }
  std::printf("trait mirror test passed\n");
  return 0;
}
