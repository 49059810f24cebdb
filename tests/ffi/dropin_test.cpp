// Drop-in acceptance program for the C++ surface of include/fourier.h (fourier::fft<T>,
// fourier::transform, namespace fourier::c), in the spirit of the reference's
// fourier-ffi/test.cpp:17-48, plus the batched extension of include/fourier_b200.h.
#include "fourier_b200.h"

#include <cmath>
#include <complex>
#include <cstdio>
#include <vector>

template <typename T> static bool round_trip() {
  std::vector<std::complex<T>> in{{1, 0}, {0, 0}, {0, 0}, {0, 0}}, out(4);
  fourier::fft<T> plan(in.size());
  plan.transform(in.data(), out.data(), fourier::transform::fft);
  plan.transform_in_place(out.data(), fourier::transform::ifft);
  for (int i = 0; i < 4; ++i)
    if (std::abs(in[i] - out[i]) > 1e-10) return false;
  fourier::fft<T> moved(std::move(plan));  // move-only handle, like the reference wrapper
  moved.transform(in.data(), out.data(), fourier::transform::sqrt_scaled_fft);
  for (int i = 0; i < 4; ++i)
    if (std::abs(out[i] - std::complex<T>(0.5, 0)) > 1e-6) return false;
  return true;
}

static bool raw_c_symbols() {
  using namespace fourier::c;
  std::vector<std::complex<double>> in{{1, 0}, {0, 0}, {0, 0}, {0, 0}}, out(4);
  fourier_fft_double* p = fourier_create_double(4);
  if (!p) return false;
  fourier_transform_double(p, in.data(), out.data(), FOURIER_TRANSFORM_FFT);
  fourier_transform_in_place_double(p, out.data(), FOURIER_TRANSFORM_IFFT);
  fourier_destroy_double(p);
  for (int i = 0; i < 4; ++i)
    if (std::abs(in[i] - out[i]) > 1e-10) return false;
  return true;
}

// batch of 3 transforms of size 6 through the extension == 3 single reference calls
static bool batch_extension() {
  const std::size_t n = 6, batch = 3;
  std::vector<std::complex<float>> in(n * batch), a(n * batch), b(n * batch);
  for (std::size_t i = 0; i < in.size(); ++i) in[i] = {float(i % 7) - 3.0f, float(i % 5) * 0.5f};
  fourier::fft<float> plan(n);
  for (std::size_t k = 0; k < batch; ++k)
    plan.transform(in.data() + k * n, a.data() + k * n, fourier::transform::fft);
  if (fourier_b200_transform_batch_float(plan.get(), in.data(), b.data(), batch,
                                         fourier::c::FOURIER_TRANSFORM_FFT) != 0)
    return false;
  for (std::size_t i = 0; i < in.size(); ++i)
    if (a[i] != b[i]) return false;
  fourier_b200_plan_info info;
  if (fourier_b200_plan_info_float(plan.get(), &info) != 0 || info.size != n) return false;
  return true;
}

int main() {
  if (!round_trip<float>()) { std::fprintf(stderr, "FAIL: fft<float>\n"); return 1; }
  if (!round_trip<double>()) { std::fprintf(stderr, "FAIL: fft<double>\n"); return 1; }
  if (!raw_c_symbols()) { std::fprintf(stderr, "FAIL: raw C symbols\n"); return 1; }
  if (!batch_extension()) { std::fprintf(stderr, "FAIL: batch extension\n"); return 1; }
  std::printf("drop-in C++ test passed\n");
  return 0;
}
