// fourier_fft.hpp -- C++ mirror of the reference's Rust operator interface over the C ABI:
//   fourier::Transform / Fft trait (fourier-algorithms/src/fft.rs:5-82) and
//   fourier::create_fft_f32 / create_fft_f64 (fourier/src/lib.rs:31-60),
// with the same names, argument meaning and failure behaviour (size mismatch = assertion; plan
// construction failure = exception where Rust would panic).  Header only; link with -lfourier.
#ifndef FOURIER_FFT_HPP_
#define FOURIER_FFT_HPP_

#include <cassert>
#include <complex>
#include <cstddef>
#include <memory>
#include <stdexcept>
#include <vector>

#include "fourier_b200.h"

namespace fourier {

enum class Transform : int { Fft = 0, Ifft = 1, UnscaledIfft = 2, SqrtScaledFft = 3, SqrtScaledIfft = 4 };
inline bool is_forward(Transform t) { return t == Transform::Fft || t == Transform::SqrtScaledFft; }

template <typename Real> class Fft {
 public:
  virtual ~Fft() = default;
  virtual std::size_t size() const = 0;
  // required method (fft.rs:48)
  virtual void transform_in_place(std::complex<Real>* input, std::size_t len, Transform t) const = 0;
  // provided methods (fft.rs:51-81)
  virtual void transform(const std::complex<Real>* input, std::complex<Real>* output, std::size_t len,
                         Transform t) const = 0;
  void fft_in_place(std::vector<std::complex<Real>>& x) const { transform_in_place(x.data(), x.size(), Transform::Fft); }
  void ifft_in_place(std::vector<std::complex<Real>>& x) const { transform_in_place(x.data(), x.size(), Transform::Ifft); }
  void fft(const std::vector<std::complex<Real>>& in, std::vector<std::complex<Real>>& out) const {
    transform(in.data(), out.data(), in.size(), Transform::Fft);
  }
  void ifft(const std::vector<std::complex<Real>>& in, std::vector<std::complex<Real>>& out) const {
    transform(in.data(), out.data(), in.size(), Transform::Ifft);
  }
};

namespace detail_fft {
template <typename Real> class GpuFft final : public Fft<Real> {
 public:
  explicit GpuFft(std::size_t n) : plan_(n), size_(n) {
    if (!plan_.get()) throw std::runtime_error("fourier: plan construction failed (no CPU fallback)");
  }
  std::size_t size() const override { return size_; }
  void transform_in_place(std::complex<Real>* x, std::size_t len, Transform t) const override {
    assert(len == size_);
    plan_.transform_in_place(x, static_cast<::fourier::transform>(static_cast<int>(t)));
  }
  void transform(const std::complex<Real>* in, std::complex<Real>* out, std::size_t len, Transform t) const override {
    assert(len == size_);
    plan_.transform(in, out, static_cast<::fourier::transform>(static_cast<int>(t)));
  }
 private:
  ::fourier::fft<Real> plan_;
  std::size_t size_;
};
}  // namespace detail_fft

inline std::unique_ptr<Fft<float>> create_fft_f32(std::size_t size) {
  return std::unique_ptr<Fft<float>>(new detail_fft::GpuFft<float>(size));
}
inline std::unique_ptr<Fft<double>> create_fft_f64(std::size_t size) {
  return std::unique_ptr<Fft<double>>(new detail_fft::GpuFft<double>(size));
}

}  // namespace fourier
#endif  // FOURIER_FFT_HPP_
