// twopass.cu -- large power-of-two N: the four-step FFT as TWO fused kernels with the intermediate
// kept in L2.
//
//   N = N1*N2, input index n = n1*N2 + n2, output index k = k1 + N1*k2
//   pass 1 (column tiles): for every n2:  A[k1][n2] = w_N^{n2*k1} * sum_{n1} x[n1*N2+n2] w_N1^{n1*k1}
//   pass 2 (row tiles):    for every k1:  X[k1 + N1*k2] = sum_{n2} A[k1][n2] w_N2^{n2*k2}
//
// Each pass is one TileFFT (tilefft.cuh): every sample is read once from global memory into
// registers, transformed by two register radix-R stages with one shared-memory exchange, and written
// once.  The reference streams the whole array once per radix-4/8 stage -- seven sweeps at N = 2^20
// (autosort/mod.rs:338-379); here HBM sees one read and one write per sample as long
// as the intermediate A of a few transforms stays resident in the 126 MB L2, which is what the
// chunking in exec_twopass() arranges.  All global accesses are >= 128-byte contiguous pieces:
//   pass 1 reads  x  as C consecutive columns (C*8 B per row),  writes A[k1][n2] the same way;
//   pass 2 reads  A  as whole contiguous rows,                  writes X as C consecutive k1.
#include <cstdlib>

#include "plan.h"
#include "twopass_kernels.cuh"

namespace fb200 {

#define FB_CHECK(expr)                                                                       \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) {                                                                 \
      set_last_error(std::string(#expr) + ": " + cudaGetErrorString(_e));                    \
      return _e;                                                                             \
    }                                                                                        \
  } while (0)

using namespace twopass;

template <typename T>
cudaError_t Plan<T>::init_twopass() {
  const TwoPassOps<T>* ops = lookup<T>(n_);
  if (!ops) return cudaErrorNotSupported;
  FB_CHECK(ops->prepare());
  n1_ = ops->n1;
  n2_ = ops->n2;
  FB_CHECK((upload_vec<T, TwPair<T>>(tw_a_, make_twa<T>(ops->ra1, ops->rb1))));
  FB_CHECK((upload_vec<T, TwPair<T>>(tw_b_, make_twa<T>(ops->ra2, ops->rb2))));
  // inter-pass twiddles in the layout of the intermediate: T[k1*N2 + n2] = w_N^{n2*k1}
  std::vector<cpx<T>> tw2(n_);
  for (size_t k1 = 0; k1 < n1_; ++k1)
    for (size_t c = 0; c < n2_; ++c) {
      double re, im;
      host_twiddle(k1 * c, n_, &re, &im);
      tw2[k1 * n2_ + c] = mk<T>((T)re, (T)im);
    }
  FB_CHECK((upload_vec<T, cpx<T>>(tw2_, tw2)));
  // transforms per chunk: the intermediate of one chunk should sit comfortably inside the L2
  size_t mb = 32;
  if (const char* env = std::getenv("FOURIER_B200_CHUNK_MB")) mb = (size_t)std::max(1, atoi(env));
  chunk_ = std::max<size_t>(1, (mb << 20) / (n_ * sizeof(C)));
  fast_ops_ = ops;
  return cudaSuccess;
}

template <typename T>
cudaError_t Plan<T>::exec_twopass(const C* in, C* out, size_t batch, int code, cudaStream_t s) {
  const auto* ops = static_cast<const TwoPassOps<T>*>(fast_ops_);
  const bool fwd = transform_is_forward(code);
  const bool do_scale = !(code == kFft || code == kUnscaledIfft);
  T scale = (T)1;
  if (code == kIfft) scale = (T)1 / (T)n_;
  else if (do_scale) scale = (T)1 / std::sqrt((T)n_);
  const size_t chunk = std::min(chunk_, batch);
  FB_CHECK(work_.reserve(chunk * n_ * sizeof(C)));
  C* scratch = (C*)work_.data();
  for (size_t b0 = 0; b0 < batch; b0 += chunk) {
    const size_t nb = std::min(chunk, batch - b0);
    FB_CHECK(ops->pass1(in + b0 * n_, scratch, tw_a_.data(), (const C*)tw2_.data(), nb, fwd, s));
    FB_CHECK(ops->pass2(scratch, out + b0 * n_, tw_b_.data(), nb, fwd, scale, do_scale, s));
    launches_ += 2;
  }
  return cudaSuccess;
}

template cudaError_t Plan<float>::init_twopass();
template cudaError_t Plan<double>::init_twopass();
template cudaError_t Plan<float>::exec_twopass(const C*, C*, size_t, int, cudaStream_t);
template cudaError_t Plan<double>::exec_twopass(const C*, C*, size_t, int, cudaStream_t);

}  // namespace fb200
