// synth.cu -- synthetic input generator for the benchmarks and the large-size parity tests.
// Same counter hash as oracle/fourier_oracle.c (fo_hash64 / fo_fill_input_*): scalar number g of the
// whole batch (re of sample s is g = 2s, im is g = 2s+1) -> splitmix64 finaliser -> U[-1, 1).
// The mapping uses only exact operations, so CPU and GPU produce identical bits.
#include "plan.h"

namespace fb200 {
namespace {

__host__ __device__ inline unsigned long long hash64(unsigned long long seed, unsigned long long counter) {
  unsigned long long z = seed + (counter + 1ull) * 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__device__ inline float unit(unsigned long long h, float) {
  return (float)(h >> 40) * (1.0f / 8388608.0f) - 1.0f;
}
__device__ inline double unit(unsigned long long h, double) {
  return (double)(h >> 11) * (1.0 / 4503599627370496.0) - 1.0;
}

template <typename T>
__global__ void __launch_bounds__(256)
fill_kernel(T* __restrict__ out, unsigned long long first, size_t count, unsigned long long seed) {
  for (size_t g = blockIdx.x * (size_t)blockDim.x + threadIdx.x; g < count;
       g += (size_t)gridDim.x * blockDim.x)
    out[g] = unit(hash64(seed, first + g), T());
}

}  // namespace

template <typename T>
cudaError_t launch_fill_input(T* out, unsigned long long first_scalar, size_t count,
                              unsigned long long seed, cudaStream_t s) {
  size_t blocks = (count + 255) / 256;
  if (blocks > 148u * 16u) blocks = 148u * 16u;
  if (blocks == 0) blocks = 1;
  fill_kernel<T><<<(unsigned)blocks, 256, 0, s>>>(out, first_scalar, count, seed);
  return cudaGetLastError();
}
template cudaError_t launch_fill_input<float>(float*, unsigned long long, size_t, unsigned long long, cudaStream_t);
template cudaError_t launch_fill_input<double>(double*, unsigned long long, size_t, unsigned long long, cudaStream_t);

}  // namespace fb200
