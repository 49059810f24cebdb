// ubench.cu -- hardware micro-measurements that the kernel design in DESIGN.md relies on (B200):
//   (1) FP32 / packed-FP32x2 / FP64 issue throughput per SM, (2) HBM bandwidth of the strided
//   "column tile" access pattern of the four-step FFT as a function of the contiguous segment size.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/ubench.cu -o build/ubench
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

template <int MODE>
__global__ void __launch_bounds__(256) fp_kernel(float* out, int iters, float a, float b) {
  // 16 independent chains per thread
  float2 v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f - i);
  const float2 aa = make_float2(a, a * 0.5f), bb = make_float2(b, b * 0.25f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) { v[i].x = fmaf(v[i].x, aa.x, bb.x); v[i].y = fmaf(v[i].y, aa.y, bb.y); }
      if (MODE == 1) { v[i].x = v[i].x + bb.x; v[i].y = v[i].y + bb.y; }
      if (MODE == 2) { v[i] = __ffma2_rn(v[i], aa, bb); }
      if (MODE == 3) { v[i] = __fadd2_rn(v[i], bb); }
      if (MODE == 4) { v[i].x = v[i].x * aa.x; v[i].y = v[i].y * aa.y; }
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += v[i].x + v[i].y;
  if (s == 12345.678f) out[0] = s;
}

template <int MODE>
__global__ void __launch_bounds__(256) fp64_kernel(double* out, int iters, double a, double b) {
  double v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) v[i] = fma(v[i], a, b);
      if (MODE == 1) v[i] = v[i] + b;
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += v[i];
  if (s == 12345.678) out[0] = s;
}

// Copy rows x cols (float2 elements) matrix tile-wise: each CTA handles a tile of SEG consecutive float2
// columns and all `rows` rows (stride `cols`), reading then writing the same pattern to dst.
template <int SEG>
__global__ void __launch_bounds__(256) tile_copy(const float2* __restrict__ src, float2* __restrict__ dst,
                                                 int rows, size_t cols, size_t tiles_per_mat) {
  const size_t mat = blockIdx.x / tiles_per_mat, tile = blockIdx.x % tiles_per_mat;
  const float2* s = src + mat * rows * cols + tile * SEG;
  float2* d = dst + mat * rows * cols + tile * SEG;
  const int c = threadIdx.x % SEG, r0 = threadIdx.x / SEG;
  constexpr int RSTEP = 256 / SEG;
  constexpr int PER = (1024 / RSTEP < 32) ? 1024 / RSTEP : 32;  // rows == 1024
  float2 v[PER];
  for (int base = 0; base < rows; base += RSTEP * PER) {
#pragma unroll
    for (int i = 0; i < PER; ++i) v[i] = s[(size_t)(base + r0 + i * RSTEP) * cols + c];
#pragma unroll
    for (int i = 0; i < PER; ++i) d[(size_t)(base + r0 + i * RSTEP) * cols + c] = v[i];
  }
}

__global__ void __launch_bounds__(256) linear_copy(const float4* __restrict__ src, float4* __restrict__ dst, size_t n) {
  for (size_t i = blockIdx.x * (size_t)256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) dst[i] = src[i];
}

template <typename F> float time_ms(F f, int reps = 5) {
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  f(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(a)); f(); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b)); if (ms < best) best = ms;
  }
  return best;
}

int main(int argc, char** argv) {
  const bool bw_only = argc > 1 && argv[1][0] == 'b';
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("device %s SMs %d clock %d kHz L2 %d MB\n", p.name, p.multiProcessorCount, clk, p.l2CacheSize >> 20);
  const int sms = p.multiProcessorCount;
  float* dout; CK(cudaMalloc(&dout, 1024));
  const int iters = 20000;
  const int blocks = sms * 8;
  const char* names[] = {"FFMA x2 scalar", "FADD x2 scalar", "FFMA2 packed", "FADD2 packed", "FMUL x2 scalar"};
  for (int mode = 0; mode < (bw_only ? 0 : 5); ++mode) {
    float ms = 0;
    if (mode == 0) ms = time_ms([&] { fp_kernel<0><<<blocks, 256>>>(dout, iters, 1.0001f, 0.5f); });
    if (mode == 1) ms = time_ms([&] { fp_kernel<1><<<blocks, 256>>>(dout, iters, 1.0001f, 0.5f); });
    if (mode == 2) ms = time_ms([&] { fp_kernel<2><<<blocks, 256>>>(dout, iters, 1.0001f, 0.5f); });
    if (mode == 3) ms = time_ms([&] { fp_kernel<3><<<blocks, 256>>>(dout, iters, 1.0001f, 0.5f); });
    if (mode == 4) ms = time_ms([&] { fp_kernel<4><<<blocks, 256>>>(dout, iters, 1.0001f, 0.5f); });
    double lane_ops = (double)blocks * 256 * iters * 16;  // scalar f32 ops (each packed op counts 2)
    printf("%-16s %8.3f ms  %7.2f Tlane-op/s  (%.1f f32 lane-ops/clk/SM @1.965GHz)\n", names[mode], ms,
           lane_ops / ms * 1e-9, lane_ops / (ms * 1e-3) / sms / 1.965e9);
  }
  for (int mode = 0; mode < (bw_only ? 0 : 2); ++mode) {
    float ms = mode == 0 ? time_ms([&] { fp64_kernel<0><<<blocks, 256>>>((double*)dout, iters / 4, 1.0001, 0.5); })
                         : time_ms([&] { fp64_kernel<1><<<blocks, 256>>>((double*)dout, iters / 4, 1.0001, 0.5); });
    double lane_ops = (double)blocks * 256 * (iters / 4) * 8;
    printf("%-16s %8.3f ms  %7.2f Tlane-op/s  (%.1f f64 lane-ops/clk/SM @1.965GHz)\n", mode == 0 ? "DFMA" : "DADD", ms,
           lane_ops / ms * 1e-9, lane_ops / (ms * 1e-3) / sms / 1.965e9);
  }

  // bandwidth: 256 matrices of 1024 x 1024 float2 (2 GiB in, 2 GiB out)
  const int rows = 1024; const size_t cols = 1024, mats = 256;
  const size_t elems = mats * rows * cols;
  float2 *src, *dst; CK(cudaMalloc(&src, elems * 8)); CK(cudaMalloc(&dst, elems * 8));
  CK(cudaMemset(src, 1, elems * 8)); CK(cudaMemset(dst, 0, elems * 8));
  {
    float ms = time_ms([&] { linear_copy<<<sms * 16, 256>>>((const float4*)src, (float4*)dst, elems / 2); });
    printf("linear float4 copy            %7.3f ms  %7.1f GB/s (read+write)\n", ms, 2.0 * elems * 8 / ms * 1e-6);
  }
  {
    float ms = time_ms([&] { CK(cudaMemcpyAsync(dst, src, elems * 8, cudaMemcpyDeviceToDevice)); });
    printf("cudaMemcpy D2D                %7.3f ms  %7.1f GB/s (read+write)\n", ms, 2.0 * elems * 8 / ms * 1e-6);
  }
#define TILE(SEG) { float ms = time_ms([&] { tile_copy<SEG><<<(unsigned)(mats * cols / SEG), 256>>>(src, dst, rows, cols, cols / SEG); }); \
    printf("tile copy, %3d B segments      %7.3f ms  %7.1f GB/s (read+write)\n", SEG * 8, ms, 2.0 * elems * 8 / ms * 1e-6); }
  TILE(4) TILE(8) TILE(16) TILE(32) TILE(64)
  return 0;
}
