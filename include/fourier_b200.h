/*
 * fourier_b200.h -- additive entry points of libfourier.so that the reference ABI has no
 * counterpart for: batched transforms, device-resident buffers, caller-supplied CUDA streams,
 * device selection and plan introspection.  Nothing here changes the eight reference symbols
 * declared in fourier.h.
 *
 * Why they exist: the reference's hot path is Fft::transform_in_place on ONE transform
 * (fourier-algorithms/src/fft.rs:48); a GPU needs many independent transforms per call to fill
 * 148 SMs and HBM, and callers that already hold data on the device must not pay PCIe.  A batch
 * is `batch` contiguous transforms of size() samples each: transform b occupies samples
 * [b*size, (b+1)*size).  Every transform of a batch is independent, exactly as if the reference's
 * transform() were called in a loop (fourier-bench/benches/fft_bench.rs:36).
 *
 * All functions return 0 on success and a non-zero cudaError_t value otherwise;
 * fourier_b200_last_error() returns a description for the calling thread.
 */
#ifndef FOURIER_B200_H_
#define FOURIER_B200_H_

#include "fourier.h"

#ifdef __cplusplus
extern "C" {
#define FB200_F ::fourier::c::fourier_fft_float
#define FB200_D ::fourier::c::fourier_fft_double
#else
#define FB200_F struct fourier_fft_float
#define FB200_D struct fourier_fft_double
#endif
#define FB200_PLAN_F const FB200_F
#define FB200_PLAN_D const FB200_D

/* Device used by plans created afterwards on this thread (default: the current CUDA device). */
int fourier_b200_set_device(int device);
int fourier_b200_get_device(void);
int fourier_b200_device_count(void);

/* Batched Fft::transform (in != out) / Fft::transform_in_place (in == out).  Pointers may be host
 * memory (pipelined H2D -> transform -> D2H, returns when `out` is complete) or device memory on
 * the plan's GPU (enqueued on the plan's stream and synchronised before returning). */
int fourier_b200_transform_batch_float(FB200_PLAN_F *plan, const void *in, void *out, size_t batch,
                                       int transform);
int fourier_b200_transform_batch_double(FB200_PLAN_D *plan, const void *in, void *out, size_t batch,
                                        int transform);

/* Device pointers only; enqueued on `cuda_stream` (a cudaStream_t, NULL = default stream) and NOT
 * synchronised: the call returns as soon as the kernels are queued.
 * Device buffers (here and above) must be 16-byte aligned (cudaMalloc / torch allocations are; a slice that
 * starts at an odd f32 sample is not): a misaligned pointer is refused with cudaErrorMisalignedAddress.
 * A plan owns ONE set of scratch buffers and work-queue counters: besides "not two threads at once" (a plan is
 * Send, not Sync, as in the reference) this means ONE STREAM AT A TIME -- do not enqueue a plan on a second
 * stream while a call on another stream is still running; use one plan per stream. */
int fourier_b200_transform_batch_async_float(FB200_PLAN_F *plan, const void *in_dev, void *out_dev,
                                             size_t batch, int transform, void *cuda_stream);
int fourier_b200_transform_batch_async_double(FB200_PLAN_D *plan, const void *in_dev, void *out_dev,
                                              size_t batch, int transform, void *cuda_stream);

/* Plan introspection (Fft::size, fourier-algorithms/src/fft.rs:45, and the chosen strategy). */
struct fourier_b200_plan_info {
  size_t size;        /* transform length N */
  int path;           /* 0 trivial, 1 onchip, 2 twopass, 3 global_stages, 4 bluestein, 5 bluestein_fused, 6 onchip_cta, 7 threepass */
  size_t inner_size;  /* Bluestein inner length next_pow2(2N-1) (bluesteins.rs:110), else 0 */
  int inner_path;
  size_t n1, n2;      /* two-pass split N = n1*n2, else 0 */
  int precision_bytes;
  int device;
  size_t table_bytes; /* twiddle / chirp tables resident in HBM */
  unsigned long long last_launches; /* kernels launched by the most recent transform call */
};
int fourier_b200_plan_info_float(FB200_PLAN_F *plan, struct fourier_b200_plan_info *out);
int fourier_b200_plan_info_double(FB200_PLAN_D *plan, struct fourier_b200_plan_info *out);
const char *fourier_b200_path_name(int path);
/* Name of the kernel that moves (nearly) all of the plan's bytes, as a profiler lists it. */
const char *fourier_b200_plan_kernel_float(FB200_PLAN_F *plan);
const char *fourier_b200_plan_kernel_double(FB200_PLAN_D *plan);

/* Plans restricted to the general one-kernel-per-stage path (testing / comparison). */
FB200_F *fourier_b200_create_general_float(size_t size);
FB200_D *fourier_b200_create_general_double(size_t size);

/* Synthetic benchmark input on the device: scalars [first_scalar, first_scalar+count) of the
 * counter-hash stream with `seed`, uniform in [-1, 1) (re of sample s is scalar 2s, im 2s+1). */
int fourier_b200_fill_input_float(void *dev_out, unsigned long long first_scalar, size_t count,
                                  unsigned long long seed, void *cuda_stream);
int fourier_b200_fill_input_double(void *dev_out, unsigned long long first_scalar, size_t count,
                                   unsigned long long seed, void *cuda_stream);

/* Building blocks of the distributed six-step transform (one huge N = N1*N2 over several GPUs with
 * all-to-all transposes between local batched FFTs; fourier_b200/distributed.py drives them):
 * batched 2-D transpose out[b][c][r] = in[b][r][c], and data[r][c] *= w_N^{(row0+r)*c} (conj if !forward)
 * with the index reduced exactly mod N.  Device pointers, enqueued on cuda_stream. */
int fourier_b200_transpose_float(const void *in_dev, void *out_dev, size_t batch, size_t rows, size_t cols,
                                 void *cuda_stream);
int fourier_b200_transpose_double(const void *in_dev, void *out_dev, size_t batch, size_t rows, size_t cols,
                                  void *cuda_stream);
/* Pack step of a pipelined exchange: batched transpose of column blocks of a row-major matrix with leading
 * dimension ld, optionally fused with the inter-step twiddle (twiddle = 0 none, 1 forward, 2 inverse):
 *   out[b*out_batch_stride + c*rows + r] = in[b*in_batch_stride + r*ld + c] * w_N^{(row0+r)*(col0+b*in_batch_stride+c)} */
int fourier_b200_pack_float(const void *in_dev, void *out_dev, size_t batch, size_t rows, size_t cols, size_t ld,
                            size_t in_batch_stride, size_t out_batch_stride, int twiddle, unsigned long long row0,
                            unsigned long long col0, unsigned long long n_total, void *cuda_stream);
int fourier_b200_pack_double(const void *in_dev, void *out_dev, size_t batch, size_t rows, size_t cols, size_t ld,
                             size_t in_batch_stride, size_t out_batch_stride, int twiddle, unsigned long long row0,
                             unsigned long long col0, unsigned long long n_total, void *cuda_stream);
/* The whole exchange as ONE kernel over NVLink peer memory: rank `me` of `nranks` holds `rows` rows of
 * ld = nranks*cb columns and stores columns [q*cb, (q+1)*cb) transposed into rank q's buffer outs[q]
 * (device pointers valid on this GPU: own memory for q == me, fourier_b200_peer_open()ed memory otherwise):
 *   outs[q][out_off + c*out_ld + r] = in[r*ld + q*cb + c] * w_N^{(row0+r)*(q*cb+c)}      (twiddle as above)
 * `outs` is a HOST array of nranks pointers.  The caller orders ranks with a stream-ordered barrier. */
int fourier_b200_exchange_float(const void *in_dev, void *const *outs, int nranks, int me, size_t rows, size_t cb,
                                size_t ld, size_t out_ld, size_t out_off, int twiddle, unsigned long long row0,
                                unsigned long long n_total, void *cuda_stream);
int fourier_b200_exchange_double(const void *in_dev, void *const *outs, int nranks, int me, size_t rows, size_t cb,
                                 size_t ld, size_t out_ld, size_t out_off, int twiddle, unsigned long long row0,
                                 unsigned long long n_total, void *cuda_stream);
/* The row FFTs of a step AND the exchange that follows them as one pass over the data: `rows` contiguous rows of the
 * plan's size N are transformed (unscaled; forward != 0: Fft, else UnscaledIfft) and the last register stage of the
 * transform stores its results straight into the destination ranks' buffers, transposed and twiddled:
 *   outs[q][c*out_ld + out_off + r] = FFT_N(in[r*N ..])[q*cb + c] * w_Ntot^{(row0+r)*(q*cb+c)},   cb = N / nranks
 * i.e. what fourier_b200_transform_batch_async_* followed by fourier_b200_exchange_* delivers, with one kernel and
 * one sweep over HBM less and the NVLink stores overlapping the butterflies.  Two-pass plans only (power-of-two N,
 * 2^11 .. 2^20 f32 / 2^9 .. 2^16 f64); nranks a power of two; rows a multiple of the tile height (32 up to N = 2^12 (f32), 8 from 2^17, else 16); twiddle 0,
 * 1 (with forward) or 2 (with inverse); `in` is left intact.  Returns cudaErrorNotSupported (801) for other plans. */
int fourier_b200_fft_rows_exchange_float(const FB200_F *plan, const void *in_dev, size_t rows,
                                         int forward, void *const *outs, int nranks, size_t out_ld, size_t out_off,
                                         int twiddle, unsigned long long row0, unsigned long long n_total,
                                         void *cuda_stream);
int fourier_b200_fft_rows_exchange_double(const FB200_D *plan, const void *in_dev, size_t rows,
                                          int forward, void *const *outs, int nranks, size_t out_ld, size_t out_off,
                                          int twiddle, unsigned long long row0, unsigned long long n_total,
                                          void *cuda_stream);
/* Buffers other ranks of the box (one process per GPU) can store into: cudaMalloc + CUDA IPC.  `handle64` is
 * the 64-byte cudaIpcMemHandle_t to send to the peers (any transport), who open it with _peer_open. */
int fourier_b200_peer_alloc(size_t bytes, void **dev_ptr, void *handle64);
int fourier_b200_peer_open(const void *handle64, void **dev_ptr);
int fourier_b200_peer_close(void *dev_ptr);
int fourier_b200_peer_free(void *dev_ptr);
/* out[b][a][i] = in[a][b][i] (i < inner contiguous): unpack step after an all-to-all */
int fourier_b200_swap_leading_float(const void *in_dev, void *out_dev, size_t a, size_t b, size_t inner,
                                    void *cuda_stream);
int fourier_b200_swap_leading_double(const void *in_dev, void *out_dev, size_t a, size_t b, size_t inner,
                                     void *cuda_stream);
int fourier_b200_twiddle_rows_float(void *data_dev, size_t rows, size_t cols, unsigned long long row0,
                                    unsigned long long n_total, int forward, void *cuda_stream);
int fourier_b200_twiddle_rows_double(void *data_dev, size_t rows, size_t cols, unsigned long long row0,
                                     unsigned long long n_total, int forward, void *cuda_stream);

const char *fourier_b200_last_error(void);
const char *fourier_b200_version(void);

#undef FB200_PLAN_F
#undef FB200_PLAN_D
#undef FB200_F
#undef FB200_D
#ifdef __cplusplus
}
#endif
#endif /* FOURIER_B200_H_ */
