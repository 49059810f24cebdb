// capi.cu -- the C ABI of libfourier.so: the eight reference symbols (include/fourier.h, replacing
// fourier-ffi/src/lib.rs:14-106) and the additive batched / device-pointer entry points
// (include/fourier_b200.h).  No torch types, plain pointers and sizes only.
//
// Error convention of the reference symbols (fourier-ffi/src/lib.rs): create returns NULL when the
// constructor fails; transform_* / destroy swallow every failure (catch_unwind) and return nothing.
// Only the C ABI is exported; everything else in the library is built with -fvisibility=hidden.
#pragma GCC visibility push(default)
#include "../../include/fourier_b200.h"
#pragma GCC visibility pop

#include <cstdint>
#include <new>

#include "plan.h"

using fb200::Plan;

namespace {

thread_local int g_device = -1;

int resolve_device() {
  if (g_device >= 0) return g_device;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  return dev;
}

// host memory (pageable or pinned) -> false; device / managed memory -> true
bool is_device_pointer(const void* p) {
  cudaPointerAttributes attr;
  if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged;
}

// The kernels read and write device buffers with 16-byte vector accesses (cpx<double> as double2, f32 pairs as
// float4, TMA tensor maps on the input): a misaligned device pointer would fault and poison the CUDA context, so it
// is refused up front.  (cudaMalloc and torch allocations are 256/512-byte aligned; a slice starting at an odd f32
// sample is not.)  Host pointers are staged through the library's own buffers and may have any alignment.
bool device_pointers_aligned(const void* in, const void* out) {
  if ((((uintptr_t)in) | ((uintptr_t)out)) & 15) {
    fb200::set_last_error("device buffers must be 16-byte aligned");
    return false;
  }
  return true;
}

template <typename T>
void* create_plan(size_t size, bool fast) {
  const int dev = resolve_device();
  if (dev < 0) {
    fb200::set_last_error("no usable CUDA device (libfourier.so has no CPU fallback)");
    return nullptr;
  }
  try {
    return Plan<T>::create(size, dev, fast);
  } catch (...) {
    fb200::set_last_error("plan construction threw");
    return nullptr;
  }
}

template <typename T>
int transform_batch(const void* plan, const void* in, void* out, size_t batch, int code) {
  if (!plan || !in || !out) { fb200::set_last_error("null argument"); return (int)cudaErrorInvalidValue; }
  auto* p = const_cast<Plan<T>*>(static_cast<const Plan<T>*>(plan));
  using C = typename Plan<T>::C;
  try {
    const bool din = is_device_pointer(in), dout = is_device_pointer(out);
    if (din != dout) {
      fb200::set_last_error("input and output must both be host or both be device memory");
      return (int)cudaErrorInvalidValue;
    }
    if (!din) return (int)p->exec_host((const C*)in, (C*)out, batch, code);
    if (!device_pointers_aligned(in, out)) return (int)cudaErrorMisalignedAddress;
    cudaError_t e = p->exec_device((const C*)in, (C*)out, batch, code, cudaStreamPerThread);
    if (e != cudaSuccess) return (int)e;
    return (int)cudaStreamSynchronize(cudaStreamPerThread);
  } catch (...) {
    fb200::set_last_error("transform threw");
    return (int)cudaErrorUnknown;
  }
}

template <typename T>
int transform_async(const void* plan, const void* in, void* out, size_t batch, int code, void* stream) {
  if (!plan || !in || !out) { fb200::set_last_error("null argument"); return (int)cudaErrorInvalidValue; }
  auto* p = const_cast<Plan<T>*>(static_cast<const Plan<T>*>(plan));
  using C = typename Plan<T>::C;
  try {
    if (!device_pointers_aligned(in, out)) return (int)cudaErrorMisalignedAddress;
    return (int)p->exec_device((const C*)in, (C*)out, batch, code, (cudaStream_t)stream);
  } catch (...) {
    fb200::set_last_error("transform threw");
    return (int)cudaErrorUnknown;
  }
}

template <typename T>
int rows_exchange(const void* plan, const void* in, size_t rows, int forward, void* const* outs, int nranks,
                  size_t out_ld, size_t out_off, int twiddle, unsigned long long row0, unsigned long long n_total,
                  void* stream) {
  if (!plan || !in || !outs) { fb200::set_last_error("null argument"); return (int)cudaErrorInvalidValue; }
  auto* p = const_cast<Plan<T>*>(static_cast<const Plan<T>*>(plan));
  using C = typename Plan<T>::C;
  try {
    if (!device_pointers_aligned(in, in)) return (int)cudaErrorMisalignedAddress;
    return (int)p->exec_rows_exchange((const C*)in, rows, forward != 0, outs, nranks, out_ld, out_off, twiddle, row0,
                                      n_total, (cudaStream_t)stream);
  } catch (...) {
    fb200::set_last_error("rows_exchange threw");
    return (int)cudaErrorUnknown;
  }
}

template <typename T>
int plan_info(const void* plan, fourier_b200_plan_info* out) {
  if (!plan || !out) return (int)cudaErrorInvalidValue;
  const auto* p = static_cast<const Plan<T>*>(plan);
  const fb200::PlanInfo i = p->info();
  out->size = i.size; out->path = i.path; out->inner_size = i.inner_size; out->inner_path = i.inner_path;
  out->n1 = i.n1; out->n2 = i.n2; out->precision_bytes = i.precision_bytes; out->device = i.device;
  out->table_bytes = i.table_bytes; out->last_launches = p->launches();
  return 0;
}

}  // namespace

using fourier::c::fourier_fft_double;
using fourier::c::fourier_fft_float;

namespace fourier {
namespace c {
extern "C" {

// ---- the eight reference symbols --------------------------------------------------------------------
fourier_fft_float* fourier_create_float(size_t size) {
  return static_cast<fourier_fft_float*>(create_plan<float>(size, true));
}
fourier_fft_double* fourier_create_double(size_t size) {
  return static_cast<fourier_fft_double*>(create_plan<double>(size, true));
}
void fourier_destroy_float(fourier_fft_float* plan) {
  try { delete reinterpret_cast<Plan<float>*>(plan); } catch (...) {}
}
void fourier_destroy_double(fourier_fft_double* plan) {
  try { delete reinterpret_cast<Plan<double>*>(plan); } catch (...) {}
}
void fourier_transform_in_place_float(const fourier_fft_float* plan, std::complex<float>* data, int t) {
  (void)transform_batch<float>(plan, data, data, 1, t);
}
void fourier_transform_in_place_double(const fourier_fft_double* plan, std::complex<double>* data, int t) {
  (void)transform_batch<double>(plan, data, data, 1, t);
}
void fourier_transform_float(const fourier_fft_float* plan, const std::complex<float>* in,
                             std::complex<float>* out, int t) {
  (void)transform_batch<float>(plan, in, out, 1, t);
}
void fourier_transform_double(const fourier_fft_double* plan, const std::complex<double>* in,
                              std::complex<double>* out, int t) {
  (void)transform_batch<double>(plan, in, out, 1, t);
}

}  // extern "C"
}  // namespace c
}  // namespace fourier

extern "C" {

// ---- additive entry points --------------------------------------------------------------------------
int fourier_b200_set_device(int device) {
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) {
    fb200::set_last_error("invalid device ordinal");
    return (int)cudaErrorInvalidDevice;
  }
  g_device = device;
  return 0;
}
int fourier_b200_get_device(void) { return resolve_device(); }
int fourier_b200_device_count(void) {
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess) { cudaGetLastError(); return 0; }
  return count;
}

int fourier_b200_transform_batch_float(const fourier_fft_float* plan, const void* in, void* out,
                                       size_t batch, int t) {
  return transform_batch<float>(plan, in, out, batch, t);
}
int fourier_b200_transform_batch_double(const fourier_fft_double* plan, const void* in, void* out,
                                        size_t batch, int t) {
  return transform_batch<double>(plan, in, out, batch, t);
}
int fourier_b200_transform_batch_async_float(const fourier_fft_float* plan, const void* in, void* out,
                                             size_t batch, int t, void* stream) {
  return transform_async<float>(plan, in, out, batch, t, stream);
}
int fourier_b200_transform_batch_async_double(const fourier_fft_double* plan, const void* in, void* out,
                                              size_t batch, int t, void* stream) {
  return transform_async<double>(plan, in, out, batch, t, stream);
}

int fourier_b200_plan_info_float(const fourier_fft_float* plan, fourier_b200_plan_info* out) {
  return plan_info<float>(plan, out);
}
int fourier_b200_plan_info_double(const fourier_fft_double* plan, fourier_b200_plan_info* out) {
  return plan_info<double>(plan, out);
}
const char* fourier_b200_path_name(int path) { return fb200::path_name((fb200::Path)path); }
const char* fourier_b200_plan_kernel_float(const fourier_fft_float* plan) {
  return plan ? reinterpret_cast<const Plan<float>*>(plan)->kernel_name() : "";
}
const char* fourier_b200_plan_kernel_double(const fourier_fft_double* plan) {
  return plan ? reinterpret_cast<const Plan<double>*>(plan)->kernel_name() : "";
}

fourier_fft_float* fourier_b200_create_general_float(size_t size) {
  return static_cast<fourier_fft_float*>(create_plan<float>(size, false));
}
fourier_fft_double* fourier_b200_create_general_double(size_t size) {
  return static_cast<fourier_fft_double*>(create_plan<double>(size, false));
}

int fourier_b200_fill_input_float(void* dev_out, unsigned long long first, size_t count,
                                  unsigned long long seed, void* stream) {
  return (int)fb200::launch_fill_input<float>((float*)dev_out, first, count, seed, (cudaStream_t)stream);
}
int fourier_b200_fill_input_double(void* dev_out, unsigned long long first, size_t count,
                                   unsigned long long seed, void* stream) {
  return (int)fb200::launch_fill_input<double>((double*)dev_out, first, count, seed, (cudaStream_t)stream);
}

int fourier_b200_transpose_float(const void* in, void* out, size_t batch, size_t rows, size_t cols, void* stream) {
  return (int)fb200::launch_transpose<float>((const float2*)in, (float2*)out, batch, rows, cols, (cudaStream_t)stream);
}
int fourier_b200_transpose_double(const void* in, void* out, size_t batch, size_t rows, size_t cols, void* stream) {
  return (int)fb200::launch_transpose<double>((const double2*)in, (double2*)out, batch, rows, cols, (cudaStream_t)stream);
}
int fourier_b200_pack_float(const void* in, void* out, size_t batch, size_t rows, size_t cols, size_t ld, size_t ibs,
                            size_t obs, int twiddle, unsigned long long row0, unsigned long long col0,
                            unsigned long long n_total, void* stream) {
  return (int)fb200::launch_pack<float>((const float2*)in, (float2*)out, batch, rows, cols, ld, ibs, obs, twiddle, row0,
                                        col0, n_total, (cudaStream_t)stream);
}
int fourier_b200_pack_double(const void* in, void* out, size_t batch, size_t rows, size_t cols, size_t ld, size_t ibs,
                             size_t obs, int twiddle, unsigned long long row0, unsigned long long col0,
                             unsigned long long n_total, void* stream) {
  return (int)fb200::launch_pack<double>((const double2*)in, (double2*)out, batch, rows, cols, ld, ibs, obs, twiddle, row0,
                                         col0, n_total, (cudaStream_t)stream);
}
int fourier_b200_exchange_float(const void* in, void* const* outs, int nranks, int me, size_t rows, size_t cb, size_t ld,
                                size_t out_ld, size_t out_off, int twiddle, unsigned long long row0,
                                unsigned long long n_total, void* stream) {
  return (int)fb200::launch_exchange<float>((const float2*)in, outs, nranks, me, rows, cb, ld, out_ld, out_off, twiddle,
                                            row0, n_total, (cudaStream_t)stream);
}
int fourier_b200_exchange_double(const void* in, void* const* outs, int nranks, int me, size_t rows, size_t cb, size_t ld,
                                 size_t out_ld, size_t out_off, int twiddle, unsigned long long row0,
                                 unsigned long long n_total, void* stream) {
  return (int)fb200::launch_exchange<double>((const double2*)in, outs, nranks, me, rows, cb, ld, out_ld, out_off, twiddle,
                                             row0, n_total, (cudaStream_t)stream);
}
int fourier_b200_fft_rows_exchange_float(const fourier_fft_float* plan, const void* in, size_t rows, int forward,
                                         void* const* outs, int nranks, size_t out_ld, size_t out_off, int twiddle,
                                         unsigned long long row0, unsigned long long n_total, void* stream) {
  return rows_exchange<float>(plan, in, rows, forward, outs, nranks, out_ld, out_off, twiddle, row0, n_total, stream);
}
int fourier_b200_fft_rows_exchange_double(const fourier_fft_double* plan, const void* in, size_t rows, int forward,
                                          void* const* outs, int nranks, size_t out_ld, size_t out_off, int twiddle,
                                          unsigned long long row0, unsigned long long n_total, void* stream) {
  return rows_exchange<double>(plan, in, rows, forward, outs, nranks, out_ld, out_off, twiddle, row0, n_total, stream);
}
int fourier_b200_peer_alloc(size_t bytes, void** dev_ptr, void* handle64) { return (int)fb200::peer_alloc(bytes, dev_ptr, handle64); }
int fourier_b200_peer_open(const void* handle64, void** dev_ptr) { return (int)fb200::peer_open(handle64, dev_ptr); }
int fourier_b200_peer_close(void* dev_ptr) { return (int)fb200::peer_close(dev_ptr); }
int fourier_b200_peer_free(void* dev_ptr) { return (int)fb200::peer_free(dev_ptr); }
int fourier_b200_swap_leading_float(const void* in, void* out, size_t a, size_t b, size_t inner, void* stream) {
  return (int)fb200::launch_swap_leading<float>((const float2*)in, (float2*)out, a, b, inner, (cudaStream_t)stream);
}
int fourier_b200_swap_leading_double(const void* in, void* out, size_t a, size_t b, size_t inner, void* stream) {
  return (int)fb200::launch_swap_leading<double>((const double2*)in, (double2*)out, a, b, inner, (cudaStream_t)stream);
}
int fourier_b200_twiddle_rows_float(void* data, size_t rows, size_t cols, unsigned long long row0,
                                    unsigned long long n_total, int forward, void* stream) {
  return (int)fb200::launch_twiddle_rows<float>((float2*)data, rows, cols, row0, n_total, forward != 0, (cudaStream_t)stream);
}
int fourier_b200_twiddle_rows_double(void* data, size_t rows, size_t cols, unsigned long long row0,
                                     unsigned long long n_total, int forward, void* stream) {
  return (int)fb200::launch_twiddle_rows<double>((double2*)data, rows, cols, row0, n_total, forward != 0, (cudaStream_t)stream);
}

const char* fourier_b200_last_error(void) { return fb200::last_error(); }
const char* fourier_b200_version(void) { return "fourier-b200 0.1.0 (sm_100a)"; }

}  // extern "C"
