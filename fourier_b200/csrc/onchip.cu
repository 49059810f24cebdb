#include "plan.h"
namespace fb200 {
template <typename T> cudaError_t Plan<T>::init_onchip() { return cudaErrorNotSupported; }
template <typename T> cudaError_t Plan<T>::exec_onchip(const C*, C*, size_t, int, cudaStream_t) { return cudaErrorNotSupported; }
template <typename T> cudaError_t Plan<T>::exec_bluestein_fused(const C*, C*, size_t, int, cudaStream_t) { return cudaErrorNotSupported; }
template cudaError_t Plan<float>::init_onchip();
template cudaError_t Plan<double>::init_onchip();
template cudaError_t Plan<float>::exec_onchip(const C*, C*, size_t, int, cudaStream_t);
template cudaError_t Plan<double>::exec_onchip(const C*, C*, size_t, int, cudaStream_t);
template cudaError_t Plan<float>::exec_bluestein_fused(const C*, C*, size_t, int, cudaStream_t);
template cudaError_t Plan<double>::exec_bluestein_fused(const C*, C*, size_t, int, cudaStream_t);
}
