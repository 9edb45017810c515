// devbuf.h — growable device buffer owned by the engine.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace sgr {
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t n) {
    if (n <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    if (n < 256) n = 256;
    cudaError_t e = cudaMalloc(&p, n);
    if (e == cudaSuccess) cap = n;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};
}  // namespace sgr
