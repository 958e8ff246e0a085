// Stand-in for <cuda_runtime.h> in CPU emulation builds (tests/emu/): "device" memory is host memory, copies are
// memcpy, streams and events do nothing.  It lets HOST code of the library that only allocates and copies
// (e.g. setup_projection of csrc/hash.cu, through the real hs_ctx of csrc/common.cuh) run unchanged without a GPU.
#pragma once
#include <stdlib.h>
#include <string.h>

typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2 };
typedef struct emu_stream *cudaStream_t;
typedef struct emu_event *cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };

static inline cudaError_t cudaMalloc(void **p, size_t bytes) {
  *p = calloc(bytes ? bytes : 1, 1);
  return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
static inline cudaError_t cudaFree(void *p) {
  free(p);
  return cudaSuccess;
}
static inline cudaError_t cudaMemcpyAsync(void *dst, const void *src, size_t bytes, cudaMemcpyKind, cudaStream_t) {
  memcpy(dst, src, bytes);
  return cudaSuccess;
}
static inline cudaError_t cudaMemsetAsync(void *dst, int v, size_t bytes, cudaStream_t) {
  memset(dst, v, bytes);
  return cudaSuccess;
}
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaGetLastError(void) { return cudaSuccess; }
static inline const char *cudaGetErrorString(cudaError_t) { return "emulation"; }
