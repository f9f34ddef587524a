// TEST INFRASTRUCTURE -- never part of the product, never loaded by it.
//
// A serial stand-in for the slice of the CUDA runtime and device intrinsics that user-eph_b200/csrc/eph_atomic.cu uses,
// so that the CPU test suite can compile THAT VERY SOURCE for the host (tests/emul/Makefile, -DEPHA_HOST_EMULATION,
// this directory first on the include path so that <cuda_runtime.h> resolves here) and check the kernels' logic and
// the C-ABI orchestration against the oracle in a container without a GPU.  Semantics: device memory is host memory,
// streams are immediate, a kernel launch runs its threads one after the other (the kernels of that file use no shared
// memory and no block-level synchronisation), one lane per atom (kLanes = 1), cross-lane shuffles see no other lane.
// What this cannot check -- real sub-warp shuffles, memory ordering, launch configuration limits -- is covered by the
// `-m gpu` tests on a B200.
#pragma once

#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)

struct double2 { double x, y; };
struct double4 { double x, y, z, w; };
inline double2 make_double2(double x, double y) { return double2{x, y}; }
inline double4 make_double4(double x, double y, double z, double w) { return double4{x, y, z, w}; }

struct uint3_emul { unsigned x = 0, y = 0, z = 0; };
namespace epha_emul {
inline thread_local uint3_emul threadIdx_, blockIdx_, blockDim_, gridDim_;
template <class F>
void launch(int grid, int block, F &&body) {
  gridDim_.x = (unsigned)grid;
  blockDim_.x = (unsigned)block;
  for (int b = 0; b < grid; ++b)
    for (int t = 0; t < block; ++t) {
      blockIdx_.x = (unsigned)b;
      threadIdx_.x = (unsigned)t;
      body();
    }
}
}  // namespace epha_emul
#define threadIdx epha_emul::threadIdx_
#define blockIdx epha_emul::blockIdx_
#define blockDim epha_emul::blockDim_
#define gridDim epha_emul::gridDim_

// ---- device intrinsics ----
inline double __hiloint2double(int hi, int lo) {
  const uint64_t u = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo;
  double d;
  std::memcpy(&d, &u, 8);
  return d;
}
inline int __double2loint(double d) {
  uint64_t u;
  std::memcpy(&u, &d, 8);
  return (int)(uint32_t)u;
}
inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
template <class T> inline T __ldcs(const T *p) { return *p; }
template <class T> inline void __stcs(T *p, T v) { *p = v; }
// no other lane exists in the serial stand-in: a shuffle contributes nothing to a sum
inline double __shfl_xor_sync(unsigned, double, int) { return 0.0; }
inline double atomicAdd(double *p, double v) { const double old = *p; *p = old + v; return old; }

// ---- runtime ----
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2 };
typedef void *cudaStream_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
enum { cudaStreamNonBlocking = 1 };
struct cudaDeviceProp { int major = 10, minor = 0, multiProcessorCount = 4; };
inline const char *cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "emulated error"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int) { *p = cudaDeviceProp(); return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = nullptr; return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
template <class T> inline cudaError_t cudaMalloc(T **p, size_t n) {
  // poison fresh storage so that reads of unwritten device memory show up as NaNs in the tests
  *p = static_cast<T *>(std::malloc(n ? n : 1));
  if (!*p) return cudaErrorMemoryAllocation;
  std::memset(*p, 0xFF, n);
  return cudaSuccess;
}
template <class T> inline cudaError_t cudaMallocHost(T **p, size_t n) { *p = static_cast<T *>(std::malloc(n ? n : 1)); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
inline cudaError_t cudaFree(void *p) { std::free(p); return cudaSuccess; }
inline cudaError_t cudaFreeHost(void *p) { std::free(p); return cudaSuccess; }
inline cudaError_t cudaMemset(void *p, int v, size_t n) { std::memset(p, v, n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void *p, int v, size_t n, cudaStream_t) { std::memset(p, v, n); return cudaSuccess; }
inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { std::memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t) { std::memmove(d, s, n); return cudaSuccess; }
