// TEST INFRASTRUCTURE -- never part of the product, never loaded by it.
//
// Host stand-in for the slice of the CUDA runtime, driver types and device intrinsics that the product's sources under
// user-eph_b200/csrc use, so that the CPU test suite can compile THOSE VERY SOURCES for the host (tests/emul/Makefile:
// cu2cpp.py rewrites the launch syntax, this directory comes first on the include path so that <cuda_runtime.h> and
// <cub/cub.cuh> resolve here) and check kernels and orchestration against the oracle in a container without a GPU.
// Device memory is host memory, streams and events are immediate, kernels run under the lock-step SIMT stand-in of
// simt.h (real sub-warp shuffles, ballots, block barriers, shared memory).  TMA box loads are synchronous copies described by
// a tensor map kept in the clear and mbarrier waits are already satisfied (cu2cpp.py maps the two PTX statements).  Not
// modelled: memory ordering, asynchrony, timing.  What this cannot
// check is covered by the `-m gpu` tests on a B200.
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>

#include <fcntl.h>
#include <sched.h>
#include <sys/mman.h>
#include <time.h>
#include <unistd.h>

#include "simt.h"

#ifdef EPHA_EMUL_EXACT_ALLOC
#include <sanitizer/asan_interface.h>
#endif

#define EPHA_HOST_EMULATION 1

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __grid_constant__
#define __shared__ static
#define __align__(n) alignas(n)

struct double2 { double x, y; };
struct alignas(32) double4 { double x, y, z, w; };
struct alignas(16) float4 { float x, y, z, w; };
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
inline float __int_as_float(int v) { float f; std::memcpy(&f, &v, 4); return f; }
inline int __float_as_int(float f) { int v; std::memcpy(&v, &f, 4); return v; }
struct int2 { int x, y; };
struct int4 { int x, y, z, w; };
inline double2 make_double2(double x, double y) { return double2{x, y}; }
inline double4 make_double4(double x, double y, double z, double w) { return double4{x, y, z, w}; }
inline int2 make_int2(int x, int y) { return int2{x, y}; }
inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }

using std::max;
using std::min;

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
inline long long __double_as_longlong(double d) { long long u; std::memcpy(&u, &d, 8); return u; }
inline double __longlong_as_double(long long u) { double d; std::memcpy(&d, &u, 8); return d; }
inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
template <class T> inline T __ldcs(const T *p) { return *p; }
template <class T> inline T __ldg(const T *p) { return *p; }
template <class T> inline T __ldcg(const T *p) { return *p; }
template <class T> inline void __stcs(T *p, T v) { *p = v; }
inline size_t __cvta_generic_to_shared(const void *p) { return reinterpret_cast<size_t>(p); }
inline double rsqrt(double x) { return 1.0 / std::sqrt(x); }

// single host thread: atomics are plain read-modify-writes
template <class T> inline T atomicAdd(T *p, T v) { const T old = *p; *p = old + v; return old; }
template <class T> inline T atomicMax(T *p, T v) { const T old = *p; if (v > old) *p = v; return old; }
template <class T> inline T atomicMin(T *p, T v) { const T old = *p; if (v < old) *p = v; return old; }
template <class T> inline T atomicOr(T *p, T v) { const T old = *p; *p = old | v; return old; }
template <class T> inline T atomicCAS(T *p, T cmp, T v) { const T old = *p; if (old == cmp) *p = v; return old; }
template <class T> inline T atomicExch(T *p, T v) { const T old = *p; *p = v; return old; }

// ---- runtime ----
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorNotReady = 600, cudaErrorNotSupported = 801 };
typedef void *cudaStream_t;
typedef void *cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaEnableDefault = 0 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
enum cudaDriverEntryPointQueryResult { cudaDriverEntryPointSuccess = 0, cudaDriverEntryPointSymbolNotFound = 1 };
struct cudaDeviceProp {
  int major = 10, minor = 0, multiProcessorCount = 2;
  size_t sharedMemPerBlockOptin = 227 * 1024;
};
inline const char *cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "emulated CUDA error"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int) { *p = cudaDeviceProp(); return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = nullptr; return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = nullptr; return cudaSuccess; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { *e = nullptr; return cudaSuccess; }
inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventQuery(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }
template <class F> inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
template <class F> inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int *n, F, int, size_t) { *n = 2; return cudaSuccess; }
template <class T> inline cudaError_t cudaMalloc(T **p, size_t n) {
  // fresh storage is poisoned so that reads of unwritten device memory show up as NaNs / wild indices in the tests
#ifdef EPHA_EMUL_EXACT_ALLOC   // AddressSanitizer build: the poisoned zone starts right behind the last byte asked for
  *p = static_cast<T *>(std::aligned_alloc(32, (n + 31) / 32 * 32));
  if (*p && n % 32) ASAN_POISON_MEMORY_REGION(reinterpret_cast<char *>(*p) + n, 32 - n % 32);
#else
  *p = static_cast<T *>(std::aligned_alloc(256, (n + 255) / 256 * 256 + 256));
#endif
  if (!*p) return cudaErrorMemoryAllocation;
  std::memset(*p, 0xFF, n);
  return cudaSuccess;
}
template <class T> inline cudaError_t cudaMallocHost(T **p, size_t n) {
  *p = static_cast<T *>(std::malloc(n ? n : 1));
  return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
// ---- memory another process can map (cudaIpc*): POSIX shared-memory objects, so that the peer-memory ghost exchange
// (csrc/eph_p2p.cuh) runs between the CPU ranks of the multi-process tests.  emul_ipc_malloc stands in for the cudaMalloc
// of an exported allocation; the handle carries the object's name and size.
struct cudaIpcMemHandle_t { char reserved[64]; };
enum { cudaIpcMemLazyEnablePeerAccess = 1 };
struct EmulIpcRegion { std::string name; size_t bytes; bool owner; };
inline std::map<void *, EmulIpcRegion> &emul_ipc_regions() { static std::map<void *, EmulIpcRegion> m; return m; }
inline cudaError_t emul_ipc_malloc(void **p, size_t n) {
  static int counter = 0;
  char name[64];
  std::snprintf(name, sizeof name, "/ephb_emul_%d_%d", (int)getpid(), counter++);
  const int fd = shm_open(name, O_CREAT | O_EXCL | O_RDWR, 0600);
  if (fd < 0) return cudaErrorMemoryAllocation;
  if (ftruncate(fd, (off_t)n) != 0) { close(fd); shm_unlink(name); return cudaErrorMemoryAllocation; }
  void *q = mmap(nullptr, n, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  close(fd);
  if (q == MAP_FAILED) { shm_unlink(name); return cudaErrorMemoryAllocation; }
  emul_ipc_regions()[q] = EmulIpcRegion{name, n, true};
  *p = q;
  return cudaSuccess;
}
inline cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *h, void *p) {
  auto it = emul_ipc_regions().find(p);
  if (it == emul_ipc_regions().end() || !it->second.owner) return cudaErrorNotSupported;
  std::memset(h, 0, sizeof *h);
  std::strncpy(h->reserved, it->second.name.c_str(), 47);
  const unsigned long long n = it->second.bytes;
  std::memcpy(h->reserved + 48, &n, 8);
  return cudaSuccess;
}
inline cudaError_t cudaIpcOpenMemHandle(void **p, cudaIpcMemHandle_t h, unsigned) {
  unsigned long long n = 0;
  std::memcpy(&n, h.reserved + 48, 8);
  h.reserved[47] = 0;
  const int fd = shm_open(h.reserved, O_RDWR, 0600);
  if (fd < 0) return cudaErrorNotSupported;
  void *q = mmap(nullptr, (size_t)n, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  close(fd);
  if (q == MAP_FAILED) return cudaErrorNotSupported;
  emul_ipc_regions()[q] = EmulIpcRegion{h.reserved, (size_t)n, false};
  *p = q;
  return cudaSuccess;
}
inline cudaError_t cudaIpcCloseMemHandle(void *p) {
  auto it = emul_ipc_regions().find(p);
  if (it == emul_ipc_regions().end() || it->second.owner) return cudaErrorNotSupported;
  munmap(p, it->second.bytes);
  emul_ipc_regions().erase(it);
  return cudaSuccess;
}
inline cudaError_t cudaFree(void *p) {
  auto it = emul_ipc_regions().find(p);
  if (it != emul_ipc_regions().end()) {
    munmap(p, it->second.bytes);
    if (it->second.owner) shm_unlink(it->second.name.c_str());
    emul_ipc_regions().erase(it);
    return cudaSuccess;
  }
  std::free(p);
  return cudaSuccess;
}
inline cudaError_t cudaFreeHost(void *p) { std::free(p); return cudaSuccess; }
inline cudaError_t cudaMemset(void *p, int v, size_t n) { std::memset(p, v, n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void *p, int v, size_t n, cudaStream_t) { std::memset(p, v, n); return cudaSuccess; }
inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { std::memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t) { std::memmove(d, s, n); return cudaSuccess; }

// ---- the two driver entry points the engine asks for ----
typedef int CUresult;
enum { CUDA_SUCCESS = 0 };
typedef void *CUstream;
typedef unsigned long long CUdeviceptr;
typedef uint32_t cuuint32_t;
typedef uint64_t cuuint64_t;
enum CUtensorMapDataType { CU_TENSOR_MAP_DATA_TYPE_FLOAT64 = 10 };
enum CUtensorMapInterleave { CU_TENSOR_MAP_INTERLEAVE_NONE = 0 };
enum CUtensorMapSwizzle { CU_TENSOR_MAP_SWIZZLE_NONE = 0 };
enum CUtensorMapL2promotion { CU_TENSOR_MAP_L2_PROMOTION_L2_128B = 2 };
enum CUtensorMapFloatOOBfill { CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE = 0 };
enum { CU_STREAM_WAIT_VALUE_GEQ = 0 };
// a tiled tensor map: what cuTensorMapEncodeTiled is told, kept in the clear (the real one is an opaque 128-byte blob)
struct alignas(64) CUtensorMap {
  unsigned char *base;
  uint64_t dims[3], strides[2];   // elements; bytes between rows / planes
  uint32_t box[3];
  uint32_t elem_bytes, rank;
  unsigned char pad[128 - 8 - 24 - 16 - 12 - 8];
};
inline CUresult emul_cuTensorMapEncodeTiled(CUtensorMap *map, CUtensorMapDataType dt, cuuint32_t rank, void *base, const cuuint64_t *dims,
                                            const cuuint64_t *strides, const cuuint32_t *box, const cuuint32_t *, CUtensorMapInterleave,
                                            CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill) {
  if (rank != 3 || dt != CU_TENSOR_MAP_DATA_TYPE_FLOAT64) return 1;
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (strides[0] & 15) || (strides[1] & 15)) return 1;   // the real encoder's alignment rules
  std::memset(map, 0, sizeof *map);
  map->base = static_cast<unsigned char *>(base);
  for (int d = 0; d < 3; ++d) { map->dims[d] = dims[d]; map->box[d] = box[d]; }
  map->strides[0] = strides[0]; map->strides[1] = strides[1];
  map->elem_bytes = 8; map->rank = 3;
  return CUDA_SUCCESS;
}
// cp.async.bulk.tensor.3d: the box starting at (c0, c1, c2), row-major in shared memory, elements outside the tensor zero
inline long long g_emul_tma_loads = 0;   // read by the tests through emul_tma_load_count() (emul_probe.cpp)
inline void emul_tma_load_3d(void *dst, const CUtensorMap *map, int c0, int c1, int c2) {
  ++g_emul_tma_loads;
  if (c0 & 1) { std::fprintf(stderr, "emulated TMA: fp64 box starts at odd x = %d (16-byte rule)\n", c0); std::abort(); }
  double *out = static_cast<double *>(dst);
  for (uint32_t z = 0; z < map->box[2]; ++z)
    for (uint32_t y = 0; y < map->box[1]; ++y)
      for (uint32_t x = 0; x < map->box[0]; ++x) {
        const long long gx = (long long)c0 + x, gy = (long long)c1 + y, gz = (long long)c2 + z;
        double v = 0.0;
        if (gx >= 0 && gy >= 0 && gz >= 0 && gx < (long long)map->dims[0] && gy < (long long)map->dims[1] && gz < (long long)map->dims[2])
          std::memcpy(&v, map->base + gx * 8 + gy * map->strides[0] + gz * map->strides[1], 8);
        *out++ = v;
      }
}
// streams are immediate: whatever a wait would wait for has already happened
inline CUresult emul_cuStreamWaitValue32(CUstream, CUdeviceptr, cuuint32_t, unsigned) { return CUDA_SUCCESS; }
inline cudaError_t cudaGetDriverEntryPoint(const char *name, void **p, unsigned long long, cudaDriverEntryPointQueryResult *q = nullptr) {
  *p = nullptr;
  if (std::strcmp(name, "cuTensorMapEncodeTiled") == 0 && !std::getenv("EPH_EMUL_NO_TMA")) *p = reinterpret_cast<void *>(&emul_cuTensorMapEncodeTiled);
  if (std::strcmp(name, "cuStreamWaitValue32") == 0) *p = reinterpret_cast<void *>(&emul_cuStreamWaitValue32);
  if (q) *q = *p ? cudaDriverEntryPointSuccess : cudaDriverEntryPointSymbolNotFound;
  return *p ? cudaSuccess : cudaErrorNotSupported;
}
