// Which 3-D fp64 TMA boxes does sm_100a accept?  (development probe for eph_grid_tma.cuh)
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ unsigned s32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__global__ void probe(const __grid_constant__ CUtensorMap map, int c0, int c1, int c2, int bytes, double *out, int n) {
  extern __shared__ __align__(128) unsigned char sm[];
  double *dst = reinterpret_cast<double *>(sm);
  unsigned long long *bar = reinterpret_cast<unsigned long long *>(sm + 65536);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(s32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(s32(bar)), "r"((unsigned)bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 :: "r"(s32(dst)), "l"(&map), "r"(c0), "r"(c1), "r"(c2), "r"(s32(bar)) : "memory");
  }
  unsigned done = 0;
  while (!done) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(s32(bar)), "r"(0u) : "memory");
  for (int t = threadIdx.x; t < n; t += blockDim.x) out[t] = dst[t];
}
int main() {
  void *p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)p;
  const int N = 64; double *g, *out; cudaMalloc(&g, sizeof(double) * N * N * N); cudaMalloc(&out, 65536);
  double *h = new double[N * N * N]; for (int i = 0; i < N * N * N; ++i) h[i] = i; cudaMemcpy(g, h, sizeof(double) * N * N * N, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 64);
  int boxes[][3] = {{36, 10, 10}, {34, 10, 10}};
  int coords[][3] = {{0, 0, 0}, {-2, -1, -1}, {30, 23, 7}, {62, 63, 63}, {-2, 0, 0}, {0, -1, 0}, {1, 0, 0}, {-1, 0, 0}};
  for (auto &b : boxes) for (auto &c : coords) {
    CUtensorMap map; cuuint64_t dims[3] = {N, N, N}, str[2] = {N * 8, N * N * 8}; cuuint32_t box[3] = {(cuuint32_t)b[0], (cuuint32_t)b[1], (cuuint32_t)b[2]}, es[3] = {1, 1, 1};
    CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, g, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    int n = b[0] * b[1] * b[2];
    if (r != CUDA_SUCCESS) { printf("box %dx%dx%d encode error %d\n", b[0], b[1], b[2], (int)r); break; }
    probe<<<1, 128, 65536 + 64>>>(map, c[0], c[1], c[2], n * 8, out, n);
    cudaError_t e = cudaDeviceSynchronize();
    double first = -1, last = -1;
    if (e == cudaSuccess) { cudaMemcpy(&first, out, 8, cudaMemcpyDeviceToHost); cudaMemcpy(&last, out + n - 1, 8, cudaMemcpyDeviceToHost); }
    printf("box %2dx%2dx%2d (%5d B) at (%2d,%2d,%2d): %s first=%g last=%g\n", b[0], b[1], b[2], n * 8, c[0], c[1], c[2], cudaGetErrorString(e), first, last);
    if (e != cudaSuccess) return 0;   // context is dead after an illegal instruction
  }
  return 0;
}
