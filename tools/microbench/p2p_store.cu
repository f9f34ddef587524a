// What a kernel that stores rows into a PEER GPU's memory over NVLink costs, piece by piece (development probe for
// csrc/eph_p2p.cuh; two GPUs of one box, one process):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o p2p_store p2p_store.cu && ./p2p_store
// Variants: local destination / peer destination; no fence / one __threadfence_system per block / per thread; rows of 32 B.
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

template <int FENCE>   // 0 none, 1 one per block (thread 0 after a barrier), 2 every thread
__global__ void __launch_bounds__(256) store_rows(int n, const int *__restrict__ index, const double4 *__restrict__ src, double4 *dst,
                                                  unsigned *done, unsigned long long *flag, unsigned long long epoch) {
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) dst[t] = src[index[t]];
  if (FENCE == 2) __threadfence_system();
  __shared__ bool last;
  __syncthreads();
  if (threadIdx.x == 0) {
    if (FENCE == 1) __threadfence_system();
    last = atomicAdd(done, 1u) == gridDim.x - 1;
    if (last) *done = 0u;
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    if (FENCE) __threadfence_system();
    *reinterpret_cast<volatile unsigned long long *>(flag) = epoch;
  }
}

int main() {
  int ndev = 0;
  CK(cudaGetDeviceCount(&ndev));
  if (ndev < 2) { printf("needs two GPUs\n"); return 0; }
  CK(cudaSetDevice(0));
  CK(cudaDeviceEnablePeerAccess(1, 0));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int natoms = 620000;
  double4 *src, *loc, *peer;
  unsigned *done;
  unsigned long long *flag_loc, *flag_peer;
  int *index;
  CK(cudaMalloc(&src, natoms * sizeof(double4)));
  CK(cudaMemset(src, 0, natoms * sizeof(double4)));
  CK(cudaMalloc(&loc, 400000 * sizeof(double4)));
  CK(cudaMalloc(&done, 4)); CK(cudaMemset(done, 0, 4));
  CK(cudaMalloc(&flag_loc, 8));
  CK(cudaSetDevice(1));
  CK(cudaMalloc(&peer, 400000 * sizeof(double4)));
  CK(cudaMalloc(&flag_peer, 8));
  CK(cudaSetDevice(0));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int n : {1000, 30000, 120000, 300000}) {
    std::vector<int> h(n);
    for (int k = 0; k < n; ++k) h[k] = (int)((1103515245ull * k + 12345ull) % natoms);
    CK(cudaMalloc(&index, n * sizeof(int)));
    CK(cudaMemcpy(index, h.data(), n * sizeof(int), cudaMemcpyHostToDevice));
    for (int where = 0; where < 2; ++where) {
      double4 *dst = where ? peer : loc;
      unsigned long long *flag = where ? flag_peer : flag_loc;
      for (int fence = 0; fence < 3; ++fence) {
        for (int blocks : {2 * prop.multiProcessorCount, (n + 255) / 256}) {
          blocks = blocks < 1 ? 1 : blocks;
          if (blocks > (n + 255) / 256) blocks = (n + 255) / 256;
          float best = 1e9f;
          for (int rep = 0; rep < 12; ++rep) {
            CK(cudaEventRecord(e0));
            if (fence == 0) store_rows<0><<<blocks, 256>>>(n, index, src, dst, done, flag, rep + 1);
            if (fence == 1) store_rows<1><<<blocks, 256>>>(n, index, src, dst, done, flag, rep + 1);
            if (fence == 2) store_rows<2><<<blocks, 256>>>(n, index, src, dst, done, flag, rep + 1);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (rep >= 2 && ms < best) best = ms;
          }
          printf("rows %6d  %-5s  fence %-10s  blocks %5d  %7.1f us  (%.0f GB/s)\n", n, where ? "peer" : "local",
                 fence == 0 ? "none" : fence == 1 ? "per block" : "per thread", blocks, best * 1e3, n * 32.0 / (best * 1e-3) / 1e9);
        }
      }
    }
    CK(cudaFree(index));
  }
  return 0;
}
