// TEST INFRASTRUCTURE: self-test of the lock-step SIMT stand-in (simt.h).  Built through cu2cpp.py like the product's
// sources.  `selftest ok` runs kernels whose results are known in closed form (block reduction through shared memory and
// __syncthreads with early-exiting threads, sub-warp shuffles under group masks with different trip counts per group,
// ballots, a 3-D launch, dynamic shared memory); `selftest deadlock` runs a kernel whose lanes name a mask that one of
// them never joins and must be stopped by the dead-lock report (exit through abort); `selftest race` runs a kernel that
// misses a barrier: its checksum must differ between SIMT_ORDER settings, which is how the host build exposes races.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <vector>

__global__ void block_sum_kernel(int n, const double *in, double *out) {
  __shared__ double part[8];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  double v = i < n ? in[i] : 0.0;
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x >= 32) return;                     // seven warps leave; the first one goes on to another barrier-free phase
  double t = threadIdx.x < (int)(blockDim.x >> 5) ? part[threadIdx.x] : 0.0;
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xFFFFFFFFu, t, o);
  if (threadIdx.x == 0) atomicAdd(out, t);
}

// groups of 4 lanes; group g sums g + 1 rounds of its lane ids, so groups of one warp shuffle different numbers of times
__global__ void group_kernel(int *out) {
  const int lane = threadIdx.x & 31, sub = lane & 3, group = threadIdx.x >> 2;
  const unsigned gmask = 0xFu << (lane & ~3);
  int acc = 0;
  for (int r = 0; r <= group % 5; ++r) {
    int v = sub + r;
    v += __shfl_xor_sync(gmask, v, 1);
    v += __shfl_xor_sync(gmask, v, 2);
    acc += v;
  }
  const unsigned b = __ballot_sync(gmask, sub >= 2);   // lanes 2, 3 of the group
  if (sub == 0) out[group] = acc * 16 + (int)((b >> (lane & ~3)) & 0xFu);
}

__global__ void dyn_smem_3d_kernel(int nx, int ny, int nz, int *out) {
  extern __shared__ int tile[];
  const int t = threadIdx.x + threadIdx.y * blockDim.x + threadIdx.z * blockDim.x * blockDim.y;
  const int n = blockDim.x * blockDim.y * blockDim.z;
  tile[t] = t;
  __syncthreads();
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y, z = blockIdx.z * blockDim.z + threadIdx.z;
  if (x < nx && y < ny && z < nz) out[x + nx * (y + ny * z)] = tile[n - 1 - t] + 1000 * (int)(blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z));
}

// a missing barrier: thread t reads its neighbour's slot without waiting for the neighbour to write it
__global__ void racy_kernel(int *out) {
  __shared__ int slot[64];
  slot[threadIdx.x] = 0;
  __syncthreads();
  slot[threadIdx.x] = 100 + threadIdx.x;
  // __syncthreads() belongs here
  out[threadIdx.x] = slot[(threadIdx.x + 1) % 64];
}

__global__ void bad_mask_kernel(int *out) {
  const int lane = threadIdx.x & 31;
  int v = lane;
  if (lane != 5) v += __shfl_xor_sync(0xFFFFFFFFu, v, 1);   // lane 5 is named in the mask but never arrives
  out[threadIdx.x] = v;
}

int main(int argc, char **argv) {
  if (argc > 1 && std::strcmp(argv[1], "deadlock") == 0) {
    std::vector<int> out(32);
    bad_mask_kernel<<<1, 32>>>(out.data());
    std::printf("not reached\n");
    return 0;
  }
  if (argc > 1 && std::strcmp(argv[1], "race") == 0) {   // the sum depends on the order the fibers are resumed in (SIMT_ORDER)
    std::vector<int> out(64);
    racy_kernel<<<1, 64>>>(out.data());
    long long sum = 0;
    for (int v : out) sum += v;
    std::printf("race checksum %lld\n", sum);
    return 0;
  }
  int bad = 0;
  {
    const int n = 1000;
    std::vector<double> in(n);
    double want = 0.0, got = 0.0;
    for (int i = 0; i < n; ++i) { in[i] = i + 1; want += in[i]; }
    block_sum_kernel<<<(n + 255) / 256, 256>>>(n, in.data(), &got);
    if (got != want) { std::printf("block_sum: %g != %g\n", got, want); ++bad; }
  }
  {
    std::vector<int> out(64, -1);
    group_kernel<<<1, 256>>>(out.data());
    for (int g = 0; g < 64; ++g) {
      int acc = 0;
      for (int r = 0; r <= g % 5; ++r) acc += 6 + 4 * r;   // sum over the group's four lanes of (sub + r)
      if (out[g] != acc * 16 + 0xC) { std::printf("group %d: %d != %d\n", g, out[g], acc * 16 + 0xC); ++bad; }
    }
  }
  {
    const int nx = 5, ny = 3, nz = 4;
    std::vector<int> out(nx * ny * nz, -1);
    dim3 block(4, 2, 2), grid((nx + 3) / 4, (ny + 1) / 2, (nz + 1) / 2);
    dyn_smem_3d_kernel<<<grid, block, 16 * sizeof(int)>>>(nx, ny, nz, out.data());
    for (int z = 0; z < nz; ++z)
      for (int y = 0; y < ny; ++y)
        for (int x = 0; x < nx; ++x) {
          const int t = (x % 4) + 4 * ((y % 2) + 2 * (z % 2));
          const int b = (x / 4) + 2 * ((y / 2) + 2 * (z / 2));
          if (out[x + nx * (y + ny * z)] != 15 - t + 1000 * b) { std::printf("3d (%d,%d,%d)\n", x, y, z); ++bad; }
        }
  }
  std::printf(bad ? "selftest FAILED\n" : "selftest ok\n");
  return bad ? 1 : 0;
}
