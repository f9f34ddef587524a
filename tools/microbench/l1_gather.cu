// L1TEX cost model probe for sm_100a: cycles per warp-wide load request for the access patterns the sweeps could use.
// Working set is L1-resident (64 KB of 32-byte records); every SM runs the same loop; we report SM cycles per request.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ double4 ld256(const double4* p){ double4 r; asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.x),"=d"(r.y),"=d"(r.z),"=d"(r.w) : "l"(p)); return r; }

// mode 0: LDG.256 random distinct   1: LDG.256 4 lanes share   2: LDG.256 contiguous   3: 2x LDG.128 random
// mode 4: LDG.256 8 lanes share     5: LDS.128x2 random (split arrays)   6: LDS.128x2, 4 lanes share   7: LDG.64 x4 random SoA
// mode 8: LDG.256 random within runs of 4 consecutive records (pairs of lanes adjacent)  9: LDS.64 x4 random
template<int MODE>
__global__ void probe(const double4* __restrict__ rec, const int* __restrict__ idx, int nrec, int iters, double* out, long long* cyc){
  extern __shared__ double2 sm[];
  double2* s_lo = sm; double2* s_hi = sm + nrec;
  for (int t = threadIdx.x; t < nrec; t += blockDim.x){ double4 r = rec[t]; s_lo[t] = make_double2(r.x, r.y); s_hi[t] = make_double2(r.z, r.w); }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  double acc = 0;
  const double* soa = reinterpret_cast<const double*>(rec);
  long long t0 = clock64();
  int cursor = (blockIdx.x * blockDim.x + threadIdx.x) * 7;
  for (int it = 0; it < iters; ++it){
    int k = idx[(cursor + it * 32) & (1<<20)-1];      // random record id, coalesced index load
    if (MODE == 1 || MODE == 6) k = __shfl_sync(0xffffffffu, k, lane & ~3);
    if (MODE == 4) k = __shfl_sync(0xffffffffu, k, lane & ~7);
    if (MODE == 2) k = (__shfl_sync(0xffffffffu, k, 0) + lane) % nrec;
    if (MODE == 8) k = ((__shfl_sync(0xffffffffu, k, lane & ~3) & ~3) + (lane & 3)) % nrec;
    if (MODE == 0 || MODE == 1 || MODE == 2 || MODE == 4 || MODE == 8){ double4 r = ld256(rec + k); acc += r.x + r.y + r.z + r.w; }
    else if (MODE == 3){ const double2* p = reinterpret_cast<const double2*>(rec + k); double2 a = __ldg(p), b = __ldg(p+1); acc += a.x+a.y+b.x+b.y; }
    else if (MODE == 5 || MODE == 6){ double2 a = s_lo[k], b = s_hi[k]; acc += a.x+a.y+b.x+b.y; }
    else if (MODE == 7){ acc += __ldg(soa + k) + __ldg(soa + nrec + k) + __ldg(soa + 2*nrec + k) + __ldg(soa + 3*nrec + k); }
    else if (MODE == 9){ const double* q = reinterpret_cast<const double*>(sm); acc += q[k] + q[nrec + k] + q[2*nrec + k] + q[3*nrec + k]; }
  }
  long long t1 = clock64();
  if (acc == 12345.678) out[0] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template<int MODE> void run(const char* name, const double4* rec, const int* idx, int nrec, double* out, long long* cyc, int nsm){
  const int iters = 4096, threads = 512;
  size_t smem = (size_t)nrec * 32;
  cudaFuncSetAttribute(probe<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  probe<MODE><<<nsm, threads, smem>>>(rec, idx, nrec, 64, out, cyc);
  probe<MODE><<<nsm, threads, smem>>>(rec, idx, nrec, iters, out, cyc);
  cudaDeviceSynchronize();
  long long h[256]; cudaMemcpy(h, cyc, nsm*sizeof(long long), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < nsm; ++i) avg += h[i]; avg /= nsm;
  double req = (double)iters * (threads/32);   // warp requests per SM (per logical record load)
  printf("%-44s %8.2f SM-cycles per warp-wide record load (32 lanes x 32 B)   err=%s\n", name, avg/req, cudaGetErrorString(cudaGetLastError()));
}

int main(){
  const int nrec = 2048; int nsm = 148;
  double4* rec; int* idx; double* out; long long* cyc;
  cudaMalloc(&rec, nrec*32); cudaMalloc(&idx, (1<<20)*4); cudaMalloc(&out, 8); cudaMalloc(&cyc, 256*8);
  int* h = (int*)malloc((1<<20)*4); srand(1); for (int i = 0; i < (1<<20); ++i) h[i] = rand() % nrec;
  cudaMemcpy(idx, h, (1<<20)*4, cudaMemcpyHostToDevice); cudaMemset(rec, 0, nrec*32);
  run<0>("LDG.256 random (32 distinct sectors)", rec, idx, nrec, out, cyc, nsm);
  run<8>("LDG.256 random runs of 4 (8 lines)", rec, idx, nrec, out, cyc, nsm);
  run<1>("LDG.256 4 lanes share (8 distinct sectors)", rec, idx, nrec, out, cyc, nsm);
  run<4>("LDG.256 8 lanes share (4 distinct sectors)", rec, idx, nrec, out, cyc, nsm);
  run<2>("LDG.256 contiguous 1024 B", rec, idx, nrec, out, cyc, nsm);
  run<3>("2 x LDG.128 random", rec, idx, nrec, out, cyc, nsm);
  run<7>("4 x LDG.64 random (SoA)", rec, idx, nrec, out, cyc, nsm);
  run<5>("2 x LDS.128 random (split arrays)", rec, idx, nrec, out, cyc, nsm);
  run<6>("2 x LDS.128, 4 lanes share", rec, idx, nrec, out, cyc, nsm);
  run<9>("4 x LDS.64 random (SoA)", rec, idx, nrec, out, cyc, nsm);
  return 0;
}
