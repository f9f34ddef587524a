// Does the texture return path add gather bandwidth on top of the LSU data pipe (sm_100a)?
// Records are 32 B; a warp-wide "record load" is 32 lanes x 32 B at random record ids (L1-resident working set).
// Modes: 0 one record via LDG.256; 1 two records via LDG.256; 2 one via LDG.256 + one via 2 x tex1Dfetch<int4>;
//        3 one record via 2 x tex1Dfetch<int4>; 4 three records via LDG.256; 5 two via LDG.256 + one via tex;
//        6 one via LDG.256 + one via 2 x LDS.128 (shared copy)
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ double4 ld256(const double4* p){ double4 r; asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.x),"=d"(r.y),"=d"(r.z),"=d"(r.w) : "l"(p)); return r; }
__device__ __forceinline__ double sum4(int4 a){ return __hiloint2double(a.y, a.x) + __hiloint2double(a.w, a.z); }

template<int MODE>
__global__ void probe(const double4* __restrict__ rec, cudaTextureObject_t tex, const int* __restrict__ idx, int nrec, int iters, double* out, long long* cyc){
  extern __shared__ double2 sm[];
  if (MODE == 6) { for (int t = threadIdx.x; t < nrec; t += blockDim.x){ double4 r = rec[t]; sm[t] = make_double2(r.x, r.y); sm[nrec + t] = make_double2(r.z, r.w); } __syncthreads(); }
  double acc = 0;
  long long t0 = clock64();
  int cursor = (blockIdx.x * blockDim.x + threadIdx.x) * 7;
  for (int it = 0; it < iters; ++it){
    const int base = (cursor + it * 96) & ((1<<20)-1);
    int k1 = idx[base], k2 = idx[(base + 32) & ((1<<20)-1)], k3 = idx[(base + 64) & ((1<<20)-1)];
    if (MODE == 0 || MODE == 1 || MODE == 2 || MODE == 4 || MODE == 5 || MODE == 6){ double4 r = ld256(rec + k1); acc += r.x + r.y + r.z + r.w; }
    if (MODE == 1 || MODE == 4 || MODE == 5){ double4 r = ld256(rec + k2); acc += r.x + r.y + r.z + r.w; }
    if (MODE == 4){ double4 r = ld256(rec + k3); acc += r.x + r.y + r.z + r.w; }
    if (MODE == 2 || MODE == 3){ int4 a = tex1Dfetch<int4>(tex, 2 * k2), b = tex1Dfetch<int4>(tex, 2 * k2 + 1); acc += sum4(a) + sum4(b); }
    if (MODE == 5){ int4 a = tex1Dfetch<int4>(tex, 2 * k3), b = tex1Dfetch<int4>(tex, 2 * k3 + 1); acc += sum4(a) + sum4(b); }
    if (MODE == 6){ double2 a = sm[k2], b = sm[nrec + k2]; acc += a.x + a.y + b.x + b.y; }
  }
  long long t1 = clock64();
  if (acc == 12345.678) out[0] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template<int MODE> void run(const char* name, const double4* rec, cudaTextureObject_t tex, const int* idx, int nrec, double* out, long long* cyc, int nsm, int threads){
  const int iters = 4096;
  size_t smem = MODE == 6 ? (size_t)nrec * 32 : 0;
  if (smem) cudaFuncSetAttribute(probe<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  probe<MODE><<<nsm, threads, smem>>>(rec, tex, idx, nrec, 64, out, cyc);
  probe<MODE><<<nsm, threads, smem>>>(rec, tex, idx, nrec, iters, out, cyc);
  cudaDeviceSynchronize();
  long long h[256]; cudaMemcpy(h, cyc, nsm*sizeof(long long), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < nsm; ++i) avg += h[i]; avg /= nsm;
  double req = (double)iters * (threads/32);
  printf("%-58s threads/SM %4d  %8.2f SM-cycles per warp iteration   err=%s\n", name, threads, avg/req, cudaGetErrorString(cudaGetLastError()));
}

int main(){
  const int nrec = 2048; int nsm = 148;
  double4* rec; int* idx; double* out; long long* cyc;
  cudaMalloc(&rec, nrec*32); cudaMalloc(&idx, (1<<20)*4); cudaMalloc(&out, 8); cudaMalloc(&cyc, 256*8);
  int* h = (int*)malloc((1<<20)*4); srand(1); for (int i = 0; i < (1<<20); ++i) h[i] = rand() % nrec;
  cudaMemcpy(idx, h, (1<<20)*4, cudaMemcpyHostToDevice); cudaMemset(rec, 0, nrec*32);
  cudaResourceDesc rd = {}; rd.resType = cudaResourceTypeLinear; rd.res.linear.devPtr = rec;
  rd.res.linear.desc = cudaCreateChannelDesc<int4>(); rd.res.linear.sizeInBytes = nrec * 32;
  cudaTextureDesc td = {}; td.readMode = cudaReadModeElementType;
  cudaTextureObject_t tex = 0; cudaCreateTextureObject(&tex, &rd, &td, nullptr);
  // L1 hit vs L1 miss (L2 hit): the same gather over working sets of 64 KB ... 8 MB
  for (int lg = 11; lg <= 18; ++lg) {
    const int n2 = 1 << lg;
    double4* rec2; cudaMalloc(&rec2, (size_t)n2 * 32); cudaMemset(rec2, 0, (size_t)n2 * 32);
    for (int i = 0; i < (1<<20); ++i) h[i] = rand() % n2;
    cudaMemcpy(idx, h, (1<<20)*4, cudaMemcpyHostToDevice);
    char name[96]; snprintf(name, sizeof name, "3 records LDG.256, working set %d KB", n2 / 32);
    run<4>(name, rec2, tex, idx, n2, out, cyc, nsm, 1024);
    cudaFree(rec2);
  }
  for (int i = 0; i < (1<<20); ++i) h[i] = rand() % nrec;
  cudaMemcpy(idx, h, (1<<20)*4, cudaMemcpyHostToDevice);
  for (int threads : {1024}) {
    run<0>("1 record LDG.256", rec, tex, idx, nrec, out, cyc, nsm, threads);
    run<1>("2 records LDG.256", rec, tex, idx, nrec, out, cyc, nsm, threads);
    run<4>("3 records LDG.256", rec, tex, idx, nrec, out, cyc, nsm, threads);
    run<3>("1 record 2 x tex1Dfetch<int4>", rec, tex, idx, nrec, out, cyc, nsm, threads);
    run<2>("1 record LDG.256 + 1 record tex", rec, tex, idx, nrec, out, cyc, nsm, threads);
    run<5>("2 records LDG.256 + 1 record tex", rec, tex, idx, nrec, out, cyc, nsm, threads);
    run<6>("1 record LDG.256 + 1 record 2 x LDS.128", rec, tex, idx, nrec, out, cyc, nsm, threads);
  }
  return 0;
}
