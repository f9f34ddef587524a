// MPI stand-in for the test harness (same idea as LAMMPS' own STUBS/mpi.h): MPI is not installed in this image.
// By default one rank, collectives are identities.  A multi-process test plugs a transport in (shim_mpi(): function
// pointers the test fills, e.g. with torch.distributed over gloo) and the same calls then really communicate, so the
// host classes can be driven on several ranks.  Test infrastructure only.
#pragma once
#include <cstring>
#include <fstream>   // eph_fdm.h uses std::ifstream without including <fstream>

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
#define MPI_COMM_WORLD 0
#define MPI_IN_PLACE ((void *)1)
#define MPI_DOUBLE 1
#define MPI_INT 2
#define MPI_CHAR 3
#define MPI_BYTE 3
#define MPI_SUM 1
#define MPI_MAX 2
#define MPI_MIN 3
#define MPI_SUCCESS 0

struct ShimMpiBackend {
  int rank = 0, size = 1;
  void (*allreduce)(void *buf, int n, int dtype, int op) = nullptr;   // in place
  void (*bcast)(void *buf, int nbytes, int root) = nullptr;
  void (*alltoallv_int)(const int *send, const int *scount, const int *sdisp, int *recv, const int *rcount, const int *rdisp) = nullptr;
  void (*barrier)() = nullptr;
};
inline ShimMpiBackend &shim_mpi() {
  static ShimMpiBackend b;
  return b;
}
static inline size_t shim_mpi_size(MPI_Datatype t) { return t == MPI_DOUBLE ? sizeof(double) : t == MPI_INT ? sizeof(int) : 1; }

static inline int MPI_Comm_rank(MPI_Comm, int *r) { *r = shim_mpi().rank; return MPI_SUCCESS; }
static inline int MPI_Comm_size(MPI_Comm, int *s) { *s = shim_mpi().size; return MPI_SUCCESS; }
static inline int MPI_Allreduce(const void *in, void *out, int n, MPI_Datatype t, MPI_Op op, MPI_Comm) {
  if (in != MPI_IN_PLACE) std::memcpy(out, in, (size_t)n * shim_mpi_size(t));
  if (shim_mpi().size > 1) shim_mpi().allreduce(out, n, t, op);
  return MPI_SUCCESS;
}
static inline int MPI_Bcast(void *buf, int n, MPI_Datatype t, int root, MPI_Comm) {
  if (shim_mpi().size > 1) shim_mpi().bcast(buf, (int)((size_t)n * shim_mpi_size(t)), root);
  return MPI_SUCCESS;
}
static inline int MPI_Alltoallv(const void *send, const int *scount, const int *sdisp, MPI_Datatype, void *recv, const int *rcount,
                                const int *rdisp, MPI_Datatype, MPI_Comm) {
  if (shim_mpi().size > 1) shim_mpi().alltoallv_int(static_cast<const int *>(send), scount, sdisp, static_cast<int *>(recv), rcount, rdisp);
  else std::memcpy(static_cast<int *>(recv) + rdisp[0], static_cast<const int *>(send) + sdisp[0], sizeof(int) * (size_t)scount[0]);
  return MPI_SUCCESS;
}
static inline int MPI_Alltoall(const void *send, int n, MPI_Datatype t, void *recv, int, MPI_Datatype, MPI_Comm c) {
  const int size = shim_mpi().size;
  int cnt[64], dsp[64];   // the harness never runs more ranks than that
  for (int r = 0; r < size && r < 64; ++r) { cnt[r] = n; dsp[r] = r * n; }
  return MPI_Alltoallv(send, cnt, dsp, t, recv, cnt, dsp, t, c);
}
static inline int MPI_Barrier(MPI_Comm) {
  if (shim_mpi().size > 1 && shim_mpi().barrier) shim_mpi().barrier();
  return MPI_SUCCESS;
}
