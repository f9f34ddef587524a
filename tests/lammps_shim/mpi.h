// Serial MPI stub for the test harness (same idea as LAMMPS' own STUBS/mpi.h):
// one rank, collectives are identities.  MPI is not installed in this image.
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
#define MPI_SUM 1
#define MPI_MAX 2
#define MPI_MIN 3
#define MPI_SUCCESS 0

static inline int MPI_Comm_rank(MPI_Comm, int *r) { *r = 0; return MPI_SUCCESS; }
static inline int MPI_Comm_size(MPI_Comm, int *s) { *s = 1; return MPI_SUCCESS; }
static inline int MPI_Allreduce(const void *in, void *out, int n, MPI_Datatype t, MPI_Op, MPI_Comm) {
  if (in != MPI_IN_PLACE) std::memcpy(out, in, (size_t)n * (t == MPI_DOUBLE ? sizeof(double) : sizeof(int)));
  return MPI_SUCCESS;
}
static inline int MPI_Bcast(void *, int, MPI_Datatype, int, MPI_Comm) { return MPI_SUCCESS; }
static inline int MPI_Barrier(MPI_Comm) { return MPI_SUCCESS; }
