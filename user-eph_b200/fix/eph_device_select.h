// Default CUDA device of a rank for the `fix .../b200` host classes: the node-local rank modulo the number of visible
// devices.  Inside LAMMPS the node-local rank comes from the launcher's environment (Open MPI, MVAPICH2, Slurm, torchrun);
// without any of those variables the world rank is used, which is right for one node.
#pragma once

#include <cstdlib>

#include "eph_b200.h"

namespace eph_b200 {

inline int default_device(int world_rank) {
  int local = world_rank;
  for (const char *name : {"OMPI_COMM_WORLD_LOCAL_RANK", "MV2_COMM_WORLD_LOCAL_RANK", "MPI_LOCALRANKID", "SLURM_LOCALID", "LOCAL_RANK"}) {
    const char *v = std::getenv(name);
    if (v && *v) { local = std::atoi(v); break; }
  }
  int ndev = 0;
  if (eph_b200_device_count(&ndev) != EPH_B200_OK || ndev < 1) return 0;   // create() reports the missing device
  return local % ndev;
}

}  // namespace eph_b200
