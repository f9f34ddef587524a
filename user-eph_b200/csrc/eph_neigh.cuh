// Device-side construction of the full neighbour list (cut-off r_c + skin) from the positions, as an alternative to
// uploading LAMMPS' list (2.2 GB at 4 M atoms): cell binning with half-cut-off cells, counting pass, exclusive scan,
// fill pass.  Produces the same pair set LAMMPS' `REQ_FULL` list would (rows of local atoms over locals and ghosts).
#pragma once

#include "eph_device.cuh"

namespace ephb {

// Capacity of every warp tile of the inner list (eph_sweeps.cuh): 32 entries per iteration, as many iterations as
// the longest LAMMPS row of the tile needs (the inner list is a subset of LAMMPS' list), rounded up to a multiple of four.
// caps[ntiles] = 0 closes the scan.
__global__ void tile_caps_kernel(int nlocal, const long long *__restrict__ offsets, int tile_atoms, int lanes,
                                 int ntiles, long long *__restrict__ caps) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w > ntiles) return;
  long long longest = 0;
  if (w < ntiles)
    for (int a = 0; a < tile_atoms; ++a) {
      const int i = w * tile_atoms + a;
      if (i < nlocal) longest = max(longest, offsets[i + 1] - offsets[i]);
    }
  // a multiple of four iterations: the packed sweeps pad every list to whole blocks of up to four (eph_packed.cuh)
  caps[w] = 32 * (((longest + lanes - 1) / lanes + 3) / 4 * 4);
}

// Boundary-first ordering of the density pass (multi-rank overlap): tiles that hold an atom some other rank needs as a
// ghost are flagged, then split into two ordered work lists with an exclusive scan of the flags.
__global__ void tile_mark_kernel(int n, const int *__restrict__ index, int tile_atoms, int nlocal, int *__restrict__ flag) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const int a = index[t];
  if (a >= 0 && a < nlocal) flag[a / tile_atoms] = 1;
}
__global__ void tile_split_kernel(int ntiles, const int *__restrict__ flag, const int *__restrict__ scan,
                                  int *__restrict__ work /* boundary tiles, then interior tiles */) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ntiles) return;
  if (flag[t]) work[scan[t]] = t;
  else work[scan[ntiles] + t - scan[t]] = t;
}

struct CellGrid {
  double lo[3];
  double inv[3];   // cells per length
  int nb[3];
};

__global__ void bbox_kernel(int n, const double *__restrict__ x, double *__restrict__ out /* lo[3], hi[3] as ordered bits */) {
  double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
  for (int a = blockIdx.x * blockDim.x + threadIdx.x; a < n; a += gridDim.x * blockDim.x)
    for (int d = 0; d < 3; ++d) {
      const double v = x[3 * (size_t)a + d];
      lo[d] = fmin(lo[d], v);
      hi[d] = fmax(hi[d], v);
    }
  for (int d = 0; d < 3; ++d) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo[d] = fmin(lo[d], __shfl_xor_sync(0xFFFFFFFFu, lo[d], o));
      hi[d] = fmax(hi[d], __shfl_xor_sync(0xFFFFFFFFu, hi[d], o));
    }
  }
  if ((threadIdx.x & 31) == 0) {
    // atomic min/max on doubles through a CAS loop (set-up time only)
    for (int d = 0; d < 3; ++d) {
      unsigned long long *pl = reinterpret_cast<unsigned long long *>(out + d), *ph = reinterpret_cast<unsigned long long *>(out + 3 + d);
      unsigned long long old = *pl;
      while (__longlong_as_double((long long)old) > lo[d]) {
        const unsigned long long prev = atomicCAS(pl, old, (unsigned long long)__double_as_longlong(lo[d]));
        if (prev == old) break;
        old = prev;
      }
      old = *ph;
      while (__longlong_as_double((long long)old) < hi[d]) {
        const unsigned long long prev = atomicCAS(ph, old, (unsigned long long)__double_as_longlong(hi[d]));
        if (prev == old) break;
        old = prev;
      }
    }
  }
}

__device__ __forceinline__ int cell_coord(double v, double lo, double inv, int nb) {
  return min(nb - 1, max(0, static_cast<int>((v - lo) * inv)));
}

__global__ void cell_id_kernel(int n, const double *__restrict__ x, CellGrid g, int *__restrict__ cell, int *__restrict__ atom) {
  int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= n) return;
  const int cx = cell_coord(x[3 * (size_t)a], g.lo[0], g.inv[0], g.nb[0]);
  const int cy = cell_coord(x[3 * (size_t)a + 1], g.lo[1], g.inv[1], g.nb[1]);
  const int cz = cell_coord(x[3 * (size_t)a + 2], g.lo[2], g.inv[2], g.nb[2]);
  cell[a] = (cz * g.nb[1] + cy) * g.nb[0] + cx;
  atom[a] = a;
}

// positions in cell order (x, y, z, atom id) and the start of every cell's run
__global__ void cell_ranges_kernel(int n, const int *__restrict__ sorted_cell, const int *__restrict__ sorted_atom,
                                   const double *__restrict__ x, double4 *__restrict__ xs, int *__restrict__ cell_start,
                                   int *__restrict__ cell_end) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  const int c = sorted_cell[s], a = sorted_atom[s];
  xs[s] = make_double4(x[3 * (size_t)a], x[3 * (size_t)a + 1], x[3 * (size_t)a + 2], bits_to_double((unsigned)a));
  if (s == 0 || sorted_cell[s - 1] != c) cell_start[c] = s;
  if (s == n - 1 || sorted_cell[s + 1] != c) cell_end[c] = s + 1;
}

// FILL = false: count neighbours of local atom i; FILL = true: write them at offsets[i]
template <bool FILL>
__global__ void __launch_bounds__(128) neighbor_pass_kernel(int nlocal, const double *__restrict__ x, CellGrid g, double cut_sq,
                                                            const double4 *__restrict__ xs, const int *__restrict__ cell_start,
                                                            const int *__restrict__ cell_end, long long *__restrict__ counts,
                                                            const long long *__restrict__ offsets, int *__restrict__ neigh) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nlocal) return;
  const double xi = x[3 * (size_t)i], yi = x[3 * (size_t)i + 1], zi = x[3 * (size_t)i + 2];
  const int cx = cell_coord(xi, g.lo[0], g.inv[0], g.nb[0]);
  const int cy = cell_coord(yi, g.lo[1], g.inv[1], g.nb[1]);
  const int cz = cell_coord(zi, g.lo[2], g.inv[2], g.nb[2]);
  long long n = 0;
  int *out = FILL ? neigh + offsets[i] : nullptr;
  for (int dz = -2; dz <= 2; ++dz) {
    const int z = cz + dz;
    if (z < 0 || z >= g.nb[2]) continue;
    for (int dy = -2; dy <= 2; ++dy) {
      const int y = cy + dy;
      if (y < 0 || y >= g.nb[1]) continue;
      // the five x cells of this row are consecutive cells: one contiguous run of the sorted array
      const int x0 = max(cx - 2, 0), x1 = min(cx + 2, g.nb[0] - 1);
      const int row = (z * g.nb[1] + y) * g.nb[0];
      int s0 = -1, s1 = -1;
      for (int c = row + x0; c <= row + x1; ++c) {
        const int b = cell_start[c], e = cell_end[c];
        if (e > b) {
          if (s0 < 0) s0 = b;
          s1 = e;
        }
      }
      for (int s = s0; s < s1; ++s) {
        const double4 p = xs[s];
        const int j = (int)double_to_bits(p.w);
        const double ddx = p.x - xi, ddy = p.y - yi, ddz = p.z - zi;
        if (j != i && ddx * ddx + ddy * ddy + ddz * ddz < cut_sq) {
          if (FILL) out[n] = j;
          ++n;
        }
      }
    }
  }
  if (!FILL) counts[i] = n;
}

}  // namespace ephb
