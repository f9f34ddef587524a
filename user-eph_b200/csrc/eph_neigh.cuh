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

// The same two passes, cooperatively: one CTA per row of cells (fixed y, z), its warps take chunks of 32 consecutive
// atoms of the row in cell order.  The atoms of a chunk share one candidate set -- the cells from two left of the
// chunk's first atom to two right of its last one, in the 25 rows around -- which the warp stages through shared memory
// 32 positions at a time (one coalesced 1 KB load) and every lane then reads as broadcasts.  The per-atom kernel above
// gathers ~1000 sectors per atom and pass (25 x 10 cell-range look-ups + ~490 candidates); this one ~75 wavefronts.
// Neighbours come out in the same order (rows by dz, dy; ascending cell order within a row), so both kernels produce
// the same list.  Rows of ghost atoms are not built (the list has rows for local atoms only).
constexpr int kNeighWarps = 4;
template <bool FILL>
__global__ void __launch_bounds__(32 * kNeighWarps) neighbor_tile_kernel(int nlocal, CellGrid g, double cut_sq, const double4 *__restrict__ xs,
                                                                         const int *__restrict__ cell_start, const int *__restrict__ cell_end,
                                                                         long long *__restrict__ counts, const long long *__restrict__ offsets,
                                                                         int *__restrict__ neigh) {
  __shared__ double4 s_cand[kNeighWarps][32];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int row = blockIdx.x;   // y + nb[1] * z
  const int ry = row % g.nb[1], rz = row / g.nb[1];
  // the row's atoms: one contiguous run of the cell-sorted array (empty cells have start = end = 0)
  int r0 = -1, r1 = -1;
  for (int c = row * g.nb[0] + lane; c < (row + 1) * g.nb[0]; c += 32) {
    const int b = cell_start[c], e = cell_end[c];
    if (e > b) { r0 = (r0 < 0 || b < r0) ? b : r0; r1 = e > r1 ? e : r1; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const int q0 = __shfl_xor_sync(0xFFFFFFFFu, r0, o), q1 = __shfl_xor_sync(0xFFFFFFFFu, r1, o);
    r0 = (r0 < 0) ? q0 : ((q0 >= 0 && q0 < r0) ? q0 : r0);
    r1 = q1 > r1 ? q1 : r1;
  }
  if (r0 < 0) return;
  for (int base = r0 + 32 * wib; base < r1; base += 32 * kNeighWarps) {
    const int sme = base + lane;
    const bool have = sme < r1;
    double4 me = make_double4(0, 0, 0, 0);
    int id = -1, cx = 0;
    if (have) {
      me = xs[sme];
      id = (int)double_to_bits(me.w);
      cx = cell_coord(me.x, g.lo[0], g.inv[0], g.nb[0]);
    }
    const bool mine = have && id < nlocal;   // ghosts have no row
    // x-cell range of the chunk (atoms are in cell order: first and last valid lane)
    const unsigned vmask = __ballot_sync(0xFFFFFFFFu, have);
    const int cxa = __shfl_sync(0xFFFFFFFFu, cx, 0), cxb = __shfl_sync(0xFFFFFFFFu, cx, 31 - __clz(vmask));
    const int x0 = max(cxa - 2, 0), x1 = min(cxb + 2, g.nb[0] - 1);
    long long n = 0;
    int *out = (FILL && mine) ? neigh + offsets[id] : nullptr;
    for (int dz = -2; dz <= 2; ++dz) {
      const int z = rz + dz;
      if (z < 0 || z >= g.nb[2]) continue;
      for (int dy = -2; dy <= 2; ++dy) {
        const int y = ry + dy;
        if (y < 0 || y >= g.nb[1]) continue;
        const int crow = (z * g.nb[1] + y) * g.nb[0];
        // the run of the cells [x0, x1] of that row
        int s0 = -1, s1 = -1;
        for (int c = crow + x0 + lane; c <= crow + x1; c += 32) {
          const int b = cell_start[c], e = cell_end[c];
          if (e > b) { s0 = (s0 < 0 || b < s0) ? b : s0; s1 = e > s1 ? e : s1; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const int q0 = __shfl_xor_sync(0xFFFFFFFFu, s0, o), q1 = __shfl_xor_sync(0xFFFFFFFFu, s1, o);
          s0 = (s0 < 0) ? q0 : ((q0 >= 0 && q0 < s0) ? q0 : s0);
          s1 = q1 > s1 ? q1 : s1;
        }
        if (s0 < 0) continue;
        // per-atom window inside the run: cells [cx - 2, cx + 2] only (what the per-atom kernel walks); outside it the
        // distance test fails anyway (cells are at least cut-off / 2 wide), so testing the whole run gives the same rows
        for (int cb = s0; cb < s1; cb += 32) {
          __syncwarp();
          if (cb + lane < s1) s_cand[wib][lane] = xs[cb + lane];
          __syncwarp();
          const int m = min(32, s1 - cb);
          if (mine) {
            for (int c = 0; c < m; ++c) {
              const double4 p = s_cand[wib][c];
              const double ddx = p.x - me.x, ddy = p.y - me.y, ddz = p.z - me.z;
              const int j = (int)double_to_bits(p.w);
              if (j != id && ddx * ddx + ddy * ddy + ddz * ddz < cut_sq) {
                if (FILL) out[n] = j;
                ++n;
              }
            }
          }
        }
      }
    }
    if (!FILL && mine) counts[id] = n;
  }
}

}  // namespace ephb
