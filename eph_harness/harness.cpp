// Harness helpers that play LAMMPS' role for tests and benchmarks: periodic
// ghost images and a cell-binned FULL neighbour list (cut-off r_c + skin) in
// CSR form, as `neighbor->add_request(this, REQ_FULL|REQ_GHOST)` would hand to
// the fix (reference fix_eph.cpp:273-275).  Host-side, OpenMP; not a hot path.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

extern "C" {

// Ghost images of a brick [lo,hi) cut out of a periodic box [0,L): every atom
// image (including shifted copies of the brick's own atoms) that lies in the
// shell of width `shell` around the brick but outside the brick itself.
// Pass 1 (out arrays NULL) returns the count; pass 2 fills them.
// xall: all atoms of the periodic box; owner_global receives the index into xall.
long long eph_harness_ghosts(long long nall, const double *xall, const double *L, const double *lo, const double *hi,
                             double shell, double *xg, long long *owner_global, int *shift_out) {
  long long count = 0;
  for (long long i = 0; i < nall; ++i) {
    const double *p = xall + 3 * i;
    for (int sx = -1; sx <= 1; ++sx) {
      double x = p[0] + sx * L[0];
      if (x < lo[0] - shell || x >= hi[0] + shell) continue;
      for (int sy = -1; sy <= 1; ++sy) {
        double y = p[1] + sy * L[1];
        if (y < lo[1] - shell || y >= hi[1] + shell) continue;
        for (int sz = -1; sz <= 1; ++sz) {
          double z = p[2] + sz * L[2];
          if (z < lo[2] - shell || z >= hi[2] + shell) continue;
          bool inside = x >= lo[0] && x < hi[0] && y >= lo[1] && y < hi[1] && z >= lo[2] && z < hi[2];
          if (inside) continue;  // that is a local atom of the brick
          if (xg) {
            xg[3 * count] = x; xg[3 * count + 1] = y; xg[3 * count + 2] = z;
            owner_global[count] = i;
            if (shift_out) { shift_out[3 * count] = sx; shift_out[3 * count + 1] = sy; shift_out[3 * count + 2] = sz; }
          }
          ++count;
        }
      }
    }
  }
  return count;
}

// Full neighbour list of atoms [0,nlocal) over all ntotal = nlocal+nghost atoms.
// Two passes: counts -> offsets (caller prefix-sums) -> fill.
// x: [ntotal][3].  If `flat` is NULL only numneigh[] is written.
void eph_harness_neighbors(long long nlocal, long long ntotal, const double *x, double cut, int *numneigh,
                           const long long *offsets, int *flat) {
  double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
  for (long long i = 0; i < ntotal; ++i)
    for (int d = 0; d < 3; ++d) {
      lo[d] = std::min(lo[d], x[3 * i + d]);
      hi[d] = std::max(hi[d], x[3 * i + d]);
    }
  int nb[3];
  double inv[3];
  for (int d = 0; d < 3; ++d) {
    nb[d] = std::max(1, (int)std::floor((hi[d] - lo[d]) / cut));
    inv[d] = nb[d] / ((hi[d] - lo[d]) * (1.0 + 1e-12) + 1e-9);
  }
  long long ncell = (long long)nb[0] * nb[1] * nb[2];
  std::vector<long long> cell_start(ncell + 1, 0);
  std::vector<int> cell_of(ntotal);
  for (long long i = 0; i < ntotal; ++i) {
    int c[3];
    for (int d = 0; d < 3; ++d) c[d] = std::min(nb[d] - 1, std::max(0, (int)((x[3 * i + d] - lo[d]) * inv[d])));
    int id = (c[2] * nb[1] + c[1]) * nb[0] + c[0];
    cell_of[i] = id;
    ++cell_start[id + 1];
  }
  for (long long c = 0; c < ncell; ++c) cell_start[c + 1] += cell_start[c];
  std::vector<long long> cursor(cell_start.begin(), cell_start.end() - 1);
  std::vector<int> sorted(ntotal);
  for (long long i = 0; i < ntotal; ++i) sorted[cursor[cell_of[i]]++] = (int)i;  // ascending index inside a cell

  const double cut2 = cut * cut;
#pragma omp parallel for schedule(dynamic, 1024)
  for (long long i = 0; i < nlocal; ++i) {
    int id = cell_of[i];
    int cx = id % nb[0], cy = (id / nb[0]) % nb[1], cz = id / (nb[0] * nb[1]);
    const double xi = x[3 * i], yi = x[3 * i + 1], zi = x[3 * i + 2];
    int n = 0;
    int *out = flat ? flat + offsets[i] : nullptr;
    for (int dz = -1; dz <= 1; ++dz) {
      int z = cz + dz;
      if (z < 0 || z >= nb[2]) continue;
      for (int dy = -1; dy <= 1; ++dy) {
        int y = cy + dy;
        if (y < 0 || y >= nb[1]) continue;
        for (int dx = -1; dx <= 1; ++dx) {
          int xx = cx + dx;
          if (xx < 0 || xx >= nb[0]) continue;
          long long c = ((long long)z * nb[1] + y) * nb[0] + xx;
          for (long long s = cell_start[c]; s < cell_start[c + 1]; ++s) {
            int j = sorted[s];
            if (j == i) continue;
            double ddx = x[3 * (long long)j] - xi, ddy = x[3 * (long long)j + 1] - yi, ddz = x[3 * (long long)j + 2] - zi;
            if (ddx * ddx + ddy * ddy + ddz * ddz < cut2) {
              if (out) out[n] = j;
              ++n;
            }
          }
        }
      }
    }
    if (out) std::sort(out, out + n);
    numneigh[i] = n;
  }
}

}  // extern "C"
