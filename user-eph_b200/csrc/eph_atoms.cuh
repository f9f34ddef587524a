// Streaming per-atom kernels around the sweeps: packing LAMMPS arrays into the
// gather-friendly records, the per-atom coupling s = alpha(rho)/rho, ghost
// fills, force accumulation, energy deposition and the integrator hooks.
#pragma once

#include "eph_device.cuh"
#include "eph_packed.cuh"

namespace ephb {

// x, v ([n][3], LAMMPS layout) -> pos4 {x,y,z,bits} and the density-pass record pv {x,y,z,bits | vx,vy,vz,0}
// (eph_sweeps.cuh); with track != 0 also the displacement check that guards the inner list (against xref, the
// positions when the inner list was built) and, with track0 != 0 -- only in a step that rebuilds the inner list from
// a LAMMPS list that is not fresh --, the largest displacement since LAMMPS built its list (against xref0).
__global__ void __launch_bounds__(256) pack_atoms_kernel(int ntotal, const double *__restrict__ x, const double *__restrict__ v,
                                  const int *__restrict__ type, const int *__restrict__ mask,
                                  const int *__restrict__ type_map, int groupbit, double4 *__restrict__ pos4,
                                  double4 *__restrict__ pv, int track, int track0, const double4 *__restrict__ xref,
                                  const double4 *__restrict__ xref0, ListState *__restrict__ st,
                                  Packed32 *__restrict__ recD, double inv_period, unsigned *__restrict__ status) {
  __shared__ double s_max[8];
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  double d0 = 0.0;
  if (a < ntotal) {
    unsigned bits = (static_cast<unsigned>(type_map[type[a] - 1]) & kElemMask) |
                    ((static_cast<unsigned>(type[a] - 1) & 0xFFu) << kTypeShift);
    if (mask[a] & groupbit) bits |= kBitGroup;
    const double px = x[3 * (size_t)a], py = x[3 * (size_t)a + 1], pz = x[3 * (size_t)a + 2];
    const double4 p4 = make_double4(px, py, pz, bits_to_double(bits));
    pos4[a] = p4;
    const double vx = v[3 * (size_t)a], vy = v[3 * (size_t)a + 1], vz = v[3 * (size_t)a + 2];
    if (pv != nullptr) {   // fp64 records: steps that walk LAMMPS' list (always, without packed records)
      pv[2 * (size_t)a] = p4;
      pv[2 * (size_t)a + 1] = make_double4(vx, vy, vz, 0.0);
    }
    if (recD != nullptr) {   // packed density-pass record (eph_packed.cuh); flags: element index
      Packed32 r;
      r.p = pack_position(px, py, pz, inv_period, bits & 3u);
      r.b = pack_vector(vx, vy, vz, status);
      recD[a] = r;
    }
    if (track) {
      const double4 r = xref[a];
      const double dx = px - r.x, dy = py - r.y, dz = pz - r.z;
      if (dx * dx + dy * dy + dz * dz > st->guard_sq) st->inner_invalid = 1u;
    }
    if (track0) {
      const double4 r0 = xref0[a];
      const double ex = px - r0.x, ey = py - r0.y, ez = pz - r0.z;
      d0 = ex * ex + ey * ey + ez * ez;
    }
  }
  if (track0) {  // block-wide max of the displacement since LAMMPS' build, one atomic per block
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) d0 = fmax(d0, __shfl_xor_sync(0xFFFFFFFFu, d0, o));
    if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = d0;
    __syncthreads();
    if (threadIdx.x == 0) {
      double m = s_max[0];
      for (int k = 1; k < (int)(blockDim.x >> 5); ++k) m = fmax(m, s_max[k]);
      atomicMax(&st->disp0_sq_bits, (unsigned long long)__double_as_longlong(m));
    }
  }
}

// fp64 density-pass records for a step that turns out to need them: the packed density pass returns at once when the
// displacement guard of pack_atoms has invalidated the inner list, and the fp64 pass that walks LAMMPS' list instead
// gathers from pv.  Launched behind pack_atoms in every packed step; does nothing while the inner list is valid.
__global__ void __launch_bounds__(256) pv_fill_kernel(int ntotal, const double *__restrict__ v, const double4 *__restrict__ pos4,
                                                      double4 *__restrict__ pv, const ListState *__restrict__ st) {
  if (st->inner_invalid == 0u) return;
  for (int a = blockIdx.x * blockDim.x + threadIdx.x; a < ntotal; a += gridDim.x * blockDim.x) {
    pv[2 * (size_t)a] = pos4[a];
    pv[2 * (size_t)a + 1] = make_double4(v[3 * (size_t)a], v[3 * (size_t)a + 1], v[3 * (size_t)a + 2], 0.0);
  }
}

constexpr double kMinInnerSkin = 0.1;   // A

struct PrepArgs {
  int nlocal, ntotal;
  const int *__restrict__ owner;      // [nghost] local owner of each ghost (single-rank images) or nullptr
  const long long *__restrict__ tag;  // [ntotal]
  const double *__restrict__ xi_inject;  // [nlocal][3] or nullptr
  const double2 *__restrict__ alpha_tab;  // [n_elements][n_beta][2]
  int n_beta;
  double inv_drho, rho_cutoff;
  unsigned long long seed, step;
  int do_random;
  double *__restrict__ rho;    // [ntotal]; ghost entries are filled here
  const double4 *__restrict__ W4;  // [nlocal] (or [ntotal] after an exchange) pair sums of the density pass
  const double4 *__restrict__ pos4;  // [ntotal] x, y, z, bits of pack_atoms
  double4 *__restrict__ puz;   // [ntotal][3] force-pass record {x,y,z,bits+valid | u = s*w, z = s*xi | var, cell} (eph_sweeps.cuh)
  const double *__restrict__ T_e;   // grid temperatures (nullptr: no grid) ...
  GridGeom grid;                    // ... and geometry: the grid cell of every local atom (EPH_FDM::get_index,
  double eta_factor;                // eph_fdm.h:494-509: six fp64 divisions) is worked out ONCE per step here, with
                                    // all lanes busy, and stored with eta_factor * sqrt(T_e(cell)) (fix_eph.cpp:829-833)
                                    // in the atom's record for the force pass and the deposition
  double *__restrict__ w;      // [nlocal][3] w_i (probe / forward-comm payload)
  double *__restrict__ xi;     // [ntotal][3] xi_i (probe / XI forward-comm slots)
  unsigned *__restrict__ status;
  // the step that (re)built the inner list validates it here (stream order: after the rho sweep)
  int built_inner;
  double skin, inner_skin;
  ListState *__restrict__ list_state;
  // packed force-pass records (eph_packed.cuh), nullptr without them: A = {position + valid/group flags | u}, B = {z},
  // var = eta_factor sqrt(T_e(cell)) per local atom
  Packed32 *__restrict__ recA;
  Block16 *__restrict__ recB;
  double *__restrict__ var;
  double inv_period;
  int puz_mode;   // fp64 record puz: 2 always, 1 only if the inner list is invalid (the fp64 force pass will run), 0 never
};

// Between the two passes, for locals and ghosts alike: ghost rho and W come from the owner (the reference's
// forward comms RHO and WI, fix_eph.cpp:870-871, :743-744, as one step), s = alpha(rho)/rho with alpha = 0
// above rho_cutoff (eph_beta.h:186-198), the rho>0 validity bit, w = s W, u = s w, xi (fix_eph.cpp:854-865)
// and z = s xi.
__global__ void prep_coupling_kernel(PrepArgs p) {
  int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= p.ntotal) return;
  if (a == 0 && p.built_inner) {
    // An inner list built now from LAMMPS' list holds every pair that is closer than r_c + s, s = min(inner_skin,
    // skin - 2 D), D = largest displacement since LAMMPS built its list (LAMMPS' list is complete up to r_c + skin at ITS
    // build time).  It therefore stands until some atom has moved s / 2 since this build: late in the life of LAMMPS'
    // list the guard simply trips sooner, instead of the engine having to walk the full list.  Below kMinInnerSkin the
    // rebuilds would come every few steps: the list is declared invalid and the engine stays on LAMMPS' list.
    const double d0 = sqrt(__longlong_as_double((long long)p.list_state->disp0_sq_bits));
    const double s = fmin(p.inner_skin, p.skin - 2.0 * d0);
    const bool ok = s >= fmin(kMinInnerSkin, p.inner_skin);
    p.list_state->inner_invalid = ok ? 0u : 1u;
    p.list_state->guard_sq = ok ? 0.25 * s * s : 0.0;
  }
  // ghosts: owner >= 0 is a local atom of this rank (periodic image); owner < 0 means the owner lives on another
  // rank and the exchange already wrote {rho, W} into the ghost's own slot
  int src = a;
  if (a >= p.nlocal && p.owner != nullptr) {
    const int o = p.owner[a - p.nlocal];
    if (o >= 0) src = o;
  }
  double rho = p.rho[src];
  if (a >= p.nlocal) p.rho[a] = rho;
  const double4 W = p.W4[src];
  double4 pa = p.pos4[a];
  unsigned bits = double_to_bits(pa.w) & ~kBitValid;
  double s = 0.0;
  if (rho > 0) {
    bits |= kBitValid;
    double alpha = 0.0;
    if (rho > p.rho_cutoff) atomicOr(p.status, 1u);
    else alpha = spline_eval(p.alpha_tab + 2 * (size_t)(bits & kElemMask) * p.n_beta, p.inv_drho, rho);
    s = alpha / rho;
  }
  pa.w = bits_to_double(bits);
  const double wx = s * W.x, wy = s * W.y, wz = s * W.z;   // w_i = alpha_i/rho_i * sum (prescaler of fix_eph.cpp:727)
  if (a < p.nlocal) {
    p.w[3 * (size_t)a] = wx; p.w[3 * (size_t)a + 1] = wy; p.w[3 * (size_t)a + 2] = wz;
  }
  double xi[3] = {0.0, 0.0, 0.0};
  if (p.do_random && (bits & kBitGroup)) {
    if (p.xi_inject) {
      // injected stream: locals and own images read the caller's array; ghosts owned elsewhere read the slot the XI
      // forward comm filled (fix_eph.cpp:863-864)
      const double *q = src < p.nlocal ? p.xi_inject + 3 * (size_t)src : p.xi + 3 * (size_t)a;
      xi[0] = q[0]; xi[1] = q[1]; xi[2] = q[2];
    } else {
      xi_stream(p.seed, p.step, p.tag[a], xi);
    }
  }
  double var = 0.0;
  int cell = 0;
  if (a < p.nlocal && p.T_e != nullptr) {
    cell = grid_index(p.grid, pa.x, pa.y, pa.z);
    if (p.do_random) var = p.eta_factor * sqrt(p.T_e[cell]);
  }
  const double ux = s * wx, uy = s * wy, uz = s * wz, zx = s * xi[0], zy = s * xi[1], zz = s * xi[2];
  if (p.puz_mode == 2 || (p.puz_mode == 1 && p.list_state->inner_invalid != 0u)) {
    double4 *rec = p.puz + 3 * (size_t)a;
    rec[0] = pa;
    rec[1] = make_double4(ux, uy, uz, zx);
    rec[2] = make_double4(zy, zz, var, bits_to_double(static_cast<unsigned>(cell)));
  }
  if (p.recA != nullptr) {
    Packed32 r;
    r.p = pack_position(pa.x, pa.y, pa.z, p.inv_period, ((bits & kBitValid) ? 1u : 0u) | ((bits & kBitGroup) ? 2u : 0u));
    r.b = pack_vector(ux, uy, uz, p.status);
    p.recA[a] = r;
    if (p.do_random) p.recB[a] = pack_vector(zx, zy, zz, p.status);
    if (a < p.nlocal) p.var[a] = var;
  }
  if (a < p.nlocal) {
    p.xi[3 * (size_t)a] = xi[0]; p.xi[3 * (size_t)a + 1] = xi[1]; p.xi[3 * (size_t)a + 2] = xi[2];
  }
}

// The one ghost exchange of a step: {rho, Wx, Wy, Wz} of the listed owned atoms into a send buffer, and from a
// receive buffer into the listed (ghost) slots.  Device buffers: the transport is NCCL (or peer memory), not the host.
__global__ void pack_payload_kernel(int n, const int *__restrict__ index, const double *__restrict__ rho,
                                    const double4 *__restrict__ W4, double4 *__restrict__ buf) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const int a = index[t];
  const double4 W = W4[a];
  buf[t] = make_double4(rho[a], W.x, W.y, W.z);
}
__global__ void unpack_payload_kernel(int n, const int *__restrict__ index, double *__restrict__ rho,
                                      double4 *__restrict__ W4, const double4 *__restrict__ buf, int ntotal) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const int a = index[t];
  if (a < 0 || a >= ntotal) return;
  const double4 b = buf[t];
  rho[a] = b.x;
  W4[a] = make_double4(b.y, b.z, b.w, 0.0);
}

struct DepositArgs {
  int nlocal;
  const double *__restrict__ x;      // LAMMPS layout, or nullptr: positions unchanged since post_force (use pos4)
  const double *__restrict__ v;
  const double4 *__restrict__ pos4;  // bits (group) from the last post_force
  const double *__restrict__ f_eph;
  const double *__restrict__ f_rng;
  double dt, dVdt;
  int do_friction, do_random;
  GridGeom grid;
  double *__restrict__ dT_e;
  double *__restrict__ E_sum;    // one double, accumulated atomically per block
};

// FixEPH::end_of_step (fix_eph.cpp:350-429) up to the grid solve: per-atom
// energy transfer, EPH_FDM::insert_energy (eph_fdm.h:172-179) with one atomic
// per distinct cell per warp, and the energy sum.  The 8-column per-atom output is
// produced on demand (peratom_kernel): nothing downstream of the step reads it on the device.
__global__ void __launch_bounds__(256) deposit_kernel(DepositArgs d) {
  __shared__ double s_part[8];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  double dE = 0.0;       // energy handed to the electrons by this atom
  double contrib = 0.0;  // the same as a power density for its grid cell
  int cell = -1;
  if (i < d.nlocal) {
    const unsigned bits = double_to_bits(d.pos4[i].w);
    const size_t o = 3 * (size_t)i;
    if (bits & kBitGroup) {
      const double vx = d.v[o], vy = d.v[o + 1], vz = d.v[o + 2];
      const double ex = d.f_eph[o], ey = d.f_eph[o + 1], ez = d.f_eph[o + 2];
      const double rx = d.f_rng[o], ry = d.f_rng[o + 1], rz = d.f_rng[o + 2];
      double dEf = 0.0, dEr = 0.0;
      if (d.do_friction) { dEf -= ex * vx * d.dt; dEf -= ey * vy * d.dt; dEf -= ez * vz * d.dt; }
      if (d.do_random) { dEr -= rx * vx * d.dt; dEr -= ry * vy * d.dt; dEr -= rz * vz * d.dt; }
      dE = dEf + dEr;
      contrib = dEf / d.dVdt + dEr / d.dVdt;  // two insert_energy calls in the reference
      if (d.x != nullptr) cell = grid_index(d.grid, d.x[o], d.x[o + 1], d.x[o + 2]);
      else {   // positions unchanged since post_force (recomputing from the packed position is cheaper than reading the
        const double4 p4 = d.pos4[i];   // cell cached in the 96-byte-stride force record: measured)
        cell = grid_index(d.grid, p4.x, p4.y, p4.z);
      }
    }
  }
  // warp-aggregated scatter-add: atoms are spatially sorted, so a warp usually
  // touches one to three cells; each distinct cell costs one fp64 atomic.
  unsigned remaining = __ballot_sync(0xFFFFFFFFu, cell >= 0);
  while (remaining) {
    const int leader = __ffs(remaining) - 1;
    const int lcell = __shfl_sync(0xFFFFFFFFu, cell, leader);
    const bool mine = (cell == lcell);
    const double sum = warp_sum(mine ? contrib : 0.0);
    if (lane == leader) atomicAdd(&d.dT_e[lcell], sum);
    remaining &= ~__ballot_sync(0xFFFFFFFFu, mine);
  }
  // energy transferred this step (E_local of fix_eph.cpp:357-404)
  dE = warp_sum(dE);
  if (lane == 0) s_part[threadIdx.x >> 5] = dE;
  __syncthreads();
  if (threadIdx.x < 32) {
    double t = (threadIdx.x < (blockDim.x >> 5)) ? s_part[threadIdx.x] : 0.0;
    t = warp_sum(t);
    if (threadIdx.x == 0 && t != 0.0) atomicAdd(d.E_sum, t);
  }
}

// Per-atom output of the fix (fix_eph.cpp:406-428): rho_i, beta(rho_i), f_EPH, f_RNG; zeros outside the group.
// Materialised when the caller asks for it (eph_b200_get_peratom) from the arrays the step left on the device.
__global__ void __launch_bounds__(256) peratom_kernel(int nlocal, const double4 *__restrict__ pos4, const double *__restrict__ rho_i,
                                                      const double *__restrict__ f_eph, const double *__restrict__ f_rng,
                                                      const double2 *__restrict__ beta_tab, int n_beta, double inv_drho,
                                                      double rho_cutoff, double *__restrict__ array8) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nlocal) return;
  const unsigned bits = double_to_bits(pos4[i].w);
  double4 lo = make_double4(0, 0, 0, 0), hi = lo;
  if (bits & kBitGroup) {
    const size_t o = 3 * (size_t)i;
    const double rho = rho_i[i];
    double beta = 0.0;  // eph_beta.h:171-184
    if (!(rho > rho_cutoff)) beta = spline_eval(beta_tab + 2 * (size_t)(bits & kElemMask) * n_beta, inv_drho, rho);
    lo = make_double4(rho, beta, f_eph[o], f_eph[o + 1]);
    hi = make_double4(f_eph[o + 2], f_rng[o], f_rng[o + 1], f_rng[o + 2]);
  }
  double4 *row = reinterpret_cast<double4 *>(array8 + 8 * (size_t)i);
  row[0] = lo;
  row[1] = hi;
}

// FixEPH::initial_integrate / final_integrate (fix_eph.cpp:305-348)
__global__ void integrate_kernel(int nlocal, double *__restrict__ x, double *__restrict__ v,
                                 const double *__restrict__ f, const int *__restrict__ type,
                                 const int *__restrict__ mask, const double *__restrict__ mass, int groupbit,
                                 double dtv, double dtf, int drift) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nlocal) return;
  if (!(mask[i] & groupbit)) return;
  const double dtfm = dtf / mass[type[i]];
  const size_t o = 3 * (size_t)i;
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    double vv = v[o + d] + dtfm * f[o + d];
    v[o + d] = vv;
    if (drift) x[o + d] += dtv * vv;
  }
}

// FixEPH::pack_forward_comm / unpack_forward_comm (fix_eph.cpp:951-1009) on
// device-resident arrays; width 1 (rho) or 3 (xi, w).
__global__ void pack_forward_kernel(int n, const int *__restrict__ list, const double *__restrict__ src, int width,
                                    int stride, double *__restrict__ buf) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * width) return;
  int k = t / width, c = t - k * width;
  buf[t] = src[(size_t)list[k] * stride + c];
}
__global__ void unpack_forward_kernel(int n, int first, double *__restrict__ dst, int width, int stride,
                                      const double *__restrict__ buf) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * width) return;
  int k = t / width, c = t - k * width;
  dst[(size_t)(first + k) * stride + c] = buf[t];
}

}  // namespace ephb
