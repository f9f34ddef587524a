// Host-side description of the electronic-temperature grid and its files.
// Keeps the reference's on-disk grammars byte-compatible:
//   grid / restart file   reference eph_fdm.h:48-119 (reader), :226-265 (writer)
//   parameter file        reference eph_fdm.h:74-104
//   heat map T_out_%06d   reference eph_fdm.h:198-224
// The arrays only live here between the file and eph_b200_set_grid(); the solve
// itself runs on the device.
#pragma once

#include <cstdint>
#include <cstdio>
#include <fstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "eph_tables.h"

namespace eph_b200 {

struct GridState {
  size_t nx = 1, ny = 1, nz = 1, steps = 1;
  double box[6] = {0, 1, 0, 1, 0, 1};  // x0 x1 y0 y1 z0 z1
  std::string parameter_filename = "NULL";
  std::vector<double> T_e, S_e, rho_e, C_e, kappa_e;
  std::vector<int16_t> flag;     // 0 constant, 1 dynamic, 2 zero-derivative wall
  std::vector<uint16_t> t_dyn;   // 1: C_e, kappa_e follow T_e
  bool has_tables = false;
  CubicTable C_e_T, kappa_e_T;
  LinearTable E_e_T;

  size_t ncell() const { return nx * ny * nz; }
  double dx() const { return (box[1] - box[0]) / nx; }
  double dy() const { return (box[3] - box[2]) / ny; }
  double dz() const { return (box[5] - box[4]) / nz; }
  void resize() {
    const size_t n = ncell();
    T_e.assign(n, 0.0); S_e.assign(n, 0.0); rho_e.assign(n, 0.0); C_e.assign(n, 0.0); kappa_e.assign(n, 0.0);
    flag.assign(n, 1); t_dyn.assign(n, 0);
  }
};

// `fix eph ... NX NY NZ NULL ...`: homogeneous grid over the simulation box (eph_fdm.h:28-46)
inline GridState make_uniform_grid(size_t nx, size_t ny, size_t nz, const double *boxlo, const double *boxhi, double T_e,
                                   double C_e, double rho_e, double kappa_e) {
  GridState g;
  g.nx = nx; g.ny = ny; g.nz = nz; g.steps = 1;
  for (int d = 0; d < 3; ++d) { g.box[2 * d] = boxlo[d]; g.box[2 * d + 1] = boxhi[d]; }
  g.resize();
  for (size_t i = 0; i < g.ncell(); ++i) { g.T_e[i] = T_e; g.C_e[i] = C_e; g.rho_e[i] = rho_e; g.kappa_e[i] = kappa_e; }
  return g;
}

inline void skip_comment_lines(std::ifstream &in, const std::string &path) {
  std::string line;
  for (int k = 0; k < 3; ++k) {
    std::getline(in, line);
    if (line.empty() || line[0] != '#') throw std::runtime_error("eph_b200: '" + path + "': expected three '#' comment lines");
  }
}

inline void load_parameter_file(GridState &g, const std::string &path) {
  std::ifstream in(path);
  if (!in.is_open()) throw std::runtime_error("eph_b200: cannot open parameter file '" + path + "'");
  skip_comment_lines(in, path);
  size_t n;
  double dT;
  if (!(in >> n >> dT)) throw std::runtime_error("eph_b200: bad parameter file header '" + path + "'");
  std::vector<double> C(n), K(n);
  for (size_t i = 0; i < n; ++i) in >> C[i] >> K[i];
  if (!in) throw std::runtime_error("eph_b200: parameter file '" + path + "' ends early");
  g.C_e_T = make_cubic_table(dT, C);
  g.kappa_e_T = make_cubic_table(dT, K);
  C[0] = 0.;  // E_e(T): running sum of C_e dT (eph_fdm.h:98-102)
  for (size_t i = 1; i < n; ++i) C[i] = C[i - 1] + C[i] * dT;
  g.E_e_T.dx = dT;
  g.E_e_T.y = std::move(C);
  g.has_tables = true;
}

inline GridState load_grid_file(const std::string &path) {
  std::ifstream in(path);
  if (!in.is_open()) throw std::runtime_error("eph_b200: cannot open grid file '" + path + "'");
  skip_comment_lines(in, path);
  GridState g;
  if (!(in >> g.nx >> g.ny >> g.nz >> g.steps) || g.nx < 1 || g.ny < 1 || g.nz < 1)
    throw std::runtime_error("eph_b200: bad grid size line in '" + path + "'");
  for (double &b : g.box) in >> b;
  in >> g.parameter_filename;
  if (!in) throw std::runtime_error("eph_b200: bad header in '" + path + "'");
  g.resize();
  if (g.parameter_filename != "NULL") load_parameter_file(g, g.parameter_filename);
  for (size_t r = 0; r != g.ncell(); ++r) {
    int i, j, k, fl, td;
    in >> i >> j >> k;
    const size_t idx = i + j * g.nx + k * g.nx * g.ny;
    if (!in || idx >= g.ncell()) throw std::runtime_error("eph_b200: bad grid record in '" + path + "'");
    in >> g.T_e[idx] >> g.S_e[idx] >> g.rho_e[idx] >> g.C_e[idx] >> g.kappa_e[idx] >> fl >> td;
    g.flag[idx] = static_cast<int16_t>(fl);
    g.t_dyn[idx] = static_cast<uint16_t>(td);
  }
  if (!in) throw std::runtime_error("eph_b200: grid file '" + path + "' ends early");
  return g;
}

// heat map `<name>_%06d`: "x y z Te" header, lower cell corner, i fastest
inline void write_heat_map(const GridState &g, const std::vector<double> &T_e, const std::string &name, int counter) {
  char fn[1200];
  std::snprintf(fn, sizeof fn, "%s_%06d", name.c_str(), counter);
  FILE *fd = std::fopen(fn, "w");
  if (!fd) throw std::runtime_error(std::string("eph_b200: cannot write '") + fn + "'");
  std::fprintf(fd, "x y z Te\n");
  const double dx = g.dx(), dy = g.dy(), dz = g.dz();
  for (int k = 0; k < (int)g.nz; ++k)
    for (int j = 0; j < (int)g.ny; ++j)
      for (int i = 0; i < (int)g.nx; ++i) {
        const size_t idx = i + j * g.nx + k * g.nx * g.ny;
        std::fprintf(fd, "%.6e %.6e %.6e %.6e\n", g.box[0] + i * dx, g.box[2] + j * dy, g.box[4] + k * dz, T_e[idx]);
      }
  std::fclose(fd);
}

// restart file: the grid-file grammar written back with %.6e
inline void write_restart(const GridState &g, const std::vector<double> &T_e, const std::string &path) {
  FILE *fd = std::fopen(path.c_str(), "w");
  if (!fd) throw std::runtime_error("eph_b200: cannot write '" + path + "'");
  std::fprintf(fd, "# A comment\n#\n#\n");
  std::fprintf(fd, "%ld %ld %ld %ld\n", (long)g.nx, (long)g.ny, (long)g.nz, (long)g.steps);
  std::fprintf(fd, "%.6e %.6e\n%.6e %.6e\n%.6e %.6e\n", g.box[0], g.box[1], g.box[2], g.box[3], g.box[4], g.box[5]);
  std::fprintf(fd, "%s\n", g.parameter_filename.c_str());
  for (int k = 0; k < (int)g.nz; ++k)
    for (int j = 0; j < (int)g.ny; ++j)
      for (int i = 0; i < (int)g.nx; ++i) {
        const size_t idx = i + j * g.nx + k * g.nx * g.ny;
        std::fprintf(fd, "%d %d %d %.6e %.6e %.6e %.6e %.6e %d %d\n", i, j, k, T_e[idx], g.S_e[idx], g.rho_e[idx],
                     g.C_e[idx], g.kappa_e[idx], (int)g.flag[idx], (int)g.t_dyn[idx]);
      }
  std::fclose(fd);
}

}  // namespace eph_b200
