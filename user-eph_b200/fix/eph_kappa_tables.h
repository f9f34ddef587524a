// Host-side tables of `fix eph/atomic`: the `.kappa` file (reference eph_kappa.h:53-151) parsed into the same tables the
// reference builds -- rho_a(r) and rho_a(r^2) splines per element (make_cubic_table of eph_tables.h, bit-identical to
// EPH_Spline), the running-sum table E(T) per element and K(T) per "pair" slot (EPH_Linear, eph_linear.h) -- plus the
// two EPH_Linear look-ups the fix's constructor needs on the host.  Compile with -ffp-contract=off.
#pragma once

#include <algorithm>
#include <cmath>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "eph_tables.h"

namespace eph_b200 {

// EPH_Linear::operator() (eph_linear.h:40-47); in the last interval the reference reads one element past its slope
// vector -- the slope is taken as 0 there
inline double linear_eval(const LinearTable &t, double x) {
  const size_t idx = static_cast<size_t>(x / t.dx);
  if (idx < t.y.size()) {
    const double dy = idx + 1 < t.y.size() ? (t.y[idx + 1] - t.y[idx]) / t.dx : 0.0;
    return t.y[idx] + dy * (x - idx * t.dx);
  }
  return 0.;
}

// EPH_Linear::reverse_lookup (eph_linear.h:50-60)
inline double linear_reverse(const LinearTable &t, double yv) {
  auto it = std::upper_bound(t.y.begin(), t.y.end(), yv);
  if (it == t.y.end() || it == t.y.begin()) return 0.;   // below the first knot the reference indexes knot -1
  const size_t idx = static_cast<size_t>(it - t.y.begin()) - 1;
  const double dy = (t.y[idx + 1] - t.y[idx]) / t.dx;
  return idx * t.dx + 1. / dy * (yv - t.y[idx]);
}

struct KappaTables {
  int n_elements = 0, n_pairs = 0;
  size_t n_r = 0, n_T = 0;
  double r_cutoff = 0, r_cutoff_sq = 0, T_max = 0, dT = 0;
  std::vector<std::string> element_name;
  std::vector<int> element_number;
  std::vector<CubicTable> rho_r, rho_r_sq;   // [n_elements]
  std::vector<LinearTable> E_T;              // [n_elements]
  std::vector<LinearTable> K_T;              // [n_pairs]

  int find(const std::string &name) const {
    for (int e = 0; e < n_elements; ++e)
      if (element_name[e] == name) return e;
    return -1;
  }
  static std::vector<double> flatten(const std::vector<LinearTable> &t) {
    std::vector<double> out;
    for (const auto &l : t) out.insert(out.end(), l.y.begin(), l.y.end());
    return out;
  }
};

// `.kappa` grammar: 3 comment lines; "n_elements NAME..."; "n_r dr r_cutoff n_T dT T_max"; per element: Z, n_r values
// of rho_a(r), n_T values of C(T); then n_pairs blocks of n_T values of K(T).
inline KappaTables load_kappa_file(const std::string &path) {
  std::ifstream in(path);
  if (!in.is_open()) throw std::runtime_error("eph_b200: cannot open kappa file '" + path + "'");
  std::string line;
  for (int k = 0; k < 3; ++k) std::getline(in, line);
  KappaTables t;
  if (!(in >> t.n_elements) || t.n_elements < 1) throw std::runtime_error("fix_eph_atomic: no elements found in kappa file");
  t.n_pairs = t.n_elements > 1 ? (t.n_elements + 1) * (t.n_elements - 1) / 2 : 1;   // eph_kappa.h:69 (sic)
  std::getline(in, line);
  std::istringstream names(line);
  t.element_name.resize(t.n_elements);
  for (auto &nm : t.element_name) names >> nm;
  double dr;
  if (!(in >> t.n_r >> dr >> t.r_cutoff >> t.n_T >> t.dT >> t.T_max))
    throw std::runtime_error("eph_b200: bad kappa file header in '" + path + "'");
  t.r_cutoff_sq = t.r_cutoff * t.r_cutoff;
  const double dr_sq = t.r_cutoff_sq / (static_cast<double>(t.n_r - 1));
  for (int e = 0; e < t.n_elements; ++e) {
    int z;
    in >> z;
    t.element_number.push_back(z);
    std::vector<double> r(t.n_r), C(t.n_T);
    for (auto &v : r) in >> v;
    CubicTable rho = make_cubic_table(dr, r);
    for (size_t j = 0; j < t.n_r; ++j) r[j] = rho(std::sqrt(j * dr_sq));   // eph_kappa.h:125-127
    t.rho_r_sq.push_back(make_cubic_table(dr_sq, r));
    t.rho_r.push_back(std::move(rho));
    for (auto &v : C) in >> v;
    if (!in) throw std::runtime_error("eph_b200: kappa file '" + path + "' ends early");
    C[0] = 0.;                                                             // E(T) = running sum of C dT, :136-140
    for (size_t j = 1; j < t.n_T; ++j) C[j] = C[j - 1] + C[j] * t.dT;
    LinearTable E;
    E.dx = t.dT;
    E.y = std::move(C);
    t.E_T.push_back(std::move(E));
  }
  for (int p = 0; p < t.n_pairs; ++p) {                                    // :145-151
    LinearTable K;
    K.dx = t.dT;
    K.y.resize(t.n_T);
    for (auto &v : K.y) in >> v;
    if (!in) throw std::runtime_error("eph_b200: kappa file '" + path + "' ends early");
    t.K_T.push_back(std::move(K));
  }
  return t;
}

}  // namespace eph_b200
