// Host-side table construction for the B200 `fix eph` path.
//
// The device kernels evaluate the reference's spline tables verbatim, so the
// tables themselves must be the ones the reference would build from the same
// `.beta` file: uniform-knot modified-Akima tangents and one cubic per interval
// stored as {a,b,c,d} in ABSOLUTE x (reference eph_spline.h:31-131), the
// rho(r) -> rho(r^2) resampling and alpha = sqrt(beta) tables of EPH_Beta
// (reference eph_beta.h:96-125) and the `.beta` grammar (eph_beta.h:39-128,
// Doc/Beta/input.beta).  tests/test_host_side.py (test_product_tables_*) checks bit-equality against the
// compiled reference.  Compile with -ffp-contract=off: the coefficients carry
// cancellation and must round exactly like the reference build.
#pragma once

#include <cmath>
#include <cstddef>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace eph_b200 {

struct CubicTable {
  double dx = 0.0, inv_dx = 0.0;
  std::vector<double> k;  // [n][4] = a, b, c, d ; value = a + x (b + x (c + x d))
  size_t size() const { return k.size() / 4; }
  double operator()(double x) const {  // eph_spline.h:134-142
    const double *c = &k[4 * static_cast<size_t>(x * inv_dx)];
    return c[0] + x * (c[1] + x * (c[2] + x * c[3]));
  }
};

// Modified-Akima cubic through uniformly spaced samples, as the reference
// constructs it.  Notation: m_i is the slope of segment i; the slope sequence is
// extended by two virtual segments at either end.  Two quirks of the reference
// are part of the table definition and are kept: the right extension starts
// from m_{n-1} = 0 (its slot is read before it is written, eph_spline.h:56-59),
// and the tangent rule for the last knot sees no slope to its right.
inline CubicTable make_cubic_table(double dx, const std::vector<double> &y) {
  const size_t n = y.size();
  if (!(dx > 0.0) || n < 5) throw std::runtime_error("eph_b200: spline needs dx > 0 and at least 5 knots");
  CubicTable t;
  t.dx = dx;
  t.inv_dx = 1. / dx;
  t.k.assign(4 * n, 0.0);

  // segment slopes m[0..n-2], plus the virtual ones
  std::vector<double> m(n, 0.0);
  for (size_t i = 0; i + 1 < n; ++i) m[i] = (y[i + 1] - y[i]) / dx;
  const double left1 = 2.0 * m[0] - m[1];       // m_{-1}
  const double left2 = 2.0 * left1 - m[0];      // m_{-2}
  const double right1 = 2.0 * m[n - 2] - 0.0;   // m_{n-1}
  const double right2 = 2.0 * right1 - 0.0;     // m_{n}
  m[n - 1] = right1;
  auto slope = [&](long i) -> double {
    if (i == -2) return left2;
    if (i == -1) return left1;
    if (i == static_cast<long>(n)) return right2;
    return m[static_cast<size_t>(i)];
  };

  // knot tangents
  std::vector<double> tan(n);
  for (size_t i = 0; i < n; ++i) {
    const long li = static_cast<long>(i);
    const double mm2 = slope(li - 2), mm1 = slope(li - 1), m0 = slope(li);
    const double mp1 = (i + 1 < n) ? slope(li + 1) : 0.0;
    const double w_next = std::fabs(slope(li + 1) - m0);   // |m_{i+1} - m_i|
    const double w_prev = std::fabs(mm1 - mm2);            // |m_{i-1} - m_{i-2}|
    double d;
    if (mm2 == mm1 && m0 != mp1) d = mm1;
    else if (m0 == mp1 && mm2 == mm1) d = m0;
    else if (mm1 == m0) d = m0;
    else if (mm2 == mm1 && m0 == mp1 && m0 != mm1) d = 0.5 * (mm1 + m0);
    else d = (mm1 * w_next + m0 * w_prev) / (w_next + w_prev);
    tan[i] = d;
  }

  // Hermite cubic on [x_i, x_{i+1}] expanded in absolute x (eph_spline.h:111-125);
  // the expression order below is the reference's and fixes the rounding.
  const double dx3 = dx * dx * dx;
  for (size_t i = 0; i + 1 < n; ++i) {
    const double p1 = i * dx, p2 = i * dx * p1, p3 = i * dx * p2;
    const double q1 = (i + 1) * dx, q2 = (i + 1) * dx * q1, q3 = (i + 1) * dx * q2;
    const double ta = tan[i], tb = tan[i + 1];
    const double d = (-ta * p1 - tb * p1 + ta * q1 + tb * q1 + 2.0 * y[i] - 2.0 * y[i + 1]) / dx3;
    const double c = (-ta + tb + 3.0 * d * p2 - 3.0 * d * q2) / 2.0 / dx;
    const double b = (c * p2 + d * p3 - c * q2 - d * q3 - y[i] + y[i + 1]) / dx;
    const double a = y[i] - b * p1 - c * p2 - d * p3;
    double *o = &t.k[4 * i];
    o[0] = a; o[1] = b; o[2] = c; o[3] = d;
  }
  t.k[4 * (n - 1)] = y[n - 1];
  return t;
}

// EPH_Linear (reference eph_linear.h): piecewise-linear table with inverse.
struct LinearTable {
  double dx = 0.0;
  std::vector<double> y;
};

struct BetaTables {
  int n_elements = 0;
  size_t n_rho = 0, n_beta = 0;
  double dr = 0, dr_sq = 0, drho = 0, r_cutoff = 0, r_cutoff_sq = 0, rho_cutoff = 0;
  std::vector<std::string> element_name;
  std::vector<int> element_number;
  std::vector<CubicTable> rho_r, rho_r_sq, alpha, beta;

  double inv_dr_sq() const { return rho_r_sq.at(0).inv_dx; }
  double inv_drho() const { return beta.at(0).inv_dx; }
  int find(const std::string &name) const {
    for (int e = 0; e < n_elements; ++e)
      if (element_name[e] == name) return e;
    return -1;
  }
  // [n_elements][n][4] flattened for eph_b200_set_tables
  static std::vector<double> flatten(const std::vector<CubicTable> &t) {
    std::vector<double> out;
    for (const auto &c : t) out.insert(out.end(), c.k.begin(), c.k.end());
    return out;
  }
};

inline void build_element_tables(BetaTables &b, std::vector<double> rho_knots, std::vector<double> beta_knots) {
  CubicTable rho = make_cubic_table(b.dr, rho_knots);
  for (size_t j = 0; j != b.n_rho; ++j) rho_knots[j] = rho(std::sqrt(j * b.dr_sq));  // eph_beta.h:108-109
  b.rho_r_sq.push_back(make_cubic_table(b.dr_sq, rho_knots));
  b.rho_r.push_back(std::move(rho));
  b.beta.push_back(make_cubic_table(b.drho, beta_knots));
  for (double &v : beta_knots) v = std::sqrt(v);  // alpha = sqrt(beta), eph_beta.h:120-121
  b.alpha.push_back(make_cubic_table(b.drho, beta_knots));
}

inline void set_beta_header(BetaTables &b, size_t n_rho, double dr, size_t n_beta, double drho, double r_cutoff) {
  b.n_rho = n_rho; b.n_beta = n_beta; b.dr = dr; b.drho = drho; b.r_cutoff = r_cutoff;
  b.r_cutoff_sq = r_cutoff * r_cutoff;
  b.rho_cutoff = drho * (n_beta - 1);
  b.dr_sq = b.r_cutoff_sq / (n_rho - 1);
}

// `.beta` grammar: 3 comment lines; "n_elements NAME..."; "n_rho dr n_beta drho r_cutoff";
// per element: Z, n_rho values of rho(r), n_beta values of beta(rho).
inline BetaTables load_beta_file(const std::string &path) {
  std::ifstream in(path);
  if (!in.is_open()) throw std::runtime_error("eph_b200: cannot open beta file '" + path + "'");
  std::string line;
  for (int k = 0; k < 3; ++k) std::getline(in, line);
  BetaTables b;
  if (!(in >> b.n_elements) || b.n_elements < 1) throw std::runtime_error("Fix eph: no elements found in input file");
  std::getline(in, line);
  std::istringstream names(line);
  b.element_name.resize(b.n_elements);
  for (auto &nm : b.element_name) names >> nm;
  size_t n_rho, n_beta;
  double dr, drho, rc;
  if (!(in >> n_rho >> dr >> n_beta >> drho >> rc)) throw std::runtime_error("eph_b200: bad beta file header in '" + path + "'");
  set_beta_header(b, n_rho, dr, n_beta, drho, rc);
  for (int e = 0; e < b.n_elements; ++e) {
    unsigned short z;
    in >> z;
    b.element_number.push_back(z);
    std::vector<double> r(n_rho), be(n_beta);
    for (auto &v : r) in >> v;
    for (auto &v : be) in >> v;
    if (!in) throw std::runtime_error("eph_b200: beta file '" + path + "' ends early");
    build_element_tables(b, std::move(r), std::move(be));
  }
  return b;
}

}  // namespace eph_b200
