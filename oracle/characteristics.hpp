/* characteristics.hpp -- CPU oracle for the characteristic decomposition behind the Riemann_invariants boundary condition.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle_impl.hpp): nothing under hexed_b200/ may include or call this.
 *
 * Follows reference include/pde.hpp:181-256 (`Navier_stokes<>::Pde<n_dim>::Characteristics`) and
 * src/Boundary_condition.cpp:82-182 (`apply_char`, `Riemann_invariants::apply_state / apply_flux`).
 *
 * Third-party arithmetic: the reference factorises the 3x3 eigenvector matrix with Eigen's
 * `ColPivHouseholderQR` (pde.hpp:188,225) and calls `.solve()` (pde.hpp:246). Eigen is NOT vendored in /root/reference and
 * its version is unpinned (no lockfile; 3.4-era API). `Qr3` below restates the PUBLISHED algorithm of Eigen 3.4's
 * ColPivHouseholderQR::computeInPlace / _solve_impl and MatrixBase::makeHouseholder / applyHouseholderOnTheLeft:
 * column pivoting on the running column norms with the LAPACK xGEQPF norm-downdate (LAWN 176), Householder vectors stored
 * essential-part-below-diagonal with tau = (beta - c0)/beta and beta = -sign(c0)*||x||, rank decision
 * `biggest_col_sq_norm < (max_norm*eps)^2/rows * (rows - k)`, least-squares "basic" solution with the columns beyond
 * nonzero_pivots set to zero. Pinned by the reference's own tests: test/test_Characteristics.cpp:4-42 (decomposition sums to
 * the state; columns are eigenvectors of the flux Jacobian) and test/test_Boundary_condition.cpp:75-117 (supersonic in/outflow).
 */
#ifndef HEXED_ORACLE_CHARACTERISTICS_HPP_
#define HEXED_ORACLE_CHARACTERISTICS_HPP_
#include <algorithm>
#include <cmath>
#include <limits>

namespace ho_impl {

struct Qr3
{
  double qr[3][3]; // row, column
  double tau[3];
  int perm[3];     // perm[i] = original column now at position i
  int nonzero_pivots;

  static double norm_tail(const double a[3][3], int col, int first)
  {
    double s = 0.;
    for (int r = first; r < 3; ++r) s += a[r][col]*a[r][col];
    return std::sqrt(s);
  }

  void compute(const double a[3][3])
  {
    const double eps = std::numeric_limits<double>::epsilon();
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) qr[r][c] = a[r][c];
    double updated[3], direct[3];
    for (int c = 0; c < 3; ++c) { direct[c] = norm_tail(qr, c, 0); updated[c] = direct[c]; }
    const double max_norm = std::max(std::max(updated[0], updated[1]), updated[2]);
    const double threshold_helper = (max_norm*eps)*(max_norm*eps)/3.;
    const double downdate_threshold = std::sqrt(eps);
    nonzero_pivots = 3;
    int transp[3];
    for (int k = 0; k < 3; ++k) {
      int big = k;
      for (int c = k + 1; c < 3; ++c) if (updated[c] > updated[big]) big = c; // first maximum wins, like maxCoeff(&index)
      const double big_sq = updated[big]*updated[big];
      if (nonzero_pivots == 3 && big_sq < threshold_helper*double(3 - k)) nonzero_pivots = k;
      transp[k] = big;
      if (big != k) {
        for (int r = 0; r < 3; ++r) std::swap(qr[r][k], qr[r][big]);
        std::swap(updated[k], updated[big]);
        std::swap(direct[k], direct[big]);
      }
      // makeHouseholderInPlace on qr[k..2][k]
      double tail_sq = 0.;
      for (int r = k + 1; r < 3; ++r) tail_sq += qr[r][k]*qr[r][k];
      const double c0 = qr[k][k];
      double beta;
      if (tail_sq <= std::numeric_limits<double>::min()) {
        tau[k] = 0.; beta = c0;
        for (int r = k + 1; r < 3; ++r) qr[r][k] = 0.;
      } else {
        beta = std::sqrt(c0*c0 + tail_sq);
        if (c0 >= 0.) beta = -beta;
        for (int r = k + 1; r < 3; ++r) qr[r][k] = qr[r][k]/(c0 - beta);
        tau[k] = (beta - c0)/beta;
      }
      qr[k][k] = beta;
      // apply H_k to the trailing columns
      if (k == 2) { /* no trailing columns */ }
      else if (tau[k] != 0.) {
        for (int c = k + 1; c < 3; ++c) {
          double tmp = 0.;
          for (int r = k + 1; r < 3; ++r) tmp += qr[r][k]*qr[r][c];
          tmp += qr[k][c];
          qr[k][c] -= tau[k]*tmp;
          for (int r = k + 1; r < 3; ++r) qr[r][c] -= tau[k]*qr[r][k]*tmp;
        }
      }
      // norm downdate
      for (int j = k + 1; j < 3; ++j) {
        if (updated[j] != 0.) {
          double temp = std::abs(qr[k][j])/updated[j];
          temp = (1. + temp)*(1. - temp);
          temp = temp < 0. ? 0. : temp;
          const double ratio = updated[j]/direct[j];
          const double temp2 = temp*ratio*ratio;
          if (temp2 <= downdate_threshold) {
            direct[j] = norm_tail(qr, j, k + 1);
            updated[j] = direct[j];
          } else updated[j] *= std::sqrt(temp);
        }
      }
    }
    for (int i = 0; i < 3; ++i) perm[i] = i;
    for (int k = 0; k < 3; ++k) std::swap(perm[k], perm[transp[k]]);
  }

  void solve(const double rhs[3], double x[3]) const
  {
    if (nonzero_pivots == 0) { x[0] = x[1] = x[2] = 0.; return; }
    double c[3] = {rhs[0], rhs[1], rhs[2]};
    for (int k = 0; k < nonzero_pivots; ++k) { // Q^T c = H_{p-1} ... H_1 H_0 c
      if (k == 2) c[2] *= 1. - tau[2]; // one-row block: scaled by 1 - tau
      else if (tau[k] != 0.) {
        double tmp = 0.;
        for (int r = k + 1; r < 3; ++r) tmp += qr[r][k]*c[r];
        tmp += c[k];
        c[k] -= tau[k]*tmp;
        for (int r = k + 1; r < 3; ++r) c[r] -= tau[k]*qr[r][k]*tmp;
      }
    }
    for (int i = nonzero_pivots - 1; i >= 0; --i) { // back substitution on the leading triangle
      double s = c[i];
      for (int j = i + 1; j < nonzero_pivots; ++j) s -= qr[i][j]*c[j];
      c[i] = s/qr[i][i];
    }
    for (int i = 0; i < nonzero_pivots; ++i) x[perm[i]] = c[i];
    for (int i = nonzero_pivots; i < 3; ++i) x[perm[i]] = 0.;
  }
};

template <int ND>
struct Characteristics
{
  static constexpr int NV = ND + 2;
  double vals[3];
  double vecs[3][3];
  Qr3 fact;
  double dir[ND];
  double mass;
  double veloc[ND];

  double nrml(const double* v) const { double s = 0.; for (int d = 0; d < ND; ++d) s += dir[d]*v[d]; return s; }
  void tang(const double* v, double* out) const { const double n = nrml(v); for (int d = 0; d < ND; ++d) out[d] = v[d] - dir[d]*n; }

  // pde.hpp:202-229
  Characteristics(const double* state, const double* direction)
  {
    double nsq = 0.;
    for (int d = 0; d < ND; ++d) nsq += direction[d]*direction[d];
    const double nrm = std::sqrt(nsq);
    for (int d = 0; d < ND; ++d) dir[d] = direction[d]/nrm;
    mass = state[ND];
    for (int d = 0; d < ND; ++d) veloc[d] = state[d]/mass;
    double vsq = 0.;
    for (int d = 0; d < ND; ++d) vsq += veloc[d]*veloc[d];
    const double pres = .4*(state[ND + 1] - .5*mass*vsq);
    const double sound_speed = std::sqrt(1.4*std::max(pres, 0.)/mass);
    vals[2] = nrml(veloc);
    vals[0] = vals[2] - sound_speed;
    vals[1] = vals[2] + sound_speed;
    const double d_mass = 1;
    for (int sign = 0; sign < 2; ++sign) {
      const double d_veloc = (2*sign - 1)*sound_speed/mass*d_mass;
      const double d_pres = 1.4*pres/mass*d_mass;
      vecs[0][sign] = d_mass*vals[2] + mass*d_veloc;
      vecs[1][sign] = d_mass;
      vecs[2][sign] = d_pres/.4 + .5*d_mass*vsq + mass*vals[2]*d_veloc;
    }
    vecs[0][2] = d_mass*vals[2];
    vecs[1][2] = d_mass;
    vecs[2][2] = .5*d_mass*vsq;
    fact.compute(vecs);
  }

  // pde.hpp:236-254; d[var][eig]
  void decomp(const double* state, double d[NV][3]) const
  {
    double mmtm[ND], corr[ND], tv[ND], tm[ND];
    for (int i = 0; i < ND; ++i) mmtm[i] = state[i];
    tang(mmtm, tm);
    tang(veloc, tv);
    for (int i = 0; i < ND; ++i) corr[i] = tm[i] - state[ND]*tv[i];
    double vdc = 0.;
    for (int i = 0; i < ND; ++i) vdc += veloc[i]*corr[i];
    const double state_1d[3] = {nrml(mmtm), state[ND], state[ND + 1] - vdc};
    double eig_basis[3];
    fact.solve(state_1d, eig_basis);
    for (int j = 0; j < 3; ++j) {
      const double e0 = vecs[0][j]*eig_basis[j], e1 = vecs[1][j]*eig_basis[j], e2 = vecs[2][j]*eig_basis[j];
      d[ND][j] = e1;
      d[ND + 1][j] = e2;
      for (int i = 0; i < ND; ++i) d[i][j] = dir[i]*e0 + tv[i]*eig_basis[j];
    }
    for (int i = 0; i < ND; ++i) d[i][2] += corr[i];
    d[ND + 1][2] += vdc;
  }
};

// src/Boundary_condition.cpp:82-95
template <int ND>
void apply_char(const double* state, const double* normal, int sign, const double* inside, const double* outside, double* result)
{
  constexpr int NV = ND + 2;
  Characteristics<ND> ch(state, normal);
  double dec[NV][3], out[NV][3];
  ch.decomp(inside, dec);
  ch.decomp(outside, out);
  for (int j = 0; j < 3; ++j) if (sign*ch.vals[j] > 0) for (int v = 0; v < NV; ++v) dec[v][j] = out[v][j];
  for (int v = 0; v < NV; ++v) result[v] = (dec[v][0] + dec[v][1]) + dec[v][2];
}

// Riemann_invariants::apply_state for one face point, src/Boundary_condition.cpp:106-132
template <int ND>
void riemann_state_point(const double* inside, const double* normal, int sign, const double* fs, double* ghost)
{
  apply_char<ND>(inside, normal, sign, inside, fs, ghost);
  ghost[ND] = std::max(ghost[ND], inside[ND]/2);
  double gsq = 0., isq = 0.;
  for (int d = 0; d < ND; ++d) { gsq += ghost[d]*ghost[d]; isq += inside[d]*inside[d]; }
  const double kin_ener = .5*gsq/ghost[ND];
  const double inside_kin_ener = .5*isq/inside[ND];
  ghost[ND + 1] = std::max(kin_ener + std::max(ghost[ND + 1] - kin_ener, (inside[ND + 1] - inside_kin_ener)/2), 0.);
}

// Riemann_invariants::apply_flux for one face point, src/Boundary_condition.cpp:152-180
template <int ND>
void riemann_flux_point(const double* cache, const double* normal, int sign, const double* inside_flux, double* ghost_flux)
{
  double zero[ND + 2];
  for (int v = 0; v < ND + 2; ++v) zero[v] = 0.;
  apply_char<ND>(cache, normal, sign, zero, inside_flux, ghost_flux);
}

} // namespace ho_impl
#endif
