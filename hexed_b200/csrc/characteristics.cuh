// characteristics.cuh -- characteristic decomposition for the Riemann_invariants boundary condition, one face point per thread.
// Replaces reference include/pde.hpp:181-256 (`Characteristics`: eigenvalues, 3x3 eigenvector matrix, least-squares-capable
// factorisation, `decomp`) and src/Boundary_condition.cpp:82-95 (`apply_char`). The reference factorises with Eigen's
// ColPivHouseholderQR; PivQr3 is the same algorithm (pivoting on running column norms with the LAPACK norm downdate,
// Householder reflectors, rank threshold (max_norm*eps)^2/rows*(rows-k), zeroed free variables) held entirely in registers:
// every loop is unrolled over compile-time indices and the data-dependent column exchange is a predicated swap.
#pragma once
#include <cfloat>

namespace hb {

struct PivQr3
{
  double a[3][3]; // [row][col], R above the diagonal, essential parts of the reflectors below
  double tau[3];
  int perm[3];
  int rank;

  __device__ __forceinline__ static void swap_d(double& x, double& y) { const double t = x; x = y; y = t; }

  template <int K>
  __device__ __forceinline__ void step(double (&upd)[3], double (&dir)[3], double threshold_helper)
  {
    // pivot: first largest running norm among columns K..2
    int big = K;
    double big_norm = upd[K];
    #pragma unroll
    for (int c = K + 1; c < 3; ++c) if (upd[c] > big_norm) { big_norm = upd[c]; big = c; }
    if (rank == 3 && big_norm*big_norm < threshold_helper*double(3 - K)) rank = K;
    #pragma unroll
    for (int c = K + 1; c < 3; ++c) {
      if (big == c) {
        #pragma unroll
        for (int r = 0; r < 3; ++r) swap_d(a[r][K], a[r][c]);
        swap_d(upd[K], upd[c]);
        swap_d(dir[K], dir[c]);
        const int t = perm[K]; perm[K] = perm[c]; perm[c] = t;
      }
    }
    // reflector for a[K..2][K]
    double tail_sq = 0.;
    #pragma unroll
    for (int r = K + 1; r < 3; ++r) tail_sq += a[r][K]*a[r][K];
    const double c0 = a[K][K];
    double beta;
    if (tail_sq <= DBL_MIN) {
      tau[K] = 0.; beta = c0;
      #pragma unroll
      for (int r = K + 1; r < 3; ++r) a[r][K] = 0.;
    } else {
      beta = sqrt(c0*c0 + tail_sq);
      if (c0 >= 0.) beta = -beta;
      const double den = c0 - beta;
      #pragma unroll
      for (int r = K + 1; r < 3; ++r) a[r][K] = a[r][K]/den;
      tau[K] = (beta - c0)/beta;
    }
    a[K][K] = beta;
    if (K < 2 && tau[K] != 0.) {
      #pragma unroll
      for (int c = K + 1; c < 3; ++c) {
        double tmp = 0.;
        #pragma unroll
        for (int r = K + 1; r < 3; ++r) tmp += a[r][K]*a[r][c];
        tmp += a[K][c];
        a[K][c] -= tau[K]*tmp;
        #pragma unroll
        for (int r = K + 1; r < 3; ++r) a[r][c] -= tau[K]*a[r][K]*tmp;
      }
    }
    #pragma unroll
    for (int j = K + 1; j < 3; ++j) {
      if (upd[j] != 0.) {
        double temp = fabs(a[K][j])/upd[j];
        temp = (1. + temp)*(1. - temp);
        temp = temp < 0. ? 0. : temp;
        const double ratio = upd[j]/dir[j];
        if (temp*ratio*ratio <= 1.4901161193847656e-08 /* sqrt(DBL_EPSILON) */) {
          double s = 0.;
          #pragma unroll
          for (int r = K + 1; r < 3; ++r) s += a[r][j]*a[r][j];
          dir[j] = sqrt(s);
          upd[j] = dir[j];
        } else upd[j] *= sqrt(temp);
      }
    }
  }

  __device__ __forceinline__ void compute()
  {
    double upd[3], dir[3];
    #pragma unroll
    for (int c = 0; c < 3; ++c) {
      dir[c] = sqrt(a[0][c]*a[0][c] + a[1][c]*a[1][c] + a[2][c]*a[2][c]);
      upd[c] = dir[c];
      perm[c] = c;
    }
    const double scaled = fmax(fmax(upd[0], upd[1]), upd[2])*DBL_EPSILON;
    const double threshold_helper = scaled*scaled/3.;
    rank = 3;
    step<0>(upd, dir, threshold_helper);
    step<1>(upd, dir, threshold_helper);
    step<2>(upd, dir, threshold_helper);
  }

  // least-squares solution with the free variables (columns beyond `rank`) zero
  __device__ __forceinline__ void solve(const double (&rhs)[3], double (&x)[3]) const
  {
    double c[3] = {rhs[0], rhs[1], rhs[2]};
    if (rank > 0 && tau[0] != 0.) {
      const double tmp = a[1][0]*c[1] + a[2][0]*c[2] + c[0];
      c[0] -= tau[0]*tmp; c[1] -= tau[0]*a[1][0]*tmp; c[2] -= tau[0]*a[2][0]*tmp;
    }
    if (rank > 1 && tau[1] != 0.) {
      const double tmp = a[2][1]*c[2] + c[1];
      c[1] -= tau[1]*tmp; c[2] -= tau[1]*a[2][1]*tmp;
    }
    if (rank > 2) c[2] *= 1. - tau[2];
    if (rank > 2) c[2] = c[2]/a[2][2];
    if (rank > 1) c[1] = (rank > 2 ? c[1] - a[1][2]*c[2] : c[1])/a[1][1];
    if (rank > 0) {
      double s = c[0];
      if (rank > 1) s -= a[0][1]*c[1];
      if (rank > 2) s -= a[0][2]*c[2];
      c[0] = s/a[0][0];
    }
    #pragma unroll
    for (int i = 0; i < 3; ++i) if (i >= rank) c[i] = 0.;
    #pragma unroll
    for (int j = 0; j < 3; ++j) x[j] = perm[0] == j ? c[0] : perm[1] == j ? c[1] : c[2];
  }
};

template <int ND>
struct Characteristics
{
  static constexpr int NV = ND + 2;
  double vals[3];
  double vecs[3][3];
  PivQr3 fact;
  double dir[ND], veloc[ND], tang_veloc[ND];
  double mass;

  __device__ __forceinline__ double nrml(const double (&v)[ND]) const
  {
    double s = 0.;
    #pragma unroll
    for (int d = 0; d < ND; ++d) s += dir[d]*v[d];
    return s;
  }

  __device__ __forceinline__ Characteristics(const double (&state)[NV], const double (&direction)[ND])
  {
    double nsq = 0.;
    #pragma unroll
    for (int d = 0; d < ND; ++d) nsq += direction[d]*direction[d];
    const double nrm = sqrt(nsq);
    mass = state[ND];
    double vsq = 0.;
    #pragma unroll
    for (int d = 0; d < ND; ++d) { dir[d] = direction[d]/nrm; veloc[d] = state[d]/mass; vsq += veloc[d]*veloc[d]; }
    const double pres = .4*(state[ND + 1] - .5*mass*vsq);
    const double sound_speed = sqrt(1.4*fmax(pres, 0.)/mass);
    vals[2] = nrml(veloc);
    vals[0] = vals[2] - sound_speed;
    vals[1] = vals[2] + sound_speed;
    #pragma unroll
    for (int d = 0; d < ND; ++d) tang_veloc[d] = veloc[d] - dir[d]*vals[2];
    #pragma unroll
    for (int sign = 0; sign < 2; ++sign) {
      const double d_veloc = (2*sign - 1)*sound_speed/mass;
      const double d_pres = 1.4*pres/mass;
      vecs[0][sign] = vals[2] + mass*d_veloc;
      vecs[1][sign] = 1.;
      vecs[2][sign] = d_pres/.4 + .5*vsq + mass*vals[2]*d_veloc;
    }
    vecs[0][2] = vals[2]; vecs[1][2] = 1.; vecs[2][2] = .5*vsq;
    #pragma unroll
    for (int r = 0; r < 3; ++r) {
      #pragma unroll
      for (int c = 0; c < 3; ++c) fact.a[r][c] = vecs[r][c];
    }
    fact.compute();
  }

  // column j of the decomposition of `state` is out[.][j]
  __device__ __forceinline__ void decomp(const double (&state)[NV], double (&out)[NV][3]) const
  {
    double mmtm[ND], corr[ND];
    #pragma unroll
    for (int d = 0; d < ND; ++d) mmtm[d] = state[d];
    const double mn = nrml(mmtm);
    double vdc = 0.;
    #pragma unroll
    for (int d = 0; d < ND; ++d) { corr[d] = (mmtm[d] - dir[d]*mn) - state[ND]*tang_veloc[d]; vdc += veloc[d]*corr[d]; }
    const double rhs[3] = {mn, state[ND], state[ND + 1] - vdc};
    double eb[3];
    fact.solve(rhs, eb);
    #pragma unroll
    for (int j = 0; j < 3; ++j) {
      const double e0 = vecs[0][j]*eb[j];
      #pragma unroll
      for (int d = 0; d < ND; ++d) out[d][j] = dir[d]*e0 + tang_veloc[d]*eb[j];
      out[ND][j] = vecs[1][j]*eb[j];
      out[ND + 1][j] = vecs[2][j]*eb[j];
    }
    #pragma unroll
    for (int d = 0; d < ND; ++d) out[d][2] += corr[d];
    out[ND + 1][2] += vdc;
  }
};

// eigenspaces whose characteristic velocity points inward (sign*eigenvalue > 0) take the `outside` value
template <int ND>
__device__ __forceinline__ void apply_char(const double (&state)[ND + 2], const double (&normal)[ND], int sign,
                                           const double (&inside)[ND + 2], const double (&outside)[ND + 2], double (&result)[ND + 2])
{
  const Characteristics<ND> ch(state, normal);
  double in_d[ND + 2][3], out_d[ND + 2][3];
  ch.decomp(inside, in_d);
  ch.decomp(outside, out_d);
  #pragma unroll
  for (int v = 0; v < ND + 2; ++v) {
    double cols[3];
    #pragma unroll
    for (int j = 0; j < 3; ++j) cols[j] = sign*ch.vals[j] > 0 ? out_d[v][j] : in_d[v][j];
    result[v] = (cols[0] + cols[1]) + cols[2];
  }
}

} // namespace hb
