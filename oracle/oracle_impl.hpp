/* oracle_impl.hpp -- CPU oracle for the Hexed per-stage DG residual update.
 *
 * TEST INFRASTRUCTURE ONLY. Nothing under hexed_b200/ may call into this file; it is
 * the checker used by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline leg.
 *
 * What it is: an Eigen-free C++/OpenMP restatement of the reference's kernels, written
 * from the algorithm (not from the source text) against the flattened mesh of
 * flat_mesh.h. Each function cites the reference lines whose behaviour it follows
 * (paths relative to /root/reference). The reference itself cannot be built in this
 * image (Eigen, HDF5 and Catch2 are absent), so parity is pinned by re-expressing the
 * reference's own closed-form test values (test/test_Max_dt.cpp, test_Derivative.cpp,
 * test_Prolong_refined.cpp, test_Restrict_refined.cpp, test_Face_permutation.cpp, ...)
 * as tests/ of this repo; the Euler/NS flux has no direct unit test in the reference
 * (test/test_pde.cpp is #if 0) and is pinned through conservation and the reference's analytic
 * marching check (test/test_Solver.cpp:555-586, tests/test_oracle_kat.py::test_marching_residual);
 * the LDG path (Neighbor average, gradient, compute_flux_diff, Neighbor_reconcile,
 * Reconcile_ldg_flux) is pinned by the reference's analytic viscous-decay check
 * (test/test_Solver.cpp:588-615, tests/test_oracle_kat.py::test_viscous_momentum_decay).
 * PARITY UNPINNED beyond line-by-line restatement: Stab_art_visc and Fix_therm_admis
 * (the reference has no dedicated test for either).
 *
 * Floating point: plain IEEE double, sums accumulated left to right in the order the
 * reference's loops imply; build with -ffp-contract=off for a machine-independent
 * answer (the Makefile also builds a contract=fast variant to measure FMA sensitivity).
 */
#ifndef HEXED_ORACLE_IMPL_HPP_
#define HEXED_ORACLE_IMPL_HPP_
#include "flat_mesh.h"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace ho_impl {

constexpr int ipow(int b, int e) { int r = 1; for (int i = 0; i < e; ++i) r *= b; return r; }
constexpr int cmax(int a, int b) { return a > b ? a : b; }

// reference include/pde.hpp:17-21
constexpr int tss_offset(int nd) { return nd + 2; }
constexpr int bulk_av_offset(int nd) { return nd + 3; }
constexpr int laplacian_av_offset(int nd) { return nd + 4; }
constexpr int forcing_offset(int nd) { return nd + 5; }
constexpr int advection_offset(int nd) { return nd + 9; }
// reference src/Element.cpp:188 -- residual cache sits after state, tss, 2 AV coefs, 4 forcing and row_size advection slots
inline int cache_offset(int nd, int rs) { return (nd + 2) + 3 + 4 + rs; }

constexpr double heat_rat = 1.4;              // include/pde.hpp:51
constexpr double specific_gas_air = 287.05287; // include/constants.hpp:49

inline double transport_coef(const ho_transport& t, double sqrt_temp)
{
  // include/Transport_model.hpp:35-38; math::pow(x, 3) multiplies 1*x*x*x left to right (include/math.hpp:29-36)
  double r = sqrt_temp/t.sqrt_ref_temp;
  double cube = 1; for (int i = 0; i < 3; ++i) cube *= r;
  return t.const_val + t.ref_val*cube*(t.ref_temp + t.temp_offset)/(sqrt_temp*sqrt_temp + t.temp_offset);
}

/* ---- row iteration: reference include/Row_index.hpp:32-63 ---- */
template <int ND, int RS>
struct Rows
{
  static constexpr int nq = ipow(RS, ND);
  static constexpr int nfq = nq/RS;
  // index of node `i_node` of row `i_fq` along dimension `i_dim`
  static int qpoint(int i_dim, int i_fq, int i_node)
  {
    const int stride = ipow(RS, ND - 1 - i_dim);
    const int i_outer = i_fq/stride, i_inner = i_fq%stride;
    return i_outer*stride*RS + i_inner + i_node*stride;
  }
};

/* ---- 1-D DG derivative: reference include/Derivative.hpp:14-61 ---- */
template <int RS>
struct Deriv
{
  double diff[RS][RS], bnd[2][RS], lift[RS][2];
  explicit Deriv(const ho_basis& b)
  {
    for (int i = 0; i < RS; ++i) {
      for (int j = 0; j < RS; ++j) diff[i][j] = b.diff_mat[i][j];
      for (int s = 0; s < 2; ++s) bnd[s][i] = b.boundary[s][i];
      const double inv_w = 1./b.weight[i];
      lift[i][0] = inv_w*b.boundary[0][i]*-1.;
      lift[i][1] = inv_w*b.boundary[1][i]*1.;
    }
  }
  // diff_mat*q + lift*(bv - boundary*q), one variable at a time
  void full(const double* q, const double* bv, double* out) const
  {
    double jump[2];
    for (int s = 0; s < 2; ++s) {
      double e = 0; for (int j = 0; j < RS; ++j) e += bnd[s][j]*q[j];
      jump[s] = bv[s] - e;
    }
    for (int i = 0; i < RS; ++i) {
      double d = 0; for (int j = 0; j < RS; ++j) d += diff[i][j]*q[j];
      double l = 0; for (int s = 0; s < 2; ++s) l += lift[i][s]*jump[s];
      out[i] = d + l;
    }
  }
  void interior(const double* q, double* out) const
  {
    for (int i = 0; i < RS; ++i) { double d = 0; for (int j = 0; j < RS; ++j) d += diff[i][j]*q[j]; out[i] = d; }
  }
  void boundary_term(const double* bv, double* out) const
  {
    for (int i = 0; i < RS; ++i) { double l = 0; for (int s = 0; s < 2; ++s) l += lift[i][s]*bv[s]; out[i] = l; }
  }
};

/* ===================== PDE definitions (reference include/pde.hpp) ===================== */

// pde.hpp:27-175
template <int ND, int RS, bool VISC>
struct Pde_ns
{
  static constexpr bool has_diffusion = VISC, has_convection = true, has_source = false;
  static constexpr int n_update = ND + 2, n_state = ND + 4, n_extrap = ND + 2;
  static constexpr int face_kind = 0; // uses face_state / face_ldg
  ho_transport dyn_visc, therm_cond;

  void fetch_extrap(int stride, const double* data, double* out) const
  { for (int v = 0; v < n_extrap; ++v) out[v] = data[v*stride]; }
  void write_update(const double* upd, int stride, double* data, bool) const
  { for (int v = 0; v < n_update; ++v) data[v*stride] += upd[v]; }

  template <int NDF>
  struct Comp
  {
    const Pde_ns& eq;
    explicit Comp(const Pde_ns& e) : eq{e}
    { for (int i = 0; i < ND; ++i) for (int j = 0; j < NDF; ++j) normal[i][j] = (i == j); }
    double state[n_state];
    double update_state[n_update];
    double normal[ND][NDF];
    double flux_conv[n_update][NDF];
    double gradient[n_extrap][ND];
    double flux_diff[n_update][ND];
    double source[n_update];
    double mass, kin_ener, pressure;
    double bulk_av, laplacian_av, sqrt_temp, dyn_visc_coef, therm_cond_coef, energy_cond;
    double char_speed, diffusivity;

    void fetch_state(int stride, const double* data)
    {
      for (int v = 0; v < ND + 2; ++v) state[v] = data[v*stride];
      state[ND + 2] = data[bulk_av_offset(ND)*stride];
      state[ND + 3] = data[laplacian_av_offset(ND)*stride];
    }
    void fetch_extrap_state(int stride, const double* data)
    {
      for (int v = 0; v < ND + 2; ++v) state[v] = data[v*stride];
      state[ND + 2] = 0.; state[ND + 3] = 0.;
      for (int v = 0; v < n_update; ++v) update_state[v] = state[v];
    }
    void scalars_conv()
    {
      mass = state[ND];
      kin_ener = 0;
      for (int i = 0; i < ND; ++i) kin_ener += state[i]*state[i];
      kin_ener *= .5/mass;
      pressure = (heat_rat - 1.)*(state[ND + 1] - kin_ener);
    }
    void compute_flux_conv()
    {
      scalars_conv();
      for (int d = 0; d < NDF; ++d) {
        double mass_flux = 0;
        for (int j = 0; j < ND; ++j) mass_flux += state[j]*normal[j][d];
        flux_conv[ND][d] = mass_flux;
        const double vol_flux = mass_flux/mass;
        flux_conv[ND + 1][d] = (state[ND + 1] + pressure)*vol_flux;
        for (int j = 0; j < ND; ++j) flux_conv[j][d] = state[j]*vol_flux + pressure*normal[j][d];
      }
    }
    void scalars_diff()
    {
      bulk_av = std::abs(state[ND + 2]);
      laplacian_av = std::abs(state[ND + 3]);
      sqrt_temp = std::sqrt(std::max((state[ND + 1] - kin_ener)/mass, 0.)*(heat_rat - 1)/specific_gas_air);
      dyn_visc_coef = transport_coef(eq.dyn_visc, sqrt_temp);
      therm_cond_coef = transport_coef(eq.therm_cond, sqrt_temp);
      energy_cond = therm_cond_coef*(heat_rat - 1)/specific_gas_air;
    }
    void compute_flux_diff()
    {
      static_assert(NDF == ND || true, "");
      if constexpr (NDF == ND) {
        scalars_diff();
        double veloc[ND], vgrad[ND][ND], stress[ND][ND];
        for (int i = 0; i < ND; ++i) veloc[i] = state[i]/mass;
        for (int i = 0; i < ND; ++i) for (int j = 0; j < ND; ++j) vgrad[i][j] = (gradient[i][j] - veloc[i]*gradient[ND][j])/mass;
        double trace = 0; for (int i = 0; i < ND; ++i) trace += vgrad[i][i];
        const double bulk = (bulk_av*mass - 2./3.*dyn_visc_coef)*trace;
        for (int i = 0; i < ND; ++i) for (int j = 0; j < ND; ++j) stress[i][j] = dyn_visc_coef*(vgrad[i][j] + vgrad[j][i]) + bulk*(i == j);
        double fd[n_update][ND];
        for (int v = 0; v < n_update; ++v) for (int j = 0; j < ND; ++j) fd[v][j] = -laplacian_av*gradient[v][j];
        for (int i = 0; i < ND; ++i) for (int j = 0; j < ND; ++j) fd[i][j] -= stress[i][j];
        for (int j = 0; j < ND; ++j) {
          double conv = 0; for (int i = 0; i < ND; ++i) conv += veloc[i]*vgrad[i][j];
          const double int_ener_grad = -state[ND + 1]/mass/mass*gradient[ND][j] + gradient[ND + 1][j]/mass - conv;
          double work = 0; for (int i = 0; i < ND; ++i) work += veloc[i]*stress[i][j];
          fd[ND + 1][j] -= work + energy_cond*int_ener_grad;
        }
        for (int v = 0; v < n_update; ++v) for (int k = 0; k < ND; ++k) {
          double s = 0; for (int j = 0; j < ND; ++j) s += fd[v][j]*normal[j][k];
          flux_diff[v][k] = s;
        }
      }
    }
    void compute_char_speed()
    {
      const double sound_speed = std::sqrt(heat_rat*(heat_rat - 1)*state[ND + 1]/state[ND]);
      double sq = 0; for (int i = 0; i < ND; ++i) sq += state[i]*state[i];
      char_speed = sound_speed + std::sqrt(sq)/state[ND];
    }
    void compute_diffusivity()
    {
      scalars_conv();
      scalars_diff();
      diffusivity = std::abs(laplacian_av) + std::max(std::abs(bulk_av) + dyn_visc_coef/mass, energy_cond/mass);
    }
    void compute_source() {}
  };
};

// pde.hpp:265-349
template <int ND, int RS>
struct Pde_advection
{
  static constexpr bool has_diffusion = false, has_convection = true, has_source = true;
  static constexpr int n_adv = RS;
  static constexpr int n_state = ND + n_adv, n_extrap = ND + n_adv, n_update = n_adv;
  static constexpr int face_kind = 2; // face_wide
  double advect_length;
  double nodes[RS]; // Gauss-Legendre nodes mapped to [-1, 1]

  void fetch_extrap(int stride, const double* data, double* out) const
  {
    for (int v = 0; v < ND; ++v) out[v] = data[v*stride];
    for (int a = 0; a < n_adv; ++a) out[ND + a] = data[(advection_offset(ND) + a)*stride];
  }
  void write_update(const double* upd, int stride, double* data, bool critical) const
  {
    const double pseudo = 1 + data[tss_offset(ND)*stride]*2/advect_length;
    for (int a = 0; a < n_adv; ++a) {
      double& d = data[(advection_offset(ND) + a)*stride];
      if (critical) d = (d + upd[a])/pseudo;
      else d += upd[a]/pseudo;
    }
  }
  template <int NDF>
  struct Comp
  {
    const Pde_advection& eq;
    explicit Comp(const Pde_advection& e) : eq{e}
    { for (int i = 0; i < ND; ++i) for (int j = 0; j < NDF; ++j) normal[i][j] = (i == j); }
    double state[n_state], update_state[n_update], normal[ND][NDF], flux_conv[n_update][NDF];
    double gradient[n_extrap][ND], flux_diff[n_update][ND];
    double source[n_update], char_speed, diffusivity;
    void fetch_state(int stride, const double* data) { eq.fetch_extrap(stride, data, state); }
    void fetch_extrap_state(int stride, const double* data)
    {
      for (int v = 0; v < n_extrap; ++v) state[v] = data[v*stride];
      for (int a = 0; a < n_adv; ++a) update_state[a] = state[ND + a];
    }
    void compute_flux_conv()
    {
      for (int d = 0; d < NDF; ++d) {
        double nv = 0; for (int j = 0; j < ND; ++j) nv += state[j]*normal[j][d];
        for (int a = 0; a < n_adv; ++a) flux_conv[a][d] = eq.nodes[a]*nv*state[ND + a];
      }
    }
    void compute_flux_diff() {}
    void compute_source() { for (int a = 0; a < n_update; ++a) source[a] = 2/eq.advect_length; }
    void compute_char_speed()
    {
      double sq = 0; for (int i = 0; i < ND; ++i) sq += state[i]*state[i];
      char_speed = std::max(1., std::sqrt(sq));
    }
    void compute_diffusivity() {}
  };
};

// pde.hpp:355-431
template <int ND, int RS>
struct Pde_smooth_av
{
  static constexpr bool has_diffusion = true, has_convection = false, has_source = true;
  static constexpr int n_state = 4, n_extrap = 3, n_update = 3;
  static constexpr int face_kind = 0;
  double diff_time, cheby;
  void fetch_extrap(int stride, const double* data, double* out) const
  { for (int v = 0; v < n_extrap; ++v) out[v] = data[(forcing_offset(ND) + 1 + v)*stride]; }
  void write_update(const double* upd, int stride, double* data, bool critical) const
  {
    const double pseudo = 1 + data[tss_offset(ND)*stride]*cheby/diff_time;
    for (int v = 0; v < n_update; ++v) {
      double& d = data[(forcing_offset(ND) + 1 + v)*stride];
      d += upd[v];
      if (critical) d /= pseudo;
    }
  }
  template <int NDF>
  struct Comp
  {
    const Pde_smooth_av& eq;
    explicit Comp(const Pde_smooth_av& e) : eq{e}
    { for (int i = 0; i < ND; ++i) for (int j = 0; j < NDF; ++j) normal[i][j] = (i == j); }
    double state[n_state], update_state[n_update], normal[ND][NDF], flux_conv[n_update][NDF];
    double gradient[n_extrap][ND], flux_diff[n_update][ND];
    double source[n_update], char_speed, diffusivity;
    void fetch_state(int stride, const double* data)
    { for (int v = 0; v < n_state; ++v) state[v] = data[(forcing_offset(ND) + v)*stride]; }
    void fetch_extrap_state(int stride, const double* data)
    { for (int v = 0; v < n_update; ++v) update_state[v] = data[v*stride]; }
    void compute_flux_conv() {}
    void compute_flux_diff()
    {
      if constexpr (NDF == ND) {
        for (int v = 0; v < n_update; ++v) for (int k = 0; k < ND; ++k) {
          double s = 0; for (int j = 0; j < ND; ++j) s += -gradient[v][j]*normal[j][k];
          flux_diff[v][k] = s;
        }
      }
    }
    void compute_source()
    {
      for (int v = 0; v < n_update; ++v) {
        const double f = std::abs(state[v]);
        source[v] = ((v == 1) ? std::sqrt(f) : f)/eq.diff_time;
      }
    }
    void compute_char_speed() {}
    void compute_diffusivity() { diffusivity = 1; }
  };
};

// pde.hpp:437-493
template <int ND, int RS>
struct Pde_fta
{
  static constexpr bool has_diffusion = true, has_convection = false, has_source = false;
  static constexpr int n_state = ND + 2, n_update = ND + 2, n_extrap = ND + 2;
  static constexpr int face_kind = 0;
  void fetch_extrap(int stride, const double* data, double* out) const
  { for (int v = 0; v < n_extrap; ++v) out[v] = data[v*stride]; }
  void write_update(const double* upd, int stride, double* data, bool) const
  { for (int v = 0; v < n_update; ++v) data[v*stride] += upd[v]; }
  template <int NDF>
  struct Comp
  {
    explicit Comp(const Pde_fta&)
    { for (int i = 0; i < ND; ++i) for (int j = 0; j < NDF; ++j) normal[i][j] = (i == j); }
    double state[n_state], update_state[n_update], normal[ND][NDF], flux_conv[n_update][NDF];
    double gradient[n_extrap][ND], flux_diff[n_update][ND];
    double source[n_update], char_speed, diffusivity;
    void fetch_state(int stride, const double* data) { for (int v = 0; v < n_state; ++v) state[v] = data[v*stride]; }
    void fetch_extrap_state(int stride, const double* data) { for (int v = 0; v < n_state; ++v) update_state[v] = data[v*stride]; }
    void compute_flux_conv() {}
    void compute_flux_diff()
    {
      if constexpr (NDF == ND) {
        for (int v = 0; v < n_update; ++v) for (int k = 0; k < ND; ++k) {
          double s = 0; for (int j = 0; j < ND; ++j) s += -gradient[v][j]*normal[j][k];
          flux_diff[v][k] = s;
        }
      }
    }
    void compute_source() {}
    void compute_char_speed() {}
    void compute_diffusivity() { diffusivity = 1; }
  };
};

/* ===================== mesh accessors ===================== */

template <int ND, int RS>
struct Geo
{
  static constexpr int nq = ipow(RS, ND), nfq = nq/RS, nv = ND + 2, n_face = 2*ND;
  ho_mesh& m;
  explicit Geo(ho_mesh& mesh) : m{mesh} {}
  int n_elem() const { return m.n_car + m.n_def; }
  double* elem(int e) const { return m.elem_data + (size_t)e*m.n_slot*nq; }
  double* state(int e) const { return elem(e); }
  double* tss(int e) const { return elem(e) + (size_t)tss_offset(ND)*nq; }
  double* cache(int e) const { return elem(e) + (size_t)cache_offset(ND, RS)*nq; }
  double* ref_nrml(int e) const { return m.ref_normals + (size_t)(e - m.n_car)*ND*ND*nq; }
  double* det(int e) const { return m.det + (size_t)(e - m.n_car)*nq; }
  double* elem_face_nrml(int e, int f) const { return m.normals + ((size_t)(e - m.n_car)*n_face + f)*ND*nfq; }
  double* nrml(int slot) const { return m.normals + (size_t)slot*ND*nfq; }
  // face storage of a given kind: 0 state, 1 ldg, 2 wide
  static int width(int kind) { return kind == 2 ? (ND + RS)*nfq : nv*nfq; }
  double* face(int kind, int slot) const
  {
    double* base = kind == 0 ? m.face_state : kind == 1 ? m.face_ldg : m.face_wide;
    return base + (size_t)slot*width(kind);
  }
};

/* ---- face permutation: reference include/Spatial.hpp:73-131, include/Kernel_connection.hpp:7-37 ---- */
struct Dir
{
  int i_dim[2]; int sign[2];
  bool flip_normal(int side) const { return sign[side] == side; }
  bool flip_tangential() const { return (i_dim[0] != i_dim[1]) && (flip_normal(0) == flip_normal(1)); }
  bool transpose() const { return (i_dim[0] == 0 && i_dim[1] == 2) || (i_dim[0] == 2 && i_dim[1] == 0); }
};

inline void perm_transpose(int nd, int rs, int n_var, const Dir& dir, double* tgt)
{
  if (nd != 3 || !dir.transpose()) return;
  const int nfq = rs*rs;
  for (int v = 0; v < n_var; ++v) {
    double* f = tgt + v*nfq;
    for (int a = 0; a < rs; ++a) for (int b = a + 1; b < rs; ++b) std::swap(f[a*rs + b], f[b*rs + a]);
  }
}
inline void perm_flip(int nd, int rs, int n_var, const Dir& dir, double* tgt)
{
  if (!dir.flip_tangential()) return;
  if (nd == 3) {
    const int nfq = rs*rs;
    // reversal along the fastest-varying face index <=> Eigen's colwise().reverse() on the column-major map
    const bool fast = (dir.i_dim[0] > 3 - dir.i_dim[0] - dir.i_dim[1]) != dir.transpose();
    for (int v = 0; v < n_var; ++v) {
      double* f = tgt + v*nfq;
      if (fast) { for (int a = 0; a < rs; ++a) std::reverse(f + a*rs, f + (a + 1)*rs); }
      else { for (int a = 0; a < rs/2; ++a) for (int b = 0; b < rs; ++b) std::swap(f[a*rs + b], f[(rs - 1 - a)*rs + b]); }
    }
  } else if (nd == 2) {
    for (int v = 0; v < n_var; ++v) std::reverse(tgt + v*rs, tgt + (v + 1)*rs);
  }
}
inline void match_faces(int nd, int rs, int n_var, const Dir& d, double* t) { perm_transpose(nd, rs, n_var, d, t); perm_flip(nd, rs, n_var, d, t); }
inline void restore_faces(int nd, int rs, int n_var, const Dir& d, double* t) { perm_flip(nd, rs, n_var, d, t); perm_transpose(nd, rs, n_var, d, t); }

/* ===================== kernels ===================== */

// reference include/Spatial.hpp:41-57
template <int ND, int RS, class Pde>
void write_face_elem(const Pde& eq, const ho_basis& b, const double* read, double* const* faces)
{
  constexpr int nq = ipow(RS, ND), nfq = nq/RS, ne = Pde::n_extrap;
  static thread_local std::vector<double> buf; buf.resize((size_t)ne*nq);
  double* extrap = buf.data();
  for (int q = 0; q < nq; ++q) {
    double g[ne]; eq.fetch_extrap(nq, read + q, g);
    for (int v = 0; v < ne; ++v) extrap[v*nq + q] = g[v];
  }
  for (int d = 0; d < ND; ++d) for (int fq = 0; fq < nfq; ++fq) for (int v = 0; v < ne; ++v) {
    for (int s = 0; s < 2; ++s) {
      double e = 0;
      for (int j = 0; j < RS; ++j) e += b.boundary[s][j]*extrap[v*nq + Rows<ND, RS>::qpoint(d, fq, j)];
      faces[2*d + s][v*nfq + fq] = e;
    }
  }
}

template <int ND, int RS, class Pde>
void write_face_all(const Pde& eq, const ho_basis& b, ho_mesh& m, int begin, int end)
{
  Geo<ND, RS> g(m);
  #pragma omp parallel for
  for (int e = begin; e < end; ++e) {
    double* faces[6];
    for (int f = 0; f < 2*ND; ++f) faces[f] = g.face(Pde::face_kind, e*2*ND + f);
    write_face_elem<ND, RS>(eq, b, g.state(e), faces);
  }
}

// reference include/Spatial.hpp:153-206 (prolong) and :230-285 (restrict); the two share the per-dimension sweep
template <int ND, int RS>
void transfer_sweep(double* var_face, const double (*mat)[8][8], const int* str, int i_face, double stretched_mult, bool divide)
{
  constexpr int nfq = ipow(RS, ND - 1);
  for (int d = 0; d < ND - 1; ++d) {
    if (str[d]) {
      for (int q = 0; q < nfq; ++q) { if (divide) var_face[q] /= stretched_mult; else var_face[q] *= stretched_mult; }
    } else {
      const int pw = ND - 2 - d;
      const int face_stride = str[ND - 2] ? 1 : ipow(2, pw);
      const int qstride = ipow(RS, pw);
      const int i_half = (i_face/face_stride)%2;
      for (int o = 0; o < nfq/(RS*qstride); ++o) for (int in = 0; in < qstride; ++in) {
        double row[RS], res[RS];
        for (int k = 0; k < RS; ++k) row[k] = var_face[(o*RS + k)*qstride + in];
        for (int i = 0; i < RS; ++i) { double s = 0; for (int k = 0; k < RS; ++k) s += mat[i_half][i][k]*row[k]; res[i] = s; }
        for (int k = 0; k < RS; ++k) var_face[(o*RS + k)*qstride + in] = res[k];
      }
    }
  }
}

// prolong needs "*= 1 + scl" for stretched dims, restrict "/= 1 + scl"; handled by dedicated drivers to keep transfer_sweep simple
template <int ND, int RS>
void prolong_refined(const ho_basis& b, ho_mesh& m, int n_var, int kind, bool scl)
{
  constexpr int nfq = ipow(RS, ND - 1), n_face_max = ipow(2, ND - 1);
  if constexpr (ND == 1) { (void)b; (void)m; (void)n_var; (void)kind; (void)scl; (void)nfq; (void)n_face_max; return; }
  else {
    Geo<ND, RS> g(m);
    #pragma omp parallel for
    for (int r = 0; r < m.n_ref; ++r) {
      const int* rf = m.ref_face + r*7;
      const int str[2] = {rf[5], rf[6]};
      int nf = n_face_max; for (int d = 0; d < ND - 1; ++d) nf /= 1 + str[d];
      const double* coarse = g.face(kind, rf[0]);
      for (int f = 0; f < nf; ++f) {
        double* fine = g.face(kind, rf[1 + f]);
        for (int v = 0; v < n_var; ++v) {
          double* vf = fine + v*nfq;
          for (int q = 0; q < nfq; ++q) vf[q] = coarse[v*nfq + q];
          transfer_sweep<ND, RS>(vf, b.prolong, str, f, 1 + scl, false);
        }
      }
    }
  }
}

template <int ND, int RS>
void restrict_refined(const ho_basis& b, ho_mesh& m, int n_var, int kind, bool scl)
{
  constexpr int nfq = ipow(RS, ND - 1), n_face_max = ipow(2, ND - 1);
  if constexpr (ND == 1) { (void)b; (void)m; (void)n_var; (void)kind; (void)scl; (void)nfq; (void)n_face_max; return; }
  else {
    Geo<ND, RS> g(m);
    #pragma omp parallel for
    for (int r = 0; r < m.n_ref; ++r) {
      const int* rf = m.ref_face + r*7;
      const int str[2] = {rf[5], rf[6]};
      int nf = n_face_max; for (int d = 0; d < ND - 1; ++d) nf /= 1 + str[d];
      double* coarse = g.face(kind, rf[0]);
      for (int i = 0; i < n_var*nfq; ++i) coarse[i] = 0.;
      for (int f = 0; f < nf; ++f) {
        double* fine = g.face(kind, rf[1 + f]); // note: the fine (mortar) data is transformed in place, as in the reference
        for (int v = 0; v < n_var; ++v) {
          double* vf = fine + v*nfq;
          transfer_sweep<ND, RS>(vf, b.restrict_, str, f, 1 + scl, true);
          for (int q = 0; q < nfq; ++q) coarse[v*nfq + q] += vf[q];
        }
      }
    }
  }
}

// reference include/Spatial.hpp:613-704
template <int ND, int RS, class Pde, bool DEF>
void neighbor(const Pde& eq, ho_mesh& m)
{
  constexpr int nfq = ipow(RS, ND - 1), ne = Pde::n_extrap, nu = Pde::n_update;
  Geo<ND, RS> g(m);
  const int n_con = DEF ? m.n_def_con : m.n_car_con;
  const int sk = Pde::face_kind;         // where the extrapolated state lives
  #pragma omp parallel for
  for (int c = 0; c < n_con; ++c) {
    int slot[2]; Dir dir;
    if constexpr (DEF) {
      const int* t = m.def_con + c*7;
      slot[0] = t[0]; slot[1] = t[1];
      dir = Dir{{t[2], t[3]}, {t[4], t[5]}};
    } else {
      const int* t = m.car_con + c*3;
      slot[0] = t[0]; slot[1] = t[1];
      dir = Dir{{t[2], t[2]}, {1, 0}}; // include/connection.hpp:37
    }
    double face[4][ne*nfq] = {};
    double face_nrml[ND*nfq];
    int sign[2] = {1, 1};
    for (int s = 0; s < 2; ++s) { const double* f = g.face(sk, slot[s]); for (int i = 0; i < ne*nfq; ++i) face[s][i] = f[i]; }
    if constexpr (DEF) {
      match_faces(ND, RS, ne, dir, face[1]);
      for (int s = 0; s < 2; ++s) sign[s] = 1 - 2*dir.flip_normal(s);
      const double* n = g.nrml(m.def_con[c*7 + 6]);
      for (int i = 0; i < ND*nfq; ++i) face_nrml[i] = n[i];
    }
    for (int q = 0; q < nfq; ++q) {
      if constexpr (Pde::has_diffusion) {
        for (int v = 0; v < ne; ++v) {
          const double avg = .5*(face[0][v*nfq + q] + face[1][v*nfq + q]);
          for (int s = 0; s < 2; ++s) face[2 + s][v*nfq + q] = avg;
        }
      }
      if constexpr (Pde::has_convection) {
        typename Pde::template Comp<1> comp[2] {typename Pde::template Comp<1>(eq), typename Pde::template Comp<1>(eq)};
        if constexpr (DEF) { for (int d = 0; d < ND; ++d) comp[0].normal[d][0] = sign[0]*face_nrml[d*nfq + q]; }
        else { for (int d = 0; d < ND; ++d) comp[0].normal[d][0] = (d == dir.i_dim[0]); }
        for (int d = 0; d < ND; ++d) comp[1].normal[d][0] = comp[0].normal[d][0];
        for (int s = 0; s < 2; ++s) comp[s].fetch_extrap_state(nfq, face[s] + q);
        for (int s = 0; s < 2; ++s) { comp[s].compute_flux_conv(); comp[s].compute_char_speed(); }
        double nsq = 0; for (int d = 0; d < ND; ++d) nsq += comp[0].normal[d][0]*comp[0].normal[d][0];
        const double nrm = std::sqrt(nsq);
        const double speed = std::max(comp[0].char_speed, comp[1].char_speed);
        for (int v = 0; v < nu; ++v) {
          const double flux = .5*(comp[0].flux_conv[v][0] + comp[1].flux_conv[v][0]
                                  + speed*nrm*(comp[0].update_state[v] - comp[1].update_state[v]));
          for (int s = 0; s < 2; ++s) face[s][v*nfq + q] = sign[s]*flux;
        }
      }
    }
    if constexpr (DEF) {
      restore_faces(ND, RS, ne, dir, face[1]);
      if constexpr (Pde::has_diffusion) restore_faces(ND, RS, ne, dir, face[3]);
    }
    for (int s = 0; s < 2; ++s) {
      if constexpr (Pde::has_convection) { double* f = g.face(sk, slot[s]); for (int i = 0; i < nu*nfq; ++i) f[i] = face[s][i]; }
      if constexpr (Pde::has_diffusion) { double* f = g.face(1, slot[s]); for (int i = 0; i < ne*nfq; ++i) f[i] = face[2 + s][i]; }
    }
  }
}

// reference include/Spatial.hpp:716-759
template <int ND, int RS, class Pde, bool DEF>
void neighbor_reconcile(ho_mesh& m)
{
  constexpr int nfq = ipow(RS, ND - 1), nu = Pde::n_update;
  Geo<ND, RS> g(m);
  const int n_con = DEF ? m.n_def_con : m.n_car_con;
  #pragma omp parallel for
  for (int c = 0; c < n_con; ++c) {
    int slot[2]; Dir dir;
    if constexpr (DEF) { const int* t = m.def_con + c*7; slot[0] = t[0]; slot[1] = t[1]; dir = Dir{{t[2], t[3]}, {t[4], t[5]}}; }
    else { const int* t = m.car_con + c*3; slot[0] = t[0]; slot[1] = t[1]; dir = Dir{{t[2], t[2]}, {1, 0}}; }
    double face[2][(ND + 2)*nfq];
    int sign[2] = {1, 1};
    for (int s = 0; s < 2; ++s) { const double* f = g.face(1, slot[s]); for (int i = 0; i < nu*nfq; ++i) face[s][i] = f[i]; }
    if constexpr (DEF) {
      // the reference permutes n_extrap variables of a buffer of which only n_update are filled; results for the filled part are identical
      match_faces(ND, RS, nu, dir, face[1]);
      for (int s = 0; s < 2; ++s) sign[s] = 1 - 2*dir.flip_normal(s);
    }
    for (int q = 0; q < nfq; ++q) for (int v = 0; v < nu; ++v) {
      double avg = 0;
      for (int s = 0; s < 2; ++s) avg += .5*sign[s]*face[s][v*nfq + q];
      for (int s = 0; s < 2; ++s) { double& f = face[s][v*nfq + q]; f = sign[s]*avg - f; }
    }
    if constexpr (DEF) restore_faces(ND, RS, nu, dir, face[1]);
    for (int s = 0; s < 2; ++s) { double* f = g.face(1, slot[s]); for (int i = 0; i < nu*nfq; ++i) f[i] = face[s][i]; }
  }
}

// reference include/Spatial.hpp:326-509
template <int ND, int RS, class Pde, bool DEF>
void local(const Pde& eq, const ho_basis& b, ho_mesh& m, ho_options o)
{
  constexpr int nq = ipow(RS, ND), nfq = nq/RS, ne = Pde::n_extrap, nu = Pde::n_update;
  using R = Rows<ND, RS>;
  Geo<ND, RS> g(m);
  const Deriv<RS> deriv(b);
  const bool stage = o.i_stage != 0;
  const double update = stage ? o.dt*(.5/b.quadratic_safety) : o.dt; // Spatial.hpp:317, Basis.cpp:11-14
  const int begin = DEF ? m.n_car : 0, end = DEF ? m.n_car + m.n_def : m.n_car;
  #pragma omp parallel
  {
    std::vector<double> time_rate((size_t)2*nu*nq), extrap((size_t)ne*nq), visc((size_t)ND*ne*nq), flux((size_t)ND*nu*nq);
    #pragma omp for
    for (int e = begin; e < end; ++e) {
      double* state = g.state(e);
      double* faces[6]; double* visc_faces[6];
      for (int f = 0; f < 2*ND; ++f) {
        faces[f] = g.face(Pde::face_kind, e*2*ND + f);
        visc_faces[f] = Pde::has_diffusion ? g.face(1, e*2*ND + f) : nullptr;
      }
      const double* tss = g.tss(e);
      const double d_pos = m.nom_size[e];
      const double* nrml = nullptr; const double* elem_det = nullptr; const double* face_nrml[6] = {};
      if constexpr (DEF) {
        elem_det = g.det(e); nrml = g.ref_nrml(e);
        for (int f = 0; f < 2*ND; ++f) face_nrml[f] = g.elem_face_nrml(e, f); // unit-normal fallback is materialised at mesh build time
      }
      std::fill(time_rate.begin(), time_rate.end(), 0.);
      std::fill(visc.begin(), visc.end(), 0.);

      if constexpr (Pde::has_diffusion) {
        // gradient (times jacobian determinant): Spatial.hpp:371-402
        for (int q = 0; q < nq; ++q) { double gv[ne]; eq.fetch_extrap(nq, state + q, gv); for (int v = 0; v < ne; ++v) extrap[v*nq + q] = gv[v]; }
        for (int d = 0; d < ND; ++d) for (int fq = 0; fq < nfq; ++fq) {
          for (int v = 0; v < ne; ++v) {
            double row[RS], bv[2];
            for (int k = 0; k < RS; ++k) row[k] = extrap[v*nq + R::qpoint(d, fq, k)];
            for (int s = 0; s < 2; ++s) bv[s] = visc_faces[2*d + s][v*nfq + fq];
            if constexpr (DEF) {
              for (int j = 0; j < ND; ++j) {
                double rn[RS], bn[2], out[RS];
                for (int k = 0; k < RS; ++k) rn[k] = nrml[(d*ND + j)*nq + R::qpoint(d, fq, k)]*row[k];
                for (int s = 0; s < 2; ++s) bn[s] = face_nrml[2*d + s][j*nfq + fq]*bv[s];
                deriv.full(rn, bn, out);
                for (int k = 0; k < RS; ++k) { double& t = visc[(j*ne + v)*nq + R::qpoint(d, fq, k)]; t = 1.*t + out[k]/d_pos; }
              }
            } else {
              double out[RS];
              deriv.full(row, bv, out);
              for (int k = 0; k < RS; ++k) visc[(d*ne + v)*nq + R::qpoint(d, fq, k)] = out[k]/d_pos;
            }
          }
        }
      }

      // pointwise flux: Spatial.hpp:405-442
      for (int q = 0; q < nq; ++q) {
        typename Pde::template Comp<ND> comp(eq);
        comp.fetch_state(nq, state + q);
        if constexpr (DEF) for (int d = 0; d < ND; ++d) for (int j = 0; j < ND; ++j) comp.normal[j][d] = nrml[(d*ND + j)*nq + q];
        if constexpr (Pde::has_convection) {
          comp.compute_flux_conv();
          for (int d = 0; d < ND; ++d) for (int v = 0; v < nu; ++v) flux[(d*nu + v)*nq + q] = comp.flux_conv[v][d];
        }
        if constexpr (Pde::has_diffusion) {
          for (int d = 0; d < ND; ++d) for (int v = 0; v < ne; ++v) comp.gradient[v][d] = visc[(d*ne + v)*nq + q];
          if constexpr (DEF) for (int d = 0; d < ND; ++d) for (int v = 0; v < ne; ++v) comp.gradient[v][d] /= elem_det[q];
          comp.compute_flux_diff();
          for (int d = 0; d < ND; ++d) for (int v = 0; v < nu; ++v) visc[(d*ne + v)*nq + q] = comp.flux_diff[v][d];
        }
        if constexpr (Pde::has_source) if (!stage) {
          comp.compute_source();
          double mult = d_pos;
          if constexpr (DEF) mult *= elem_det[q];
          for (int v = 0; v < nu; ++v) time_rate[(nu + v)*nq + q] = mult*comp.source[v];
        }
      }

      // residual: Spatial.hpp:445-470
      for (int d = 0; d < ND; ++d) for (int fq = 0; fq < nfq; ++fq) {
        if constexpr (Pde::has_convection) {
          for (int v = 0; v < nu; ++v) {
            double row[RS], bv[2], out[RS];
            for (int k = 0; k < RS; ++k) row[k] = flux[(d*nu + v)*nq + R::qpoint(d, fq, k)];
            for (int s = 0; s < 2; ++s) bv[s] = faces[2*d + s][v*nfq + fq];
            deriv.full(row, bv, out);
            for (int k = 0; k < RS; ++k) { double& t = time_rate[v*nq + R::qpoint(d, fq, k)]; t = 1.*t + -out[k]; }
          }
        }
        if constexpr (Pde::has_diffusion) {
          for (int v = 0; v < nu; ++v) {
            double row[RS], out[RS];
            for (int k = 0; k < RS; ++k) row[k] = visc[(d*ne + v)*nq + R::qpoint(d, fq, k)];
            for (int s = 0; s < 2; ++s) {
              double ext = 0; for (int k = 0; k < RS; ++k) ext += b.boundary[s][k]*row[k];
              visc_faces[2*d + s][v*nfq + fq] = ext;
            }
            deriv.interior(row, out);
            for (int k = 0; k < RS; ++k) { double& t = time_rate[(nu + v)*nq + R::qpoint(d, fq, k)]; t = 1.*t + -out[k]; }
          }
        }
      }

      // modal filter: Spatial.hpp:473-481
      if (o.use_filter) {
        for (int d = 0; d < ND; ++d) for (int fq = 0; fq < nfq; ++fq) for (int v = 0; v < 2*nu; ++v) {
          double row[RS], out[RS];
          for (int k = 0; k < RS; ++k) row[k] = time_rate[v*nq + R::qpoint(d, fq, k)];
          for (int i = 0; i < RS; ++i) { double s = 0; for (int k = 0; k < RS; ++k) s += b.filter[i][k]*row[k]; out[i] = s; }
          for (int k = 0; k < RS; ++k) time_rate[v*nq + R::qpoint(d, fq, k)] = out[k];
        }
      }

      // update: Spatial.hpp:484-503
      double* ref_state = g.cache(e);
      for (int q = 0; q < nq; ++q) {
        double upd[nu]; for (int v = 0; v < nu; ++v) upd[v] = 0;
        double mult = update*tss[q]/d_pos;
        if constexpr (DEF) mult /= elem_det[q];
        for (int v = 0; v < nu; ++v) {
          double u = time_rate[v*nq + q];
          if (stage) u -= ref_state[v*nq + q];
          else {
            if constexpr (Pde::has_convection) ref_state[v*nq + q] = u;
            if constexpr (Pde::has_diffusion || Pde::has_source) u += time_rate[(nu + v)*nq + q];
          }
          u *= mult;
          if (o.compute_residual) ref_state[v*nq + q] = u;
          else upd[v] = u;
        }
        eq.write_update(upd, nq, state + q, !Pde::has_diffusion && !stage);
      }
      if constexpr (!Pde::has_diffusion) write_face_elem<ND, RS>(eq, b, state, faces);
    }
  }
}

// reference include/Spatial.hpp:543-594
template <int ND, int RS, class Pde, bool DEF>
void reconcile_ldg_flux(const Pde& eq, const ho_basis& b, ho_mesh& m, ho_options o)
{
  constexpr int nq = ipow(RS, ND), nfq = nq/RS, nu = Pde::n_update;
  using R = Rows<ND, RS>;
  Geo<ND, RS> g(m);
  const Deriv<RS> deriv(b);
  const int begin = DEF ? m.n_car : 0, end = DEF ? m.n_car + m.n_def : m.n_car;
  #pragma omp parallel
  {
    std::vector<double> time_rate((size_t)nu*nq);
    #pragma omp for
    for (int e = begin; e < end; ++e) {
      double* state = g.state(e);
      const double* tss = g.tss(e);
      const double d_pos = m.nom_size[e];
      const double* elem_det = DEF ? g.det(e) : nullptr;
      std::fill(time_rate.begin(), time_rate.end(), 0.);
      for (int d = 0; d < ND; ++d) for (int fq = 0; fq < nfq; ++fq) for (int v = 0; v < nu; ++v) {
        double bv[2], out[RS];
        for (int s = 0; s < 2; ++s) bv[s] = g.face(1, e*2*ND + 2*d + s)[v*nfq + fq];
        deriv.boundary_term(bv, out);
        for (int k = 0; k < RS; ++k) { double& t = time_rate[v*nq + R::qpoint(d, fq, k)]; t = 1.*t + -out[k]; }
      }
      if (o.use_filter) {
        for (int d = 0; d < ND; ++d) for (int fq = 0; fq < nfq; ++fq) for (int v = 0; v < nu; ++v) {
          double row[RS], out[RS];
          for (int k = 0; k < RS; ++k) row[k] = time_rate[v*nq + R::qpoint(d, fq, k)];
          for (int i = 0; i < RS; ++i) { double s = 0; for (int k = 0; k < RS; ++k) s += b.filter[i][k]*row[k]; out[i] = s; }
          for (int k = 0; k < RS; ++k) time_rate[v*nq + R::qpoint(d, fq, k)] = out[k];
        }
      }
      // the update lands in the state, or in the residual cache when only the residual is wanted.
      // `write_update` is handed the chosen base pointer, exactly as the reference does (Spatial.hpp:579,587)
      double* to_update = o.compute_residual ? g.cache(e) : state;
      for (int q = 0; q < nq; ++q) {
        double upd[nu];
        double mult = o.dt*tss[q]/d_pos;
        if constexpr (DEF) mult /= elem_det[q];
        for (int v = 0; v < nu; ++v) upd[v] = time_rate[v*nq + q]*mult;
        eq.write_update(upd, nq, to_update + q, true);
      }
      double* faces[6];
      for (int f = 0; f < 2*ND; ++f) faces[f] = g.face(Pde::face_kind, e*2*ND + f);
      write_face_elem<ND, RS>(eq, b, state, faces);
    }
  }
}

// reference include/Spatial.hpp:784-828 and include/math.hpp:207-218
template <int ND, int RS, class Pde, bool DEF>
double max_dt(const Pde& eq, const ho_basis& b, ho_mesh& m, bool is_local, double safety_conv, double safety_diff)
{
  constexpr int nq = ipow(RS, ND), n_vert = ipow(2, ND);
  Geo<ND, RS> g(m);
  const double max_cfl_c = (-2*b.quadratic_safety/b.min_eig_convection)*safety_conv; // Basis.cpp:6-9
  const double max_cfl_d = -2/b.min_eig_diffusion*safety_diff;
  const int begin = DEF ? m.n_car : 0, end = DEF ? m.n_car + m.n_def : m.n_car;
  double dt = std::numeric_limits<double>::max();
  #pragma omp parallel for reduction(min:dt)
  for (int e = begin; e < end; ++e) {
    double* state = g.state(e);
    double* tss = g.tss(e);
    for (int q = 0; q < nq; ++q) {
      double vals[n_vert];
      for (int i = 0; i < n_vert; ++i) vals[i] = m.vertex_tss[(size_t)e*n_vert + i];
      int stride = n_vert;
      for (int d = 0; d < ND; ++d) {
        const double coord = b.node[(q/ipow(RS, ND - 1 - d))%RS];
        stride /= 2;
        for (int i = 0; i < stride; ++i) vals[i] += coord*(vals[i + stride] - vals[i]);
      }
      const double spacing = vals[0];
      typename Pde::template Comp<ND> comp(eq);
      comp.fetch_state(nq, state + q);
      double scale = 0;
      if constexpr (Pde::has_convection) { comp.compute_char_speed(); scale += comp.char_speed/max_cfl_c/spacing; }
      if constexpr (Pde::has_diffusion) { comp.compute_diffusivity(); scale += comp.diffusivity/max_cfl_d/spacing/spacing; }
      if (is_local) tss[q] = 1./scale;
      else { tss[q] = 1.; dt = std::min(dt, 1./scale); }
    }
  }
  return is_local ? 1. : dt;
}

// reference src/stabilizing_art_visc.cpp:8-61
template <int ND, int RS>
void stab_art_visc(const ho_basis& b, ho_mesh& m, double char_speed)
{
  constexpr int nq = ipow(RS, ND), nfq = nq/RS;
  using R = Rows<ND, RS>;
  Geo<ND, RS> g(m);
  const double ramp_center = -4.25*std::log(RS - 1)/std::log(10);
  const double half_width = 0.5;
  double qw[nq], fw[nfq], proj[RS];
  for (int q = 0; q < nq; ++q) { double w = 1; for (int d = 0; d < ND; ++d) w *= b.weight[(q/ipow(RS, ND - 1 - d))%RS]; qw[q] = w; }
  for (int q = 0; q < nfq; ++q) { double w = 1; for (int d = 0; d < ND - 1; ++d) w *= b.weight[(q/ipow(RS, ND - 2 - d))%RS]; fw[q] = w; }
  for (int i = 0; i < RS; ++i) proj[i] = b.orthogonal[RS - 1][i]*b.weight[i];
  #pragma omp parallel for
  for (int e = 0; e < m.n_car + m.n_def; ++e) {
    const double* state = g.state(e);
    double ind[nq]; double norm_sq = 0;
    for (int q = 0; q < nq; ++q) { ind[q] = 1./state[ND*nq + q]; norm_sq += ind[q]*ind[q]*qw[q]; }
    double nonsmooth = 0;
    for (int d = 0; d < ND; ++d) for (int fq = 0; fq < nfq; ++fq) {
      double dot = 0; for (int k = 0; k < RS; ++k) dot += ind[R::qpoint(d, fq, k)]*proj[k];
      nonsmooth += dot*dot*fw[fq];
    }
    nonsmooth /= norm_sq*ND;
    double indicator = std::log(nonsmooth)/std::log(10);
    if (indicator <= ramp_center - half_width) indicator = 0;
    else if (indicator < ramp_center + half_width) indicator = .5*(1 + std::sin(M_PI*(indicator - ramp_center)/2/half_width));
    else indicator = 1;
    m.uncert[e] = (RS - 1)*char_speed*m.nom_size[e]*indicator;
  }
}


} // namespace ho_impl
#endif
