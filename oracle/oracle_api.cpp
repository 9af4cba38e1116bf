/* oracle_api.cpp -- extern "C" front of the CPU oracle (declared in flat_mesh.h).
 * TEST INFRASTRUCTURE (see oracle_impl.hpp): PDE-independent kernels live here, PDE-dependent
 * ones are routed to the ho_part_<nd>_<pde> translation units. */
#include "oracle_impl.hpp"
#include "oracle_call.h"
#include "characteristics.hpp"
#include <type_traits>

namespace {
using namespace ho_impl;

const ho_part_fn parts[3][5] = {
  {ho_part_1_0, ho_part_1_1, ho_part_1_2, ho_part_1_3, ho_part_1_4},
  {ho_part_2_0, ho_part_2_1, ho_part_2_2, ho_part_2_3, ho_part_2_4},
  {ho_part_3_0, ho_part_3_1, ho_part_3_2, ho_part_3_3, ho_part_3_4},
};

int route(int pde, ho_call& c)
{
  const int nd = c.mesh->n_dim;
  if (nd < 1 || nd > 3 || pde < 0 || pde > 4) return 1;
  return parts[nd - 1][pde](&c);
}

template <class F>
int dispatch_dims(int nd, int rs, F&& f)
{
  #define HO_CASE(ND, RS) if (nd == ND && rs == RS) { f(std::integral_constant<int, ND>{}, std::integral_constant<int, RS>{}); return 0; }
  #define HO_ROW(ND) HO_CASE(ND, 2) HO_CASE(ND, 3) HO_CASE(ND, 4) HO_CASE(ND, 5) HO_CASE(ND, 6) HO_CASE(ND, 7) HO_CASE(ND, 8)
  HO_ROW(1) HO_ROW(2) HO_ROW(3)
  #undef HO_ROW
  #undef HO_CASE
  return 1;
}

int n_extrap_of(int pde, int nd, int rs)
{ return pde == HO_ADVECTION ? nd + rs : pde == HO_SMOOTH_AV ? 3 : nd + 2; }
} // namespace

extern "C" {

int ho_num_threads(void)
{
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

int ho_compute_euler(const ho_basis* b, ho_mesh* m, ho_options o)
{ ho_call c{}; c.op = HO_OP_CONV_STAGE; c.basis = b; c.mesh = m; c.opts = o; return route(HO_EULER, c); }

int ho_compute_advection(const ho_basis* b, ho_mesh* m, ho_options o, double advect_length)
{ ho_call c{}; c.op = HO_OP_CONV_STAGE; c.basis = b; c.mesh = m; c.opts = o; c.p0 = advect_length; return route(HO_ADVECTION, c); }

int ho_compute_navier_stokes(const ho_basis* b, ho_mesh* m, ho_options o, ho_callback cb, void* user, ho_transport visc, ho_transport cond)
{
  ho_call c{}; c.op = HO_OP_DIFF_STAGE; c.basis = b; c.mesh = m; c.opts = o; c.flux_bc = cb; c.user = user; c.visc = visc; c.cond = cond;
  return route(HO_NAVIER_STOKES, c);
}

int ho_compute_smooth_av(const ho_basis* b, ho_mesh* m, ho_options o, ho_callback cb, void* user, double diff_time, double cheby_step)
{
  ho_call c{}; c.op = HO_OP_DIFF_STAGE; c.basis = b; c.mesh = m; c.opts = o; c.flux_bc = cb; c.user = user; c.p0 = diff_time; c.p1 = cheby_step;
  return route(HO_SMOOTH_AV, c);
}

int ho_compute_fix_therm_admis(const ho_basis* b, ho_mesh* m, ho_options o, ho_callback cb, void* user)
{ ho_call c{}; c.op = HO_OP_DIFF_STAGE; c.basis = b; c.mesh = m; c.opts = o; c.flux_bc = cb; c.user = user; return route(HO_FIX_THERM_ADMIS, c); }

int ho_max_dt(int pde, const ho_basis* b, ho_mesh* m, double sc, double sd, int local_time,
              ho_transport visc, ho_transport cond, double advect_length, double* dt_out)
{
  // reference src/kernels_max_dt.cpp:14-21; smooth_av is constructed with (1., 1.)
  ho_call c{}; c.op = HO_OP_MAX_DT; c.basis = b; c.mesh = m; c.visc = visc; c.cond = cond;
  c.p0 = pde == HO_SMOOTH_AV ? 1. : advect_length; c.p1 = 1.;
  c.safety_conv = sc; c.safety_diff = sd; c.local_time = local_time; c.dt_out = dt_out;
  return route(pde, c);
}

int ho_compute_write_face(int pde, const ho_basis* b, ho_mesh* m)
{
  // reference src/kernels_convective.cpp:43-56: advection is built with (1.), smooth_av with (1., 1.)
  ho_call c{}; c.op = HO_OP_WRITE_FACE; c.basis = b; c.mesh = m; c.p0 = 1.; c.p1 = 1.;
  return route(pde, c);
}

int ho_compute_prolong(int pde, const ho_basis* b, ho_mesh* m, int scale, int offset)
{
  const int kind = pde == HO_ADVECTION ? 2 : offset ? 1 : 0;
  return dispatch_dims(m->n_dim, m->row_size, [&](auto nd, auto rs) {
    prolong_refined<decltype(nd)::value, decltype(rs)::value>(*b, *m, n_extrap_of(pde, m->n_dim, m->row_size), kind, scale);
  });
}

int ho_compute_restrict(int pde, const ho_basis* b, ho_mesh* m, int scale, int offset)
{
  const int kind = pde == HO_ADVECTION ? 2 : offset ? 1 : 0;
  return dispatch_dims(m->n_dim, m->row_size, [&](auto nd, auto rs) {
    restrict_refined<decltype(nd)::value, decltype(rs)::value>(*b, *m, n_extrap_of(pde, m->n_dim, m->row_size), kind, scale);
  });
}

int ho_face_permutation(int n_dim, int row_size, int n_var, const int dir[4], int restore, double* data)
{
  Dir d{{dir[0], dir[1]}, {dir[2], dir[3]}};
  if (restore) restore_faces(n_dim, row_size, n_var, d, data);
  else match_faces(n_dim, row_size, n_var, d, data);
  return 0;
}

int ho_stabilizing_art_visc(const ho_basis* b, ho_mesh* m, double char_speed)
{
  return dispatch_dims(m->n_dim, m->row_size, [&](auto nd, auto rs) {
    stab_art_visc<decltype(nd)::value, decltype(rs)::value>(*b, *m, char_speed);
  });
}

int ho_neighbor(int pde, int deformed, ho_mesh* m, int i_stage, ho_transport visc, ho_transport cond, double p0, double p1)
{
  ho_basis dummy{}; dummy.row_size = m->row_size;
  ho_call c{}; c.op = HO_OP_NEIGHBOR; c.basis = &dummy; c.mesh = m; c.opts.i_stage = i_stage; c.deformed = deformed;
  c.visc = visc; c.cond = cond; c.p0 = p0; c.p1 = p1;
  return route(pde, c);
}

int ho_local(int pde, int deformed, const ho_basis* b, ho_mesh* m, ho_options o, ho_transport visc, ho_transport cond, double p0, double p1)
{
  ho_call c{}; c.op = HO_OP_LOCAL; c.basis = b; c.mesh = m; c.opts = o; c.deformed = deformed;
  c.visc = visc; c.cond = cond; c.p0 = p0; c.p1 = p1;
  return route(pde, c);
}

int ho_neighbor_reconcile(int pde, int deformed, ho_mesh* m)
{
  ho_basis dummy{}; dummy.row_size = m->row_size;
  ho_call c{}; c.op = HO_OP_NEIGHBOR_RECONCILE; c.basis = &dummy; c.mesh = m; c.deformed = deformed;
  return route(pde, c);
}

int ho_reconcile_ldg_flux(int pde, int deformed, const ho_basis* b, ho_mesh* m, ho_options o, ho_transport visc, ho_transport cond, double p0, double p1)
{
  ho_call c{}; c.op = HO_OP_RECONCILE_LDG; c.basis = b; c.mesh = m; c.opts = o; c.deformed = deformed;
  c.visc = visc; c.cond = cond; c.p0 = p0; c.p1 = p1;
  return route(pde, c);
}

int ho_derivative(const ho_basis* b, int n_var, const double* q, const double* bv, double* result)
{
  // q: [n_var][row_size], bv: [n_var][2], result: [n_var][row_size]
  return dispatch_dims(1, b->row_size, [&](auto, auto rs) {
    constexpr int RS = decltype(rs)::value;
    Deriv<RS> d(*b);
    for (int v = 0; v < n_var; ++v) d.full(q + v*RS, bv + v*2, result + v*RS);
  });
}

int ho_bc_freestream(ho_mesh* m, int n_bc, const int* ghost_slot, const double* fs)
{
  // reference src/Boundary_condition.cpp:66-76
  const int nfq = ipow(m->row_size, m->n_dim - 1), nv = m->n_dim + 2;
  for (int i = 0; i < n_bc; ++i) {
    double* gf = m->face_state + (size_t)ghost_slot[i]*nv*nfq;
    for (int v = 0; v < nv; ++v) for (int q = 0; q < nfq; ++q) gf[v*nfq + q] = fs[v];
  }
  return 0;
}

int ho_bc_copy(ho_mesh* m, int n_bc, const int* inside_slot, const int* ghost_slot)
{
  // copy_state touches both halves of the face storage (src/Boundary_condition.cpp:12-23)
  const int w = (m->n_dim + 2)*ipow(m->row_size, m->n_dim - 1);
  for (int i = 0; i < n_bc; ++i) {
    std::memcpy(m->face_state + (size_t)ghost_slot[i]*w, m->face_state + (size_t)inside_slot[i]*w, sizeof(double)*w);
    if (m->face_ldg) std::memcpy(m->face_ldg + (size_t)ghost_slot[i]*w, m->face_ldg + (size_t)inside_slot[i]*w, sizeof(double)*w);
  }
  return 0;
}

int ho_bc_nonpenetration(ho_mesh* m, int n_bc, const int* inside_slot, const int* ghost_slot, const int* normal_slot)
{
  // reference src/Boundary_condition.cpp:301-327
  const int nd = m->n_dim, nfq = ipow(m->row_size, nd - 1), w = (nd + 2)*nfq;
  for (int i = 0; i < n_bc; ++i) {
    double* gh = m->face_state + (size_t)ghost_slot[i]*w;
    const double* in = m->face_state + (size_t)inside_slot[i]*w;
    const double* n = m->normals + (size_t)normal_slot[i]*nd*nfq;
    for (int k = 0; k < w; ++k) gh[k] = in[k];
    for (int q = 0; q < nfq; ++q) {
      double dot = 0., nsq = 0.;
      for (int d = 0; d < nd; ++d) { dot += gh[d*nfq + q]*n[d*nfq + q]; nsq += n[d*nfq + q]*n[d*nfq + q]; }
      for (int d = 0; d < nd; ++d) gh[d*nfq + q] -= 2*dot*n[d*nfq + q]/nsq;
    }
  }
  return 0;
}

/* Characteristics (reference include/pde.hpp:181-256): eigvals[3], decomp[(nd+2)][3] of `state1` about `state` along `direction` */
int ho_characteristics(int nd, const double* state, const double* direction, const double* state1, double* eigvals, double* decomp)
{
  auto run = [&](auto ndc) {
    constexpr int ND = decltype(ndc)::value;
    Characteristics<ND> ch(state, direction);
    double d[ND + 2][3];
    ch.decomp(state1, d);
    for (int j = 0; j < 3; ++j) eigvals[j] = ch.vals[j];
    for (int v = 0; v < ND + 2; ++v) for (int j = 0; j < 3; ++j) decomp[v*3 + j] = d[v][j];
  };
  if (nd == 1) run(std::integral_constant<int, 1>{});
  else if (nd == 2) run(std::integral_constant<int, 2>{});
  else if (nd == 3) run(std::integral_constant<int, 3>{});
  else return 1;
  return 0;
}

/* Riemann_invariants::apply_state (reference src/Boundary_condition.cpp:97-139); `cache` [n_bc][nv*nfq] receives the inside state */
int ho_bc_riemann_state(ho_mesh* m, int n_bc, const int* inside_slot, const int* ghost_slot, const int* normal_slot, const double* fs, double* cache)
{
  const int nd = m->n_dim, nfq = ipow(m->row_size, nd - 1), nv = nd + 2, w = nv*nfq;
  if (nd < 1 || nd > 3) return 1;
  for (int i = 0; i < n_bc; ++i) {
    double* gh = m->face_state + (size_t)ghost_slot[i]*w;
    const double* in = m->face_state + (size_t)inside_slot[i]*w;
    const double* n = m->normals + (size_t)normal_slot[i]*nd*nfq;
    const int sign = 1 - 2*((inside_slot[i] % (2*nd)) % 2);
    for (int q = 0; q < nfq; ++q) {
      double inside[5], nrml[3], ghost[5];
      for (int v = 0; v < nv; ++v) inside[v] = in[v*nfq + q];
      for (int d = 0; d < nd; ++d) nrml[d] = n[d*nfq + q];
      if (nd == 1) riemann_state_point<1>(inside, nrml, sign, fs, ghost);
      else if (nd == 2) riemann_state_point<2>(inside, nrml, sign, fs, ghost);
      else riemann_state_point<3>(inside, nrml, sign, fs, ghost);
      for (int v = 0; v < nv; ++v) gh[v*nfq + q] = ghost[v];
    }
    for (int k = 0; k < w; ++k) cache[(size_t)i*w + k] = in[k];
  }
  return 0;
}

/* Riemann_invariants::apply_flux (reference src/Boundary_condition.cpp:141-182) on the LDG halves of the faces */
int ho_bc_riemann_flux(ho_mesh* m, int n_bc, const int* inside_slot, const int* ghost_slot, const int* normal_slot, const double* cache)
{
  const int nd = m->n_dim, nfq = ipow(m->row_size, nd - 1), nv = nd + 2, w = nv*nfq;
  if (nd < 1 || nd > 3 || !m->face_ldg) return 1;
  for (int i = 0; i < n_bc; ++i) {
    double* gh = m->face_ldg + (size_t)ghost_slot[i]*w;
    const double* in = m->face_ldg + (size_t)inside_slot[i]*w;
    const double* n = m->normals + (size_t)normal_slot[i]*nd*nfq;
    const int sign = 1 - 2*((inside_slot[i] % (2*nd)) % 2);
    for (int q = 0; q < nfq; ++q) {
      double flux[5], sc[5], nrml[3], ghost[5];
      for (int v = 0; v < nv; ++v) { flux[v] = in[v*nfq + q]; sc[v] = cache[(size_t)i*w + v*nfq + q]; }
      for (int d = 0; d < nd; ++d) nrml[d] = n[d*nfq + q];
      if (nd == 1) riemann_flux_point<1>(sc, nrml, sign, flux, ghost);
      else if (nd == 2) riemann_flux_point<2>(sc, nrml, sign, flux, ghost);
      else riemann_flux_point<3>(sc, nrml, sign, flux, ghost);
      for (int v = 0; v < nv; ++v) gh[v*nfq + q] = ghost[v];
    }
  }
  return 0;
}

} // extern "C"
