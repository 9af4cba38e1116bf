/* oracle_part.cpp -- one (n_dim, PDE) slice of the CPU oracle; compile with -DHO_ND=<1|2|3> -DHO_PDE=<0..4>.
 * TEST INFRASTRUCTURE (see oracle_impl.hpp). Row sizes 2..8 are instantiated, mirroring the reference's
 * run-time -> compile-time lookup (include/kernel_factory.hpp:64-119, config::max_row_size = 8);
 * define HO_ONLY_RS=<n> to build a single row size (used for the quick -march=native rebuild on the GPU box). */
#include "oracle_impl.hpp"
#include "oracle_call.h"

#ifndef HO_ND
#error "define HO_ND"
#endif
#ifndef HO_PDE
#error "define HO_PDE"
#endif

namespace {
using namespace ho_impl;

template <int ND, int RS>
auto make_pde(const ho_call& c)
{
  if constexpr (HO_PDE == HO_EULER) return Pde_ns<ND, RS, false>{c.visc, c.cond};
  else if constexpr (HO_PDE == HO_NAVIER_STOKES) return Pde_ns<ND, RS, true>{c.visc, c.cond};
  else if constexpr (HO_PDE == HO_ADVECTION) {
    // pde.hpp:281: nodes of Gauss_legendre(row_size) mapped to [-1, 1]; the oracle is only driven with Gauss-Legendre bases
    Pde_advection<ND, RS> p; p.advect_length = c.p0;
    for (int i = 0; i < RS; ++i) p.nodes[i] = 2*c.basis->node[i] - 1;
    return p;
  }
  else if constexpr (HO_PDE == HO_SMOOTH_AV) return Pde_smooth_av<ND, RS>{c.p0, c.p1};
  else return Pde_fta<ND, RS>{};
}

template <int RS>
int run(ho_call& c)
{
  constexpr int ND = HO_ND;
  auto eq = make_pde<ND, RS>(c);
  using P = decltype(eq);
  const ho_basis& b = *c.basis;
  ho_mesh& m = *c.mesh;
  switch (c.op) {
    case HO_OP_CONV_STAGE: // reference src/kernels_convective.cpp:8-16
      if constexpr (P::has_convection && !P::has_diffusion) {
        neighbor<ND, RS, P, false>(eq, m);
        neighbor<ND, RS, P, true>(eq, m);
        restrict_refined<ND, RS>(b, m, P::n_extrap, P::face_kind, true);
        local<ND, RS, P, false>(eq, b, m, c.opts);
        local<ND, RS, P, true>(eq, b, m, c.opts);
        prolong_refined<ND, RS>(b, m, P::n_extrap, P::face_kind, false);
        return 0;
      } else return 3;
    case HO_OP_DIFF_STAGE: // reference src/kernels_diffusive.cpp:8-26
      if constexpr (P::has_diffusion) {
        neighbor<ND, RS, P, false>(eq, m);
        neighbor<ND, RS, P, true>(eq, m);
        restrict_refined<ND, RS>(b, m, P::n_extrap, 0, true);
        restrict_refined<ND, RS>(b, m, P::n_extrap, 1, false);
        local<ND, RS, P, false>(eq, b, m, c.opts);
        local<ND, RS, P, true>(eq, b, m, c.opts);
        if (!c.opts.i_stage) {
          prolong_refined<ND, RS>(b, m, P::n_extrap, 1, true);
          if (c.flux_bc) c.flux_bc(c.user);
          neighbor_reconcile<ND, RS, P, false>(m);
          neighbor_reconcile<ND, RS, P, true>(m);
          restrict_refined<ND, RS>(b, m, P::n_extrap, 1, true);
          reconcile_ldg_flux<ND, RS, P, false>(eq, b, m, c.opts);
          reconcile_ldg_flux<ND, RS, P, true>(eq, b, m, c.opts);
        }
        prolong_refined<ND, RS>(b, m, P::n_extrap, 0, false);
        return 0;
      } else return 3;
    case HO_OP_MAX_DT: { // reference src/kernels_max_dt.cpp:8-12
      const double car = max_dt<ND, RS, P, false>(eq, b, m, c.local_time, c.safety_conv, c.safety_diff);
      const double def = max_dt<ND, RS, P, true>(eq, b, m, c.local_time, c.safety_conv, c.safety_diff);
      *c.dt_out = std::min(car, def);
      return 0;
    }
    case HO_OP_WRITE_FACE:
      write_face_all<ND, RS>(eq, b, m, 0, m.n_car + m.n_def);
      return 0;
    case HO_OP_NEIGHBOR:
      if (c.deformed) neighbor<ND, RS, P, true>(eq, m); else neighbor<ND, RS, P, false>(eq, m);
      return 0;
    case HO_OP_LOCAL:
      if (c.deformed) local<ND, RS, P, true>(eq, b, m, c.opts); else local<ND, RS, P, false>(eq, b, m, c.opts);
      return 0;
    case HO_OP_NEIGHBOR_RECONCILE:
      if constexpr (P::has_diffusion) {
        if (c.deformed) neighbor_reconcile<ND, RS, P, true>(m); else neighbor_reconcile<ND, RS, P, false>(m);
        return 0;
      } else return 3;
    case HO_OP_RECONCILE_LDG:
      if constexpr (P::has_diffusion) {
        if (c.deformed) reconcile_ldg_flux<ND, RS, P, true>(eq, b, m, c.opts); else reconcile_ldg_flux<ND, RS, P, false>(eq, b, m, c.opts);
        return 0;
      } else return 3;
  }
  return 4;
}
} // namespace

#define HO_CAT_(a, b, c) ho_part_##a##_##b
#define HO_CAT(a, b) HO_CAT_(a, b, 0)
extern "C" int HO_CAT(HO_ND, HO_PDE)(ho_call* c)
{
  switch (c->mesh->row_size) {
#ifdef HO_ONLY_RS
    case HO_ONLY_RS: return run<HO_ONLY_RS>(*c);
#else
    case 2: return run<2>(*c);
    case 3: return run<3>(*c);
    case 4: return run<4>(*c);
    case 5: return run<5>(*c);
    case 6: return run<6>(*c);
    case 7: return run<7>(*c);
    case 8: return run<8>(*c);
#endif
  }
  return 1; // "demand for invalid kernel" (kernel_factory.hpp:114-116)
}
