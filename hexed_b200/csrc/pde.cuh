/* pde.cuh -- device-side PDE policies for the generic kernels (generic_part.cu).
 *
 * One policy per PDE of the reference (include/pde.hpp): which element slots feed the state / extrapolated variables,
 * how an update is written back, and the pointwise physics ("Comp", the analogue of `Pde::Computation<n_dim_flux>`).
 * The element data of the reference is one slot-major array per element (src/Element.cpp:114-142); on the device the slot
 * groups live in separate arrays (state, tss, av, forcing, adv, cache), so a policy names reference SLOT indices and
 * `ElemData::slot` resolves them.
 */
#ifndef HB_PDE_CUH_
#define HB_PDE_CUH_
#include "common.cuh"

namespace hb {

constexpr double heat_rat_ns = 1.4;               // include/pde.hpp:51
constexpr double specific_gas_air = 287.05287;    // include/constants.hpp:49

enum { PDE_EULER = 0, PDE_NAVIER_STOKES = 1, PDE_ADVECTION = 2, PDE_SMOOTH_AV = 3, PDE_FIX_THERM_ADMIS = 4 };

/* device view of the per-element arrays */
struct ElemData
{
  double *state, *tss, *av, *forcing, *adv, *cache;
  template <int ND, int RS>
  __device__ __forceinline__ double* slot(int e, int s) const
  {
    constexpr int nv = ND + 2, nq = ipow(RS, ND), cs = nv > RS ? nv : RS;
    if (s < nv) return state + ((size_t)e*nv + s)*nq;
    if (s == nv) return tss + (size_t)e*nq;
    if (s < nv + 3) return av + ((size_t)e*2 + (s - nv - 1))*nq;
    if (s < nv + 7) return forcing + ((size_t)e*4 + (s - nv - 3))*nq;
    if (s < nv + 7 + RS) return adv + ((size_t)e*RS + (s - nv - 7))*nq;
    return cache + ((size_t)e*cs + (s - nv - 7 - RS))*nq;
  }
};

__device__ __forceinline__ double transport_coef(const hexed_b200_transport& t, double inv_sqrt_ref_temp, double sqrt_temp)
{
  // include/Transport_model.hpp:35-38; math::pow(x, 3) = ((1*x)*x)*x. sqrt_temp/sqrt_ref_temp is a multiplication by the reciprocal
  // formed once on the host (<= 1 ulp from the reference's division; an FP64 division is ~30 instructions per point on the device)
  const double r = sqrt_temp*inv_sqrt_ref_temp;
  const double cube = r*r*r;
  return t.const_val + t.ref_val*cube*(t.ref_temp + t.temp_offset)/(sqrt_temp*sqrt_temp + t.temp_offset);
}

/* single-precision restatement for the max_dt screen (generic_part.cu); the uniform parameters are converted once per thread */
__device__ __forceinline__ float transport_coef_f(const hexed_b200_transport& t, double inv_sqrt_ref_temp, float sqrt_temp, float temp)
{
  const float r = sqrt_temp*(float)inv_sqrt_ref_temp;
  return (float)t.const_val + (float)(t.ref_val*(t.ref_temp + t.temp_offset))*(r*r*r)*__fdividef(1.f, temp + (float)t.temp_offset);
}

/* ---------------- Navier-Stokes / Euler: include/pde.hpp:27-175 ---------------- */
template <int ND, int RS, bool VISC>
struct PdeNs
{
  static constexpr bool has_diffusion = VISC, has_convection = true, has_source = false;
  static constexpr int n_update = ND + 2, n_state = ND + 4, n_extrap = ND + 2, face_kind = 0;
  /* Max_dt screen: 1/scale (the local time step, reference include/Spatial.hpp:808-822) in single precision, or 0 where the estimate
   * cannot be trusted to 1.5e-5 relative: the temperature is a difference of total and kinetic energy, so above Mach ~7 (internal
   * energy below 3 % of the total) and for inadmissible states the caller must take the FP64 path. Error budget where a value is
   * returned: inputs and arithmetic ~1e-6; temperature <= 1.2e-7*(E + ke)/(E - ke) <= 8e-6, Sutherland's law amplifies it by <= 1.5;
   * the two terms of the scale are positive, so the sum is no worse than its worst term. The screen margin is 1e-4. */
  static constexpr bool has_float_screen = VISC;
  static constexpr float screen_margin = 1.0001f;
  __device__ static float screen_dt(const float (&s)[n_state], const PdeParams& p, float inv_c_h, float inv_d_h2)
  {
    const float rho = s[ND], en = s[ND + 1];
    float sq = 0;
    #pragma unroll
    for (int i = 0; i < ND; ++i) sq += s[i]*s[i];
    const float inv = __fdividef(1.f, rho);
    const float int_ener = en - .5f*sq*inv;
    if (!(int_ener > 0.03f*en)) return 0.f;
    const float speed = __fsqrt_rn((float)(heat_rat_ns*(heat_rat_ns - 1))*en*inv) + __fsqrt_rn(sq)*inv;
    constexpr float gm1_over_r = (float)((heat_rat_ns - 1)/specific_gas_air);
    const float temp = int_ener*inv*gm1_over_r;
    const float sqrt_temp = __fsqrt_rn(temp);
    const float visc = transport_coef_f(p.visc, p.visc_inv_sqrt_ref, sqrt_temp, temp);
    const float cond = transport_coef_f(p.cond, p.cond_inv_sqrt_ref, sqrt_temp, temp)*gm1_over_r;
    const float diffusivity = fabsf(s[ND + 3]) + fmaxf(fabsf(s[ND + 2]) + visc*inv, cond*inv);
    const float r = __fdividef(1.f, speed*inv_c_h + diffusivity*inv_d_h2);
    return (r > 0.f && r < 3.0e38f) ? r : 0.f;
  }
  static constexpr bool needs_av = VISC, needs_forcing = false, needs_adv = false;
  __device__ static constexpr int extrap_slot(int v) { return v; }
  __device__ static constexpr int state_slot(int i) { return i < ND + 2 ? i : i + 1; } // ND+2 -> bulk av (ND+3), ND+3 -> laplacian av (ND+4)
  __device__ static constexpr int update_slot(int v) { return v; }
  __device__ static void write_update(const PdeParams&, const double (&upd)[n_update], double* const (&tgt)[n_update], double, bool)
  {
    #pragma unroll
    for (int v = 0; v < n_update; ++v) *tgt[v] += upd[v];
  }

  template <int NDF>
  struct Comp
  {
    double state[n_state], update_state[n_update], normal[ND][NDF], flux_conv[n_update][NDF];
    double gradient[n_extrap][ND], flux_diff[n_update][ND], source[n_update];
    double mass, inv_mass, kin_ener, pressure, bulk_av, laplacian_av, dyn_visc_coef, energy_cond, char_speed, diffusivity;
    __device__ Comp()
    {
      #pragma unroll
      for (int i = 0; i < ND; ++i)
        #pragma unroll
        for (int j = 0; j < NDF; ++j) normal[i][j] = (i == j);
    }
    __device__ void from_extrap(const double (&x)[n_extrap])
    {
      #pragma unroll
      for (int v = 0; v < ND + 2; ++v) { state[v] = x[v]; update_state[v] = x[v]; }
      state[ND + 2] = 0.; state[ND + 3] = 0.;
    }
    __device__ void scalars_conv()
    {
      // FP64 division costs ~30 instructions on the device: one IEEE reciprocal of the mass per point, then multiplications
      // (<= 1 ulp per use from the reference's divisions, include/pde.hpp:94-158; covered by the 1e-11 / 1e-13 bars)
      mass = state[ND];
      inv_mass = 1./mass;
      kin_ener = 0;
      #pragma unroll
      for (int i = 0; i < ND; ++i) kin_ener += state[i]*state[i];
      kin_ener *= .5*inv_mass;
      pressure = (heat_rat_ns - 1.)*(state[ND + 1] - kin_ener);
    }
    __device__ void compute_flux_conv(const PdeParams&)
    {
      scalars_conv();
      #pragma unroll
      for (int d = 0; d < NDF; ++d) {
        double mass_flux = 0;
        #pragma unroll
        for (int j = 0; j < ND; ++j) mass_flux += state[j]*normal[j][d];
        flux_conv[ND][d] = mass_flux;
        const double vol_flux = mass_flux*inv_mass;
        flux_conv[ND + 1][d] = (state[ND + 1] + pressure)*vol_flux;
        #pragma unroll
        for (int j = 0; j < ND; ++j) flux_conv[j][d] = state[j]*vol_flux + pressure*normal[j][d];
      }
    }
    __device__ void scalars_diff(const PdeParams& p)
    {
      bulk_av = fabs(state[ND + 2]);
      laplacian_av = fabs(state[ND + 3]);
      constexpr double gm1_over_r = (heat_rat_ns - 1)/specific_gas_air;
      const double sqrt_temp = sqrt(fmax((state[ND + 1] - kin_ener)*inv_mass, 0.)*gm1_over_r);
      dyn_visc_coef = transport_coef(p.visc, p.visc_inv_sqrt_ref, sqrt_temp);
      const double therm_cond_coef = transport_coef(p.cond, p.cond_inv_sqrt_ref, sqrt_temp);
      energy_cond = therm_cond_coef*gm1_over_r;
    }
    __device__ void compute_flux_diff(const PdeParams& p)
    {
      if constexpr (NDF == ND) {
        scalars_diff(p);
        double veloc[ND], vgrad[ND][ND], stress[ND][ND], fd[n_update][ND];
        #pragma unroll
        for (int i = 0; i < ND; ++i) veloc[i] = state[i]*inv_mass;
        #pragma unroll
        for (int i = 0; i < ND; ++i)
          #pragma unroll
          for (int j = 0; j < ND; ++j) vgrad[i][j] = (gradient[i][j] - veloc[i]*gradient[ND][j])*inv_mass;
        double trace = 0;
        #pragma unroll
        for (int i = 0; i < ND; ++i) trace += vgrad[i][i];
        const double bulk = (bulk_av*mass - 2./3.*dyn_visc_coef)*trace;
        #pragma unroll
        for (int i = 0; i < ND; ++i)
          #pragma unroll
          for (int j = 0; j < ND; ++j) stress[i][j] = dyn_visc_coef*(vgrad[i][j] + vgrad[j][i]) + (i == j ? bulk : 0.);
        #pragma unroll
        for (int v = 0; v < n_update; ++v)
          #pragma unroll
          for (int j = 0; j < ND; ++j) fd[v][j] = -laplacian_av*gradient[v][j];
        #pragma unroll
        for (int i = 0; i < ND; ++i)
          #pragma unroll
          for (int j = 0; j < ND; ++j) fd[i][j] -= stress[i][j];
        #pragma unroll
        for (int j = 0; j < ND; ++j) {
          double conv = 0, work = 0;
          #pragma unroll
          for (int i = 0; i < ND; ++i) { conv += veloc[i]*vgrad[i][j]; work += veloc[i]*stress[i][j]; }
          const double int_ener_grad = -state[ND + 1]*inv_mass*inv_mass*gradient[ND][j] + gradient[ND + 1][j]*inv_mass - conv;
          fd[ND + 1][j] -= work + energy_cond*int_ener_grad;
        }
        #pragma unroll
        for (int v = 0; v < n_update; ++v)
          #pragma unroll
          for (int k = 0; k < ND; ++k) {
            double s = 0;
            #pragma unroll
            for (int j = 0; j < ND; ++j) s += fd[v][j]*normal[j][k];
            flux_diff[v][k] = s;
          }
      }
    }
    __device__ void compute_source(const PdeParams&) {}
    __device__ void compute_char_speed()
    {
      const double sound_speed = sqrt(heat_rat_ns*(heat_rat_ns - 1)*state[ND + 1]/state[ND]);
      double sq = 0;
      #pragma unroll
      for (int i = 0; i < ND; ++i) sq += state[i]*state[i];
      char_speed = sound_speed + sqrt(sq)/state[ND];
    }
    __device__ void compute_diffusivity(const PdeParams& p)
    {
      scalars_conv();
      scalars_diff(p);
      diffusivity = fabs(laplacian_av) + fmax(fabs(bulk_av) + dyn_visc_coef*inv_mass, energy_cond*inv_mass);
    }
  };
};

/* ---------------- Advection: include/pde.hpp:265-349 ---------------- */
template <int ND, int RS>
struct PdeAdvection
{
  static constexpr bool has_diffusion = false, has_convection = true, has_source = true;
  static constexpr bool has_float_screen = false;
  static constexpr int n_adv = RS, n_state = ND + RS, n_extrap = ND + RS, n_update = RS, face_kind = 2;
  static constexpr bool needs_av = false, needs_forcing = false, needs_adv = true;
  __device__ static constexpr int extrap_slot(int v) { return v < ND ? v : ND + 9 + (v - ND); }
  __device__ static constexpr int state_slot(int i) { return extrap_slot(i); }
  __device__ static constexpr int update_slot(int v) { return ND + 9 + v; }
  __device__ static void write_update(const PdeParams& p, const double (&upd)[n_update], double* const (&tgt)[n_update], double tss, bool critical)
  {
    const double pseudo = 1 + tss*2/p.p0;
    #pragma unroll
    for (int a = 0; a < n_adv; ++a) {
      double d = *tgt[a];
      if (critical) d = (d + upd[a])/pseudo;
      else d += upd[a]/pseudo;
      *tgt[a] = d;
    }
  }
  template <int NDF>
  struct Comp
  {
    double state[n_state], update_state[n_update], normal[ND][NDF], flux_conv[n_update][NDF];
    double gradient[1][1], flux_diff[1][1], source[n_update], char_speed, diffusivity;
    __device__ Comp()
    {
      #pragma unroll
      for (int i = 0; i < ND; ++i)
        #pragma unroll
        for (int j = 0; j < NDF; ++j) normal[i][j] = (i == j);
    }
    __device__ void from_extrap(const double (&x)[n_extrap])
    {
      #pragma unroll
      for (int v = 0; v < n_extrap; ++v) state[v] = x[v];
      #pragma unroll
      for (int a = 0; a < n_adv; ++a) update_state[a] = x[ND + a];
    }
    __device__ void compute_flux_conv(const PdeParams& p)
    {
      #pragma unroll
      for (int d = 0; d < NDF; ++d) {
        double nv = 0;
        #pragma unroll
        for (int j = 0; j < ND; ++j) nv += state[j]*normal[j][d];
        #pragma unroll
        for (int a = 0; a < n_adv; ++a) flux_conv[a][d] = p.adv_nodes[a]*nv*state[ND + a];
      }
    }
    __device__ void compute_flux_diff(const PdeParams&) {}
    __device__ void compute_source(const PdeParams& p)
    {
      #pragma unroll
      for (int a = 0; a < n_update; ++a) source[a] = 2/p.p0;
    }
    __device__ void compute_char_speed()
    {
      double sq = 0;
      #pragma unroll
      for (int i = 0; i < ND; ++i) sq += state[i]*state[i];
      char_speed = fmax(1., sqrt(sq));
    }
    __device__ void compute_diffusivity(const PdeParams&) {}
  };
};

/* ---------------- pure diffusion PDEs: Smooth_art_visc include/pde.hpp:355-431, Fix_therm_admis :437-493 ---------------- */
template <int ND, int NU, int NE, int NS>
struct DiffusionComp
{
  double state[NS], update_state[NU], normal[ND][ND], flux_conv[1][1];
  double gradient[NE][ND], flux_diff[NU][ND], source[NU], char_speed, diffusivity;
  __device__ DiffusionComp()
  {
    #pragma unroll
    for (int i = 0; i < ND; ++i)
      #pragma unroll
      for (int j = 0; j < ND; ++j) normal[i][j] = (i == j);
  }
  __device__ void compute_flux_conv(const PdeParams&) {}
  __device__ void compute_flux_diff(const PdeParams&)
  {
    #pragma unroll
    for (int v = 0; v < NU; ++v)
      #pragma unroll
      for (int k = 0; k < ND; ++k) {
        double s = 0;
        #pragma unroll
        for (int j = 0; j < ND; ++j) s += -gradient[v][j]*normal[j][k];
        flux_diff[v][k] = s;
      }
  }
  __device__ void compute_char_speed() {}
  __device__ void compute_diffusivity(const PdeParams&) { diffusivity = 1; }
};

template <int ND, int RS>
struct PdeSmoothAv
{
  static constexpr bool has_diffusion = true, has_convection = false, has_source = true;
  static constexpr bool has_float_screen = false;
  static constexpr int n_state = 4, n_extrap = 3, n_update = 3, face_kind = 0;
  static constexpr bool needs_av = false, needs_forcing = true, needs_adv = false;
  __device__ static constexpr int extrap_slot(int v) { return ND + 5 + 1 + v; }
  __device__ static constexpr int state_slot(int i) { return ND + 5 + i; }
  __device__ static constexpr int update_slot(int v) { return ND + 5 + 1 + v; }
  __device__ static void write_update(const PdeParams& p, const double (&upd)[n_update], double* const (&tgt)[n_update], double tss, bool critical)
  {
    const double pseudo = 1 + tss*p.p1/p.p0;
    #pragma unroll
    for (int v = 0; v < n_update; ++v) {
      double d = *tgt[v] + upd[v];
      if (critical) d /= pseudo;
      *tgt[v] = d;
    }
  }
  template <int NDF>
  struct Comp : DiffusionComp<ND, n_update, n_extrap, n_state>
  {
    __device__ void from_extrap(const double (&x)[n_extrap])
    {
      #pragma unroll
      for (int v = 0; v < n_update; ++v) this->update_state[v] = x[v];
    }
    __device__ void compute_source(const PdeParams& p)
    {
      #pragma unroll
      for (int v = 0; v < n_update; ++v) {
        const double f = fabs(this->state[v]);
        this->source[v] = ((v == 1) ? sqrt(f) : f)/p.p0;
      }
    }
  };
};

template <int ND, int RS>
struct PdeFta
{
  static constexpr bool has_diffusion = true, has_convection = false, has_source = false;
  static constexpr bool has_float_screen = false;
  static constexpr int n_state = ND + 2, n_extrap = ND + 2, n_update = ND + 2, face_kind = 0;
  static constexpr bool needs_av = false, needs_forcing = false, needs_adv = false;
  __device__ static constexpr int extrap_slot(int v) { return v; }
  __device__ static constexpr int state_slot(int i) { return i; }
  __device__ static constexpr int update_slot(int v) { return v; }
  __device__ static void write_update(const PdeParams&, const double (&upd)[n_update], double* const (&tgt)[n_update], double, bool)
  {
    #pragma unroll
    for (int v = 0; v < n_update; ++v) *tgt[v] += upd[v];
  }
  template <int NDF>
  struct Comp : DiffusionComp<ND, n_update, n_extrap, n_state>
  {
    __device__ void from_extrap(const double (&x)[n_extrap])
    {
      #pragma unroll
      for (int v = 0; v < n_update; ++v) this->update_state[v] = x[v];
    }
    __device__ void compute_source(const PdeParams&) {}
  };
};

template <int PDE, int ND, int RS> struct PdeSelect;
template <int ND, int RS> struct PdeSelect<PDE_EULER, ND, RS> { using type = PdeNs<ND, RS, false>; };
template <int ND, int RS> struct PdeSelect<PDE_NAVIER_STOKES, ND, RS> { using type = PdeNs<ND, RS, true>; };
template <int ND, int RS> struct PdeSelect<PDE_ADVECTION, ND, RS> { using type = PdeAdvection<ND, RS>; };
template <int ND, int RS> struct PdeSelect<PDE_SMOOTH_AV, ND, RS> { using type = PdeSmoothAv<ND, RS>; };
template <int ND, int RS> struct PdeSelect<PDE_FIX_THERM_ADMIS, ND, RS> { using type = PdeFta<ND, RS>; };

} // namespace hb
#endif
