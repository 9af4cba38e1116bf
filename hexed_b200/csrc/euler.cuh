/* euler.cuh -- pointwise Euler physics shared by the kernels.
 * Follows pde::Navier_stokes<false>::Pde::Computation (reference include/pde.hpp:94-122,160-166):
 * state order [momentum_0..nd-1, mass, total energy], heat ratio 1.4. Divisions by the mass are done with one
 * IEEE reciprocal per point and multiplications (<= 1 ulp from the reference's divisions; covered by the 1e-11 bar). */
#ifndef HB_EULER_CUH_
#define HB_EULER_CUH_
#include "common.cuh"

namespace hb {

constexpr double heat_rat = 1.4;

template <int ND>
struct EulerPoint
{
  double s[ND + 2];
  double inv_mass, pressure;
  __device__ __forceinline__ void scalars()
  {
    inv_mass = 1./s[ND];
    double ke = 0;
    #pragma unroll
    for (int i = 0; i < ND; ++i) ke += s[i]*s[i];
    ke *= .5*inv_mass;
    pressure = (heat_rat - 1.)*(s[ND + 1] - ke);
  }
  /* flux through the (area-weighted) direction n */
  __device__ __forceinline__ void flux(const double (&n)[ND], double (&f)[ND + 2]) const
  {
    double mass_flux = 0;
    #pragma unroll
    for (int j = 0; j < ND; ++j) mass_flux += s[j]*n[j];
    const double vol_flux = mass_flux*inv_mass;
    f[ND] = mass_flux;
    f[ND + 1] = (s[ND + 1] + pressure)*vol_flux;
    #pragma unroll
    for (int j = 0; j < ND; ++j) f[j] = s[j]*vol_flux + pressure*n[j];
  }
  /* flux along reference direction d of a Cartesian element (unit normal e_d) */
  __device__ __forceinline__ void flux_axis(int d, double (&f)[ND + 2]) const
  {
    const double mass_flux = s[d];
    const double vol_flux = mass_flux*inv_mass;
    f[ND] = mass_flux;
    f[ND + 1] = (s[ND + 1] + pressure)*vol_flux;
    #pragma unroll
    for (int j = 0; j < ND; ++j) f[j] = s[j]*vol_flux + (j == d ? pressure : 0.);
  }
  __device__ __forceinline__ double char_speed() const
  {
    const double sound = sqrt(heat_rat*(heat_rat - 1)*s[ND + 1]*inv_mass);
    double sq = 0;
    #pragma unroll
    for (int i = 0; i < ND; ++i) sq += s[i]*s[i];
    return sound + sqrt(sq)*inv_mass;
  }
  /* c_spacing/char_speed() (the local time step of Max_dt, reference include/Spatial.hpp:808-822) multiplied through by the density:
   * one division and no reciprocal of the mass (FP64 division is ~30 instructions and max_dt_euler_kernel is FP64-issue-bound);
   * differs from the two-division form by a few ulp, 3 orders inside the 1e-13 bar. Does not need scalars(). */
  __device__ __forceinline__ double cfl_time_step(double c_spacing) const
  {
    double sq = 0;
    #pragma unroll
    for (int i = 0; i < ND; ++i) sq += s[i]*s[i];
    return c_spacing*s[ND]/(sqrt(heat_rat*(heat_rat - 1)*s[ND + 1]*s[ND]) + sqrt(sq));
  }
};

/* n-linear interpolation of the vertex time-step scale to quadrature point q (reference include/math.hpp:207-218 as used by
 * Spatial::Max_dt, include/Spatial.hpp:800-806); `vt` = the element's 2^ND vertex values */
template <int ND, int RS>
__device__ __forceinline__ double interp_vertex_spacing(const double* vt, const Ops& ops, int q)
{
  constexpr int n_vert = ipow(2, ND);
  double vals[n_vert];
  #pragma unroll
  for (int i = 0; i < n_vert; ++i) vals[i] = vt[i];
  int stride = n_vert;
  #pragma unroll
  for (int d = 0; d < ND; ++d) {
    const double coord = ops.node[(q/ipow(RS, ND - 1 - d)) % RS];
    stride /= 2;
    #pragma unroll
    for (int i = 0; i < n_vert/2; ++i) if (i < stride) vals[i] += coord*(vals[i + stride] - vals[i]);
  }
  return vals[0];
}

} // namespace hb
#endif
