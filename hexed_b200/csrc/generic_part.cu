/* generic_part.cu -- PDE-generic kernels (one translation unit per PDE: compile with -DHB_PDE=<1..4>).
 *
 * Covers the reference kernels for pde::Navier_stokes<true>, pde::Advection, pde::Smooth_art_visc and pde::Fix_therm_admis:
 *   Spatial<Pde, is_deformed>::Neighbor            include/Spatial.hpp:613-704   (LLF flux and/or LDG average state)
 *   Spatial<Pde, is_deformed>::Local               include/Spatial.hpp:326-509   (gradient, convective/diffusive flux, source, update)
 *   Spatial<Pde, is_deformed>::Neighbor_reconcile  include/Spatial.hpp:716-759
 *   Spatial<Pde, is_deformed>::Reconcile_ldg_flux  include/Spatial.hpp:543-594
 *   Spatial<Pde, is_deformed>::Max_dt              include/Spatial.hpp:784-828
 *   Spatial<Pde, false>::Write_face                include/Spatial.hpp:41-70
 * The inviscid Euler path has its own kernels (local_euler.cu, neighbor_euler.cu); these share their structure:
 * one thread per quadrature point, line contractions through shared memory, connection tables for the faces.
 */
#include "pde.cuh"

#ifndef HB_PDE
#error "compile with -DHB_PDE=<1..4>"
#endif

namespace hb {
namespace {

struct GArgs
{
  ElemData ed;
  const double* nom; const double* refn; const double* det; const double* normals; const double* vtss;
  double* faces; double* faces_ldg; int face_width;
  const int* con; const int* perm; int n_con;
  int elem_begin, elem_end, n_car;
  double update; int stage, compute_residual, use_filter;
  const double* dt_dev; // non-null: the time step lives on the device (hexed_b200_update_*) and multiplies `update`
  double max_cfl_c, max_cfl_d, inv_max_cfl_c, inv_max_cfl_d; int is_local; unsigned long long* global_min; int write_tss = 1; // max_dt, global stepping: 0 = time_step_scale already holds 1 everywhere
  PdeParams pp;
};

template <class K> int set_smem(hexed_b200_ctx* c, K k, size_t smem)
{
  if (smem > 48*1024) {
    HB_CUDA(c, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // several such CTAs per SM only fit with the largest shared-memory carve-out
    HB_CUDA(c, cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
  }
  return 0;
}

template <int ND, int RS> struct Line
{
  int stride, node, base, fq;
  __device__ Line(int d, int q)
  {
    stride = 1;
    for (int i = 0; i < ND - 1 - d; ++i) stride *= RS;
    node = (q/stride) % RS;
    base = q - node*stride;
    fq = (q/(stride*RS))*stride + q % stride;
  }
};

/* ---------------- Neighbor ---------------- */
template <int ND, int RS, class P, bool DEF>
__global__ void __launch_bounds__(128, 8)
g_neighbor_kernel(GArgs a)
{
  constexpr int nfq = ipow(RS, ND - 1), ne = P::n_extrap, nu = P::n_update, wl = (ND + 2)*nfq;
  const long long gid = (long long)blockIdx.x*blockDim.x + threadIdx.x;
  const int con = (int)(gid/nfq), q0 = (int)(gid % nfq);
  if (con >= a.n_con) return;
  const int* tab = a.con + (size_t)con*4;
  const int slot0 = tab[0], slot1 = tab[1];
  int q1 = q0;
  double n[ND];
  double sign0 = 1., sign1 = 1.;
  if constexpr (DEF) {
    const int code = tab[2];
    q1 = a.perm[code*nfq + q0];
    sign0 = ((code/9) % 2) ? 1. : -1.;   // flip_normal(0) = (face_sign[0] == 0)   include/Kernel_connection.hpp:19
    sign1 = ((code/18) % 2) ? -1. : 1.;  // flip_normal(1) = (face_sign[1] == 1)
    const double* nr = a.normals + (size_t)tab[3]*ND*nfq;
    #pragma unroll
    for (int d = 0; d < ND; ++d) n[d] = sign0*nr[d*nfq + q0];
  } else {
    #pragma unroll
    for (int d = 0; d < ND; ++d) n[d] = (d == tab[2]) ? 1. : 0.;
  }
  double* f0 = a.faces + (size_t)slot0*a.face_width + q0;
  double* f1 = a.faces + (size_t)slot1*a.face_width + q1;
  double x0[ne], x1[ne];
  #pragma unroll
  for (int v = 0; v < ne; ++v) { x0[v] = f0[v*nfq]; x1[v] = f1[v*nfq]; }
  if constexpr (P::has_diffusion) {
    double* l0 = a.faces_ldg + (size_t)slot0*wl + q0;
    double* l1 = a.faces_ldg + (size_t)slot1*wl + q1;
    #pragma unroll
    for (int v = 0; v < ne; ++v) { const double avg = .5*(x0[v] + x1[v]); l0[v*nfq] = avg; l1[v*nfq] = avg; }
  }
  if constexpr (P::has_convection) {
    typename P::template Comp<1> c0, c1;
    #pragma unroll
    for (int d = 0; d < ND; ++d) { c0.normal[d][0] = n[d]; c1.normal[d][0] = n[d]; }
    c0.from_extrap(x0); c1.from_extrap(x1);
    c0.compute_flux_conv(a.pp); c0.compute_char_speed();
    c1.compute_flux_conv(a.pp); c1.compute_char_speed();
    double nsq = 0;
    #pragma unroll
    for (int d = 0; d < ND; ++d) nsq += n[d]*n[d];
    const double speed = fmax(c0.char_speed, c1.char_speed)*sqrt(nsq);
    #pragma unroll
    for (int v = 0; v < nu; ++v) {
      const double flux = .5*(c0.flux_conv[v][0] + c1.flux_conv[v][0] + speed*(c0.update_state[v] - c1.update_state[v]));
      f0[v*nfq] = sign0*flux;
      f1[v*nfq] = sign1*flux;
    }
  }
}

/* ---------------- Neighbor_reconcile ---------------- */
template <int ND, int RS, class P, bool DEF>
__global__ void __launch_bounds__(128)
g_neighbor_reconcile_kernel(GArgs a)
{
  constexpr int nfq = ipow(RS, ND - 1), nu = P::n_update, wl = (ND + 2)*nfq;
  const long long gid = (long long)blockIdx.x*blockDim.x + threadIdx.x;
  const int con = (int)(gid/nfq), q0 = (int)(gid % nfq);
  if (con >= a.n_con) return;
  const int* tab = a.con + (size_t)con*4;
  int q1 = q0;
  double sign0 = 1., sign1 = 1.;
  if constexpr (DEF) {
    const int code = tab[2];
    q1 = a.perm[code*nfq + q0];
    sign0 = ((code/9) % 2) ? 1. : -1.;
    sign1 = ((code/18) % 2) ? -1. : 1.;
  }
  double* l0 = a.faces_ldg + (size_t)tab[0]*wl + q0;
  double* l1 = a.faces_ldg + (size_t)tab[1]*wl + q1;
  // every load before the first store: interleaved, each store had to be assumed to alias the next variable's loads, which chained
  // n_update memory round trips per thread (202 cycles of long-scoreboard stall per issue, 4.0 TB/s, profiles/r01g_ncu_full_ns.md)
  double g0[nu], g1[nu];
  #pragma unroll
  for (int v = 0; v < nu; ++v) { g0[v] = l0[v*nfq]; g1[v] = l1[v*nfq]; }
  #pragma unroll
  for (int v = 0; v < nu; ++v) {
    double avg = 0;
    avg += .5*sign0*g0[v];
    avg += .5*sign1*g1[v];
    l0[v*nfq] = sign0*avg - g0[v];
    l1[v*nfq] = sign1*avg - g1[v];
  }
}

/* ---------------- Local ---------------- */
template <int ND, int RS, class P>
struct GCfg
{
  static constexpr int nq = ipow(RS, ND), nfq = nq/RS, ne = P::n_extrap, nu = P::n_update;
  static constexpr int epb = nq >= 128 ? 1 : (128 + nq - 1)/nq;
  static constexpr int threads = epb*nq;
  static constexpr int nbuf = (ND > 2 ? ND : 2)*(nu > ne ? nu : ne)*nq;  // flux of one kind / filter scratch / face-extrapolation scratch
  static constexpr int nx = P::has_diffusion ? ne*nq : 0;                   // variables whose gradient is taken
  static constexpr int nface = P::has_convection ? 2*ND*nu*nfq : 0;         // numerical convective flux on the element's faces
  static constexpr int nvface = P::has_diffusion ? 2*ND*ne*nfq : 0;         // LDG face state
  static constexpr int nops = 3*RS*RS + 2*RS;                               // dfull, diff, filter, lift
  static constexpr int per_elem = nbuf + nx + nface + nvface;
  static constexpr int smem_doubles = nops + epb*per_elem;
  // Reconcile_ldg_flux needs only one scratch field set and the LDG faces: a smaller footprint keeps more CTAs resident
  static constexpr int rec_buf = (nu > ne ? nu : ne)*nq;
  static constexpr int rec_per_elem = rec_buf + nvface;
  static constexpr int rec_smem_doubles = nops + epb*rec_per_elem;
};

template <int ND, int RS, class P>
__device__ __forceinline__ void load_ops(double* s_ops, const Ops& ops, const FilterOp& filt, int t, int nt)
{
  for (int i = t; i < RS*RS; i += nt) {
    s_ops[i] = ops.dfull[i/RS][i % RS];
    s_ops[RS*RS + i] = ops.diff[i/RS][i % RS];
    s_ops[2*RS*RS + i] = filt.filter[i/RS][i % RS];
  }
  for (int i = t; i < 2*RS; i += nt) s_ops[3*RS*RS + i] = ops.lift[i/2][i % 2];
}

/* extrapolate [n_var][nq] values held in shared memory to the 2*ND faces of element e: dst[(e*2ND + f)*width + v*nfq + fq] */
template <int ND, int RS>
__device__ __forceinline__ void extrapolate_faces(const double* sval, int n_var, double* dst_elem, int width, const Ops& ops, int q)
{
  constexpr int nq = ipow(RS, ND), nfq = nq/RS;
  for (int item = q; item < ND*n_var*nfq; item += nq) {
    const int d = item/(n_var*nfq), v = (item/nfq) % n_var, fq = item % nfq;
    int stride = 1;
    for (int i = 0; i < ND - 1 - d; ++i) stride *= RS;
    const int base = (fq/stride)*stride*RS + fq % stride;
    double e0 = 0, e1 = 0;
    #pragma unroll
    for (int k = 0; k < RS; ++k) {
      const double x = sval[v*nq + base + k*stride];
      e0 += ops.bnd[0][k]*x;
      e1 += ops.bnd[1][k]*x;
    }
    dst_elem[(size_t)(2*d)*width + v*nfq + fq] = e0;
    dst_elem[(size_t)(2*d + 1)*width + v*nfq + fq] = e1;
  }
}

template <int ND, int RS, class P, bool DEF>
__global__ void __launch_bounds__(GCfg<ND, RS, P>::threads)
g_local_kernel(GArgs a, Ops ops, FilterOp filt)
{
  using C = GCfg<ND, RS, P>;
  constexpr int nq = C::nq, nfq = C::nfq, ne = C::ne, nu = C::nu, nv = ND + 2, wl = nv*nfq;
  constexpr int cache_slot0 = nv + 7 + RS;
  HB_DYN_SMEM(double, smem);
  const int t = threadIdx.x;
  const int le = t/nq, q = t % nq;
  const int e = a.elem_begin + blockIdx.x*C::epb + le;
  const bool active = e < a.elem_end;
  double* s_dfull = smem; double* s_diff = smem + RS*RS; double* s_filt = smem + 2*RS*RS; double* s_lift = smem + 3*RS*RS;
  double* sbuf = smem + C::nops + le*C::per_elem;
  double* sx = sbuf + C::nbuf;
  double* sface = sx + C::nx;
  double* svface = sface + C::nface;
  load_ops<ND, RS, P>(smem, ops, filt, t, C::threads);

  typename P::template Comp<ND> comp;
  double det = 1., tss = 0., nom = 1., inv_nom = 1.;
  if (active) {
    if constexpr (P::has_convection) {
      const double* base = a.faces + (size_t)e*2*ND*a.face_width;
      for (int i = q; i < 2*ND*nu*nfq; i += nq) sface[i] = base[(size_t)(i/(nu*nfq))*a.face_width + i % (nu*nfq)];
    }
    if constexpr (P::has_diffusion) {
      const double* base = a.faces_ldg + (size_t)e*2*ND*wl;
      for (int i = q; i < 2*ND*ne*nfq; i += nq) svface[i] = base[(size_t)(i/(ne*nfq))*wl + i % (ne*nfq)];
    }
    #pragma unroll
    for (int i = 0; i < P::n_state; ++i) comp.state[i] = a.ed.template slot<ND, RS>(e, P::state_slot(i))[q];
    tss = a.ed.tss[(size_t)e*nq + q];
    nom = a.nom[e];
    inv_nom = 1./nom; // one reciprocal instead of a division per gradient term (<= 1 ulp each; FP64 division is ~30 instructions)
    if constexpr (DEF) {
      det = a.det[(size_t)(e - a.n_car)*nq + q];
      const double* rn = a.refn + (size_t)(e - a.n_car)*ND*ND*nq;
      #pragma unroll
      for (int d = 0; d < ND; ++d)
        #pragma unroll
        for (int j = 0; j < ND; ++j) comp.normal[j][d] = rn[(d*ND + j)*nq + q];
    }
    if constexpr (P::has_diffusion) {
      #pragma unroll
      for (int v = 0; v < ne; ++v) sx[v*nq + q] = a.ed.template slot<ND, RS>(e, P::extrap_slot(v))[q];
    }
  }
  __syncthreads();

  // gradient of the extrapolated variables (times the jacobian determinant for deformed elements): Spatial.hpp:371-402
  if constexpr (P::has_diffusion) {
    if (active) {
      #pragma unroll
      for (int v = 0; v < ne; ++v)
        #pragma unroll
        for (int j = 0; j < ND; ++j) comp.gradient[v][j] = 0.;
      #pragma unroll
      for (int d = 0; d < ND; ++d) {
        const Line<ND, RS> ln(d, q);
        double m[RS];
        #pragma unroll
        for (int k = 0; k < RS; ++k) m[k] = s_dfull[ln.node*RS + k];
        const double l0 = s_lift[ln.node*2], l1 = s_lift[ln.node*2 + 1];
        if constexpr (DEF) {
          const double* rn = a.refn + (size_t)(e - a.n_car)*ND*ND*nq;
          const double* fn = a.normals + (size_t)(e - a.n_car)*2*ND*ND*nfq;
          #pragma unroll
          for (int j = 0; j < ND; ++j) {
            double nk[RS];
            #pragma unroll
            for (int k = 0; k < RS; ++k) nk[k] = rn[(d*ND + j)*nq + ln.base + k*ln.stride];
            const double fn0 = fn[((2*d)*ND + j)*nfq + ln.fq], fn1 = fn[((2*d + 1)*ND + j)*nfq + ln.fq];
            #pragma unroll
            for (int v = 0; v < ne; ++v) {
              double acc = 0;
              #pragma unroll
              for (int k = 0; k < RS; ++k) acc += m[k]*(nk[k]*sx[v*nq + ln.base + k*ln.stride]);
              acc += l0*(fn0*svface[((2*d)*ne + v)*nfq + ln.fq]);
              acc += l1*(fn1*svface[((2*d + 1)*ne + v)*nfq + ln.fq]);
              comp.gradient[v][j] += acc*inv_nom;
            }
          }
        } else {
          #pragma unroll
          for (int v = 0; v < ne; ++v) {
            double acc = 0;
            #pragma unroll
            for (int k = 0; k < RS; ++k) acc += m[k]*sx[v*nq + ln.base + k*ln.stride];
            acc += l0*svface[((2*d)*ne + v)*nfq + ln.fq];
            acc += l1*svface[((2*d + 1)*ne + v)*nfq + ln.fq];
            comp.gradient[v][d] = acc*inv_nom;
          }
        }
      }
      if constexpr (DEF) {
        const double inv_det = 1./det;
        #pragma unroll
        for (int v = 0; v < ne; ++v)
          #pragma unroll
          for (int j = 0; j < ND; ++j) comp.gradient[v][j] *= inv_det;
      }
    }
  }

  // convective flux and its derivative: Spatial.hpp:405-419,445-456
  double r0[nu], r1[nu];
  #pragma unroll
  for (int v = 0; v < nu; ++v) { r0[v] = 0.; r1[v] = 0.; }
  if constexpr (P::has_convection) {
    if (active) {
      comp.compute_flux_conv(a.pp);
      #pragma unroll
      for (int d = 0; d < ND; ++d)
        #pragma unroll
        for (int v = 0; v < nu; ++v) sbuf[(d*nu + v)*nq + q] = comp.flux_conv[v][d];
    }
    __syncthreads();
    if (active) {
      #pragma unroll
      for (int d = 0; d < ND; ++d) {
        const Line<ND, RS> ln(d, q);
        double m[RS];
        #pragma unroll
        for (int k = 0; k < RS; ++k) m[k] = s_dfull[ln.node*RS + k];
        const double l0 = s_lift[ln.node*2], l1 = s_lift[ln.node*2 + 1];
        #pragma unroll
        for (int v = 0; v < nu; ++v) {
          const double* row = sbuf + (d*nu + v)*nq + ln.base;
          double acc = 0;
          #pragma unroll
          for (int k = 0; k < RS; ++k) acc += m[k]*row[k*ln.stride];
          acc += l0*sface[((2*d)*nu + v)*nfq + ln.fq];
          acc += l1*sface[((2*d + 1)*nu + v)*nfq + ln.fq];
          r0[v] -= acc;
        }
      }
    }
    __syncthreads();
  }
  // source: Spatial.hpp:436-441
  if constexpr (P::has_source) {
    if (active && !a.stage) {
      comp.compute_source(a.pp);
      double mult = nom;
      if constexpr (DEF) mult *= det;
      #pragma unroll
      for (int v = 0; v < nu; ++v) r1[v] = mult*comp.source[v];
    }
  }
  // diffusive flux, its interior derivative and its extrapolation to the LDG faces: Spatial.hpp:420-435,458-468
  if constexpr (P::has_diffusion) {
    if (active) {
      comp.compute_flux_diff(a.pp);
      #pragma unroll
      for (int d = 0; d < ND; ++d)
        #pragma unroll
        for (int v = 0; v < nu; ++v) sbuf[(d*nu + v)*nq + q] = comp.flux_diff[v][d];
    }
    __syncthreads();
    if (active) {
      #pragma unroll
      for (int d = 0; d < ND; ++d) {
        const Line<ND, RS> ln(d, q);
        double m[RS];
        #pragma unroll
        for (int k = 0; k < RS; ++k) m[k] = s_diff[ln.node*RS + k];
        #pragma unroll
        for (int v = 0; v < nu; ++v) {
          const double* row = sbuf + (d*nu + v)*nq + ln.base;
          double acc = 0;
          #pragma unroll
          for (int k = 0; k < RS; ++k) acc += m[k]*row[k*ln.stride];
          r1[v] -= acc;
        }
      }
      double* dst = a.faces_ldg + (size_t)e*2*ND*wl;
      for (int item = q; item < ND*nu*nfq; item += nq) {
        const int d = item/(nu*nfq), v = (item/nfq) % nu, fq = item % nfq;
        int stride = 1;
        for (int i = 0; i < ND - 1 - d; ++i) stride *= RS;
        const int base = (fq/stride)*stride*RS + fq % stride;
        double e0 = 0, e1 = 0;
        #pragma unroll
        for (int k = 0; k < RS; ++k) {
          const double x = sbuf[(d*nu + v)*nq + base + k*stride];
          e0 += ops.bnd[0][k]*x;
          e1 += ops.bnd[1][k]*x;
        }
        dst[(size_t)(2*d)*wl + v*nfq + fq] = e0;
        dst[(size_t)(2*d + 1)*wl + v*nfq + fq] = e1;
      }
    }
    __syncthreads();
  }

  // modal filter of both parts of the time rate: Spatial.hpp:473-481
  if (a.use_filter) {
    #pragma unroll
    for (int d = 0; d < ND; ++d) {
      if (active) {
        #pragma unroll
        for (int v = 0; v < nu; ++v) { sbuf[v*nq + q] = r0[v]; sbuf[(nu + v)*nq + q] = r1[v]; }
      }
      __syncthreads();
      if (active) {
        const Line<ND, RS> ln(d, q);
        #pragma unroll
        for (int v = 0; v < nu; ++v) {
          double acc0 = 0, acc1 = 0;
          for (int k = 0; k < RS; ++k) {
            const double f = s_filt[ln.node*RS + k];
            acc0 += f*sbuf[v*nq + ln.base + k*ln.stride];
            acc1 += f*sbuf[(nu + v)*nq + ln.base + k*ln.stride];
          }
          r0[v] = acc0; r1[v] = acc1;
        }
      }
      __syncthreads();
    }
  }

  // update: Spatial.hpp:484-503
  if (active) {
    double mult = (a.dt_dev ? *a.dt_dev*a.update : a.update)*tss/nom;
    if constexpr (DEF) mult /= det;
    double upd[nu]; double* tgt[nu];
    #pragma unroll
    for (int v = 0; v < nu; ++v) {
      double* cache = a.ed.template slot<ND, RS>(e, cache_slot0 + v) + q;
      double u = r0[v];
      if (a.stage) u -= *cache;
      else {
        if constexpr (P::has_convection) *cache = u;
        if constexpr (P::has_diffusion || P::has_source) u += r1[v];
      }
      u *= mult;
      upd[v] = 0.;
      if (a.compute_residual) *cache = u;
      else upd[v] = u;
      tgt[v] = a.ed.template slot<ND, RS>(e, P::update_slot(v)) + q;
    }
    P::write_update(a.pp, upd, tgt, tss, !P::has_diffusion && !a.stage);
  }
  // face extrapolation of the updated state (inviscid PDEs only): Spatial.hpp:507
  if constexpr (!P::has_diffusion) {
    if (active) {
      #pragma unroll
      for (int v = 0; v < ne; ++v) sbuf[v*nq + q] = a.ed.template slot<ND, RS>(e, P::extrap_slot(v))[q];
    }
    __syncthreads();
    if (active) extrapolate_faces<ND, RS>(sbuf, ne, a.faces + (size_t)e*2*ND*a.face_width, a.face_width, ops, q);
  }
}

#if HB_PDE == 1 /* PDE_NAVIER_STOKES */
/* ---------------- Navier-Stokes Local, 3-D, line-task formulation ----------------
 * Same arithmetic as g_local_kernel<3, RS, PdeNs<3, RS, true>, DEF> without the modal filter (reference include/Spatial.hpp:326-509
 * for pde::Navier_stokes<true>), reorganised because the first profile of the point-per-thread kernel (profiles/r01c_ncu_full_ns.md)
 * showed it instruction- and occupancy-bound: 220 registers -> one 216-thread CTA per SM, 4 000 instructions per point, 9 % of the
 * HBM roofline. Here one thread owns a LINE of row_size points wherever a 1-D operator is applied (the products along the line are
 * formed once and reused for row_size outputs, operator entries are constant-bank operands) and a POINT only for the physics:
 *   P1  gradient of the state: one sub-phase per reference direction d; task = (line of d, physical component j) accumulates
 *       D_d(n_dj * u) + lifted LDG face terms into G[v][j]                          (Spatial.hpp:371-402)
 *   P2  pointwise: convective and diffusive flux in reference directions; F <- flux_conv, G <- flux_diff   (:405-435)
 *   P3  line (d, l): F <- -(D_d flux_conv + lifted numerical flux) in place; diffusive flux extrapolated to faces 2d, 2d+1 -> LDG
 *       face storage; G <- -diff_mat_d flux_diff in place                           (:445-468)
 *   P4  pointwise: sum over d, residual cache, update                                (:484-503)
 * Every input of the element (state, LDG faces, flux faces, normals, face normals, determinant, tss, AV coefficients) is one
 * contiguous run, fetched by 1-D bulk TMA copies onto one mbarrier at the top (the second profile, profiles/r01e_ncu_full_ns.md,
 * showed the version with per-thread global loads stalled on them: long-scoreboard 6.5 cycles per issue at 18 % occupancy);
 * with G and F that is 105 KB at row size 6 -> 2 CTAs of 128 threads per SM, one loading while the other computes. */
template <int RS, bool DEF>
struct NsCfg
{
  static constexpr int ND = 3, nq = RS*RS*RS, nfq = RS*RS, nv = 5, n_line = ND*nfq;
  static constexpr int threads = ((n_line + 31)/32)*32;
  // Shared-memory plan, sized so that THREE CTAs are resident per SM at row size 6 (72.6 KB each; the first version staged every input
  // and kept G and F apart: 105 KB, two CTAs, 12 % occupancy, profiles/r01h_ncu_full_ns.md):
  //   state | flux faces | region A = [LDG faces | reference normals | face normals] | G
  // fetched by 1-D bulk TMA copies onto one mbarrier. F (the convective flux, then its derivative) ALIASES region A: the LDG faces and
  // face normals are dead after P1 (a barrier separates P1 from P2), and F's fields 5..13 coincide field for field with the nine
  // reference normals, which the thread that writes F[.][q] has itself just read at the same q and nobody else reads at that q.
  // The per-point scalars (tss, determinant, AV coefficients; used once or twice per point) are read straight from HBM after a
  // prefetch issued at the top of the kernel.
  static constexpr int s_state = 0, s_fc = s_state + nv*nq, s_ldg = s_fc + 2*ND*nv*nfq;
  static constexpr int s_nrml = s_ldg + 2*ND*nv*nfq, s_fn = s_nrml + (DEF ? ND*ND*nq : 0);
  static constexpr int a_end = s_fn + (DEF ? 2*ND*ND*nfq : 0);
  static constexpr int s_flux = s_nrml - nv*nq; // aliases region A so that F field 5 lands on the first reference normal (= s_ldg at row size 6)
  static_assert(s_flux >= s_ldg, "the LDG faces must be at least as large as nv fields (row size <= 6)");
  static constexpr int s_grad = (a_end > s_flux + ND*nv*nq) ? a_end : s_flux + ND*nv*nq;
  static constexpr int smem_doubles = s_grad + ND*nv*nq;
  static constexpr size_t smem_bytes = sizeof(double)*smem_doubles + 2*sizeof(mbar_t);
};

template <int RS, bool DEF>
__global__ void __launch_bounds__(NsCfg<RS, DEF>::threads, 3)
ns_local_line_kernel(GArgs a, Ops ops)
{
  using C = NsCfg<RS, DEF>;
  using P = PdeNs<3, RS, true>;
  constexpr int ND = 3, nq = C::nq, nfq = C::nfq, nv = C::nv, wl = nv*nfq, T = C::threads;
  constexpr int cs = nv > RS ? nv : RS;
  HB_DYN_SMEM(double, smem);
  double* S = smem + C::s_state;
  const double* fldg = smem + C::s_ldg;
  const double* fc = smem + C::s_fc;
  const double* rn = smem + C::s_nrml;
  const double* fn = smem + C::s_fn;
  double* G = smem + C::s_grad;
  double* F = smem + C::s_flux;
  mbar_t* bar = reinterpret_cast<mbar_t*>(smem + C::smem_doubles);
  const int t = threadIdx.x;
  const int e = a.elem_begin + blockIdx.x;
  if (e >= a.elem_end) return;
  if (t == 0) { mbar_init(bar, 1); mbar_init_fence(); }
  __syncthreads();
  if (t == 0) {
    constexpr unsigned b_field = sizeof(double)*nq, b_face = sizeof(double)*2*ND*nv*nfq;
    mbar_arrive_expect_tx(bar, nv*b_field + 2*b_face + (DEF ? ND*ND*b_field + sizeof(double)*2*ND*ND*nfq : 0u));
    bulk_g2s(S, a.ed.state + (size_t)e*nv*nq, nv*b_field, bar);
    bulk_g2s(smem + C::s_ldg, a.faces_ldg + (size_t)e*2*ND*wl, b_face, bar);
    bulk_g2s(smem + C::s_fc, a.faces + (size_t)e*2*ND*wl, b_face, bar);
    if constexpr (DEF) {
      bulk_g2s(smem + C::s_nrml, a.refn + (size_t)(e - a.n_car)*ND*ND*nq, ND*ND*b_field, bar);
      bulk_g2s(smem + C::s_fn, a.normals + (size_t)(e - a.n_car)*2*ND*ND*nfq, sizeof(double)*2*ND*ND*nfq, bar);
    }
  }
  // per-point scalars: prefetch their cache lines now, read them in P2 / P4
  const double* g_tss = a.ed.tss + (size_t)e*nq;
  const double* g_av = a.ed.av + (size_t)e*2*nq;
  [[maybe_unused]] const double* g_det = DEF ? a.det + (size_t)(e - a.n_car)*nq : nullptr;
  {
    constexpr int lines = (nq*8 + 127)/128; // 128-byte lines per field
    for (int i = t; i < (DEF ? 4 : 3)*lines; i += T) {
      const int fld = i/lines, off = (i % lines)*16;
      prefetch_l1(fld == 0 ? g_tss + off : fld == 1 ? g_av + off : fld == 2 ? g_av + nq + off : g_det + off);
    }
  }
  const double nom = a.nom[e];
  const double inv_nom = 1./nom; // reciprocals instead of divisions, see g_local_kernel
  mbar_wait(bar, 0);

  /* ---- P1: gradient ---- */
  if constexpr (DEF) {
    const int l = t % nfq, j = t/nfq; // task of every sub-phase: line l of the current direction, physical component j
    #pragma unroll
    for (int d = 0; d < ND; ++d) {
      if (t < ND*nfq) {
        const int stride = d == 0 ? RS*RS : d == 1 ? RS : 1;
        const int q0 = d == 0 ? l : d == 1 ? (l/RS)*RS*RS + l % RS : l*RS;
        double nk[RS]; // the 1/nominal-size factor of the gradient (Spatial.hpp:398) is folded into the normals once per line
        #pragma unroll
        for (int k = 0; k < RS; ++k) nk[k] = rn[(d*ND + j)*nq + q0 + k*stride]*inv_nom;
        const double fn0 = fn[((2*d)*ND + j)*nfq + l]*inv_nom, fn1 = fn[((2*d + 1)*ND + j)*nfq + l]*inv_nom;
        #pragma unroll
        for (int v = 0; v < nv; ++v) {
          double p[RS];
          #pragma unroll
          for (int k = 0; k < RS; ++k) p[k] = nk[k]*S[v*nq + q0 + k*stride];
          const double b0 = fn0*fldg[((2*d)*nv + v)*nfq + l], b1 = fn1*fldg[((2*d + 1)*nv + v)*nfq + l];
          double r[RS];
          line_deriv_eo<RS>(ops, p, b0, b1, r);
          #pragma unroll
          for (int i = 0; i < RS; ++i) {
            double* g = G + (v*ND + j)*nq + q0 + i*stride;
            if (d == 0) *g = r[i]; else *g += r[i];
          }
        }
      }
      __syncthreads();
    }
  } else {
    if (t < C::n_line) {
      const int d = t/nfq, l = t % nfq;
      const int stride = d == 0 ? RS*RS : d == 1 ? RS : 1;
      const int q0 = d == 0 ? l : d == 1 ? (l/RS)*RS*RS + l % RS : l*RS;
      #pragma unroll
      for (int v = 0; v < nv; ++v) {
        double p[RS];
        #pragma unroll
        for (int k = 0; k < RS; ++k) p[k] = S[v*nq + q0 + k*stride];
        const double b0 = fldg[((2*d)*nv + v)*nfq + l], b1 = fldg[((2*d + 1)*nv + v)*nfq + l];
        double r[RS];
        line_deriv_eo<RS>(ops, p, b0, b1, r);
        #pragma unroll
        for (int i = 0; i < RS; ++i) G[(v*ND + d)*nq + q0 + i*stride] = r[i]*inv_nom;
      }
    }
    __syncthreads();
  }

  /* ---- P2: pointwise fluxes ---- */
  for (int q = t; q < nq; q += T) {
    typename P::template Comp<ND> comp;
    #pragma unroll
    for (int v = 0; v < nv; ++v) comp.state[v] = S[v*nq + q];
    comp.state[nv] = g_av[q];
    comp.state[nv + 1] = g_av[nq + q];
    if constexpr (DEF) {
      const double inv_det = 1./g_det[q];
      #pragma unroll
      for (int d = 0; d < ND; ++d)
        #pragma unroll
        for (int j = 0; j < ND; ++j) comp.normal[j][d] = rn[(d*ND + j)*nq + q];
      #pragma unroll
      for (int v = 0; v < nv; ++v)
        #pragma unroll
        for (int j = 0; j < ND; ++j) comp.gradient[v][j] = G[(v*ND + j)*nq + q]*inv_det;
    } else {
      #pragma unroll
      for (int v = 0; v < nv; ++v)
        #pragma unroll
        for (int j = 0; j < ND; ++j) comp.gradient[v][j] = G[(v*ND + j)*nq + q];
    }
    comp.compute_flux_conv(a.pp);
    comp.compute_flux_diff(a.pp);
    #pragma unroll
    for (int d = 0; d < ND; ++d)
      #pragma unroll
      for (int v = 0; v < nv; ++v) {
        F[(d*nv + v)*nq + q] = comp.flux_conv[v][d];
        G[(d*nv + v)*nq + q] = comp.flux_diff[v][d]; // all of this point's gradient entries are in registers by now
      }
  }
  __syncthreads();

  /* ---- P3: line derivatives in place, diffusive flux to the LDG faces ---- */
  if (t < C::n_line) {
    const int d = t/nfq, l = t % nfq;
    const int stride = d == 0 ? RS*RS : d == 1 ? RS : 1;
    const int q0 = d == 0 ? l : d == 1 ? (l/RS)*RS*RS + l % RS : l*RS;
    double* fl = a.faces_ldg + (size_t)e*2*ND*wl;
    #pragma unroll
    for (int v = 0; v < nv; ++v) {
      double* row = F + (d*nv + v)*nq + q0;
      double f[RS];
      #pragma unroll
      for (int k = 0; k < RS; ++k) f[k] = row[k*stride];
      const double b0 = fc[((2*d)*nv + v)*nfq + l], b1 = fc[((2*d + 1)*nv + v)*nfq + l];
      double r[RS];
      line_deriv_eo<RS, true>(ops, f, b0, b1, r);
      #pragma unroll
      for (int i = 0; i < RS; ++i) row[i*stride] = r[i];
      double* drow = G + (d*nv + v)*nq + q0;
      #pragma unroll
      for (int k = 0; k < RS; ++k) f[k] = drow[k*stride];
      double e0, e1;
      face_extrap_eo<RS>(ops, f, e0, e1);
      fl[(size_t)(2*d)*wl + v*nfq + l] = e0;
      fl[(size_t)(2*d + 1)*wl + v*nfq + l] = e1;
      line_diff_eo<RS, true>(ops, f, r);
      #pragma unroll
      for (int i = 0; i < RS; ++i) drow[i*stride] = r[i];
    }
  }
  __syncthreads();

  /* ---- P4: combine and update ---- */
  for (int q = t; q < nq; q += T) {
    double mult; // update*tss/nom/det with one division (<= 1 ulp)
    const double update = (a.dt_dev ? *a.dt_dev*a.update : a.update);
    if constexpr (DEF) mult = update*g_tss[q]/(nom*g_det[q]);
    else mult = update*g_tss[q]/nom;
    #pragma unroll
    for (int v = 0; v < nv; ++v) {
      double r0 = 0., r1 = 0.;
      #pragma unroll
      for (int d = 0; d < ND; ++d) { r0 += F[(d*nv + v)*nq + q]; r1 += G[(d*nv + v)*nq + q]; }
      double* cache = a.ed.cache + ((size_t)e*cs + v)*nq + q;
      double u = r0;
      *cache = u;
      u += r1;
      u *= mult;
      if (a.compute_residual) *cache = u;
      else a.ed.state[((size_t)e*nv + v)*nq + q] = S[v*nq + q] + u;
    }
  }
}

/* ---------------- Navier-Stokes Local, 3-D row size 6, bank-conflict-free layout ----------------
 * Same phases and the same arithmetic, operation for operation, as ns_local_line_kernel<6, DEF> (the two give bit-identical results), laid
 * out for the shared-memory pipe, which is what bounds that kernel: profiles/r01j_ncu_full_ns.md has it 82 % busy with 31 % of its
 * wavefronts (1 930 of 6 230 per element) caused by bank conflicts -- 8-byte accesses of consecutive lines of dimension 1 (stride 6) and
 * dimension 2 (stride 1) in a dense [6][6][6] field hit every bank twice. Here
 *   - every field in shared memory has a PLANE pitch of 38 doubles (field pitch 228): the six dimension-1 line starts of consecutive planes
 *     then fall 6 banks apart, so 16 consecutive lines of dimension 1 cover the 16 8-byte banks exactly once; dimension 0 (stride 38) keeps
 *     consecutive lanes on consecutive doubles. The bulk copies deliver a field plane by plane (30 + 54 copies of 288 B issued by the 32
 *     lanes of warp 0 instead of 2 copies by one thread);
 *   - dimension-2 lines are contiguous: 16-byte accesses, lines dealt to lanes so that every quarter-warp holds one line of each residue
 *     (x + y) mod 8, i.e. eight distinct 16-byte bank groups (c_ns_diag; a search also made the face points of a half-warp distinct);
 *   - tasks are dealt half-warp by half-warp (16 lines of one direction / one physical component), the four left-over lines of each group
 *     of 36 share the last half-warps;
 *   - point tasks take 16 consecutive points of ONE plane per half-warp (the 6 x 4 left-over points: two half-warps of 12 lanes whose
 *     plane offsets are 6 banks apart), so the 2-double gap between planes never falls inside a half-warp;
 *   - the face normals (two per task and direction) are read from HBM after an L2 prefetch instead of being staged, which pays for the
 *     padding: 72.5 KB per CTA as before, three resident CTAs. */
__constant__ unsigned char c_ns_diag[36] = {0, 29, 35, 8, 4, 15, 11, 17,  23, 34, 12, 18, 14, 25, 26, 22,  28, 1, 7, 3, 24, 20, 21, 27,
                                            33, 6, 2, 13, 9, 10, 31, 32,  19, 30, 5, 16};

/* PX = 38: the padded layout described above; PX = 36: the dense layout (one bulk copy per array, natural line and point order) with the same
 * task dealing, 16-byte accesses and early global loads -- kept to measure what the padding itself buys (HEXED_B200_OPT_NS_LOCAL_LAYOUT = 2) */
template <bool DEF, int PX_ = 38>
struct NsPadCfg
{
  static constexpr int RS = 6, ND = 3, nq = 216, nfq = 36, nv = 5, n_line = 108, threads = 128;
  static constexpr int PX = PX_, FP = 6*PX; // plane and field pitch in shared memory
  //   state | flux faces | region A = [LDG faces (padded to nv field pitches) | reference normals] | (free) | G ; F aliases region A as in NsCfg
  static constexpr int s_state = 0, s_fc = s_state + nv*FP, s_ldg = s_fc + 2*ND*nv*nfq, s_nrml = s_ldg + nv*FP;
  static constexpr int a_end = s_nrml + (DEF ? ND*ND*FP : 0);
  static constexpr int s_flux = s_nrml - nv*FP; // = s_ldg: F field 5 lands on the first reference normal
  static_assert(2*ND*nv*nfq <= nv*FP, "LDG faces must fit the first nv field pitches of F");
  static constexpr int s_grad = (a_end > s_flux + ND*nv*FP) ? a_end : s_flux + ND*nv*FP;
  static constexpr int smem_doubles = s_grad + ND*nv*FP;
  static constexpr size_t smem_bytes = sizeof(double)*smem_doubles + 2*sizeof(mbar_t);
};

/* line task (direction d, line l) of thread slot t for the phases whose tasks are lines of all three directions: warps 0 / 1 / 2 take 32
 * lines of direction 0 / 1 / 2 (warp 2 through c_ns_diag, with 16-byte accesses), warp 3 the left-over four lines of each */
template <int PX>
__device__ __forceinline__ int ns_pad_diag(int slot) { if constexpr (PX == 38) return c_ns_diag[slot]; else return slot; }

template <int PX>
__device__ __forceinline__ bool ns_pad_line(int t, int& d, int& l)
{
  if (t < 96) { d = t/32; l = d == 2 ? ns_pad_diag<PX>(t % 32) : t % 32; return true; }
  if (t < 100) { d = 0; l = 32 + (t - 96); return true; }
  if (t < 104) { d = 2; l = ns_pad_diag<PX>(32 + (t - 100)); return true; }
  if (t >= 112 && t < 116) { d = 1; l = 32 + (t - 112); return true; }
  d = 0; l = 0;
  return false;
}

/* point of slot s = t + 128*pass (see the header comment); -1: none */
template <int PX>
__device__ __forceinline__ int ns_pad_point(int s)
{
  if constexpr (PX != 38) return s < 216 ? s : -1;
  const int hw = s/16, lane = s % 16;
  if (hw < 12) return (hw/2)*36 + (hw % 2)*16 + lane;
  if (hw < 14 && lane < 12) return ((hw - 12)*3 + lane/4)*36 + 32 + lane % 4;
  return -1;
}

template <bool DEF, int PX_>
__global__ void __launch_bounds__(128, 3)
ns_local_pad_kernel(GArgs a, Ops ops)
{
  using C = NsPadCfg<DEF, PX_>;
  using P = PdeNs<3, 6, true>;
  constexpr int RS = 6, ND = 3, nq = C::nq, nfq = C::nfq, nv = C::nv, wl = nv*nfq, T = C::threads, PX = C::PX, FP = C::FP;
  constexpr int cs = nv > RS ? nv : RS;
  HB_DYN_SMEM(double, smem);
  double* S = smem + C::s_state;
  const double* fldg = smem + C::s_ldg;
  const double* fc = smem + C::s_fc;
  const double* rn = smem + C::s_nrml;
  double* G = smem + C::s_grad;
  double* F = smem + C::s_flux;
  mbar_t* bar = reinterpret_cast<mbar_t*>(smem + C::smem_doubles);
  const int t = threadIdx.x;
  const int e = a.elem_begin + blockIdx.x;
  if (e >= a.elem_end) return;
  constexpr unsigned b_plane = sizeof(double)*nfq, b_face = sizeof(double)*2*ND*nv*nfq;
  if (t == 0) { mbar_init(bar, 1); mbar_init_fence(); }
  __syncthreads();
  // (the one arrival that carries the byte count may come before or after the other lanes' copies: the phase cannot complete without it)
  if (t == 0) mbar_arrive_expect_tx(bar, nv*RS*b_plane + 2*b_face + (DEF ? ND*ND*RS*b_plane : 0u));
  if constexpr (PX == nfq) { // dense: one copy per array
    if (t == 0) bulk_g2s(S, a.ed.state + (size_t)e*nv*nq, nv*RS*b_plane, bar);
    if (t == 1) bulk_g2s(smem + C::s_ldg, a.faces_ldg + (size_t)e*2*ND*wl, b_face, bar);
    if (t == 2) bulk_g2s(smem + C::s_fc, a.faces + (size_t)e*2*ND*wl, b_face, bar);
    if constexpr (DEF) { if (t == 3) bulk_g2s(smem + C::s_nrml, a.refn + (size_t)(e - a.n_car)*ND*ND*nq, ND*ND*RS*b_plane, bar); }
  } else if (t < 32) { // one 288-byte copy per (field, plane) into the padded layout + the two face blocks
    const double* g_state = a.ed.state + (size_t)e*nv*nq;
    for (int c = t; c < nv*RS; c += 32) bulk_g2s(S + (c/RS)*FP + (c % RS)*PX, g_state + (size_t)c*nfq, b_plane, bar);
    if constexpr (DEF) {
      const double* g_rn = a.refn + (size_t)(e - a.n_car)*ND*ND*nq;
      for (int c = t; c < ND*ND*RS; c += 32) bulk_g2s(smem + C::s_nrml + (c/RS)*FP + (c % RS)*PX, g_rn + (size_t)c*nfq, b_plane, bar);
    }
    if (t == 30) bulk_g2s(smem + C::s_ldg, a.faces_ldg + (size_t)e*2*ND*wl, b_face, bar);
    if (t == 31) bulk_g2s(smem + C::s_fc, a.faces + (size_t)e*2*ND*wl, b_face, bar);
  }
  // per-point scalars and the face normals: prefetch their cache lines now, read them in P1 / P2 / P4
  const double* g_tss = a.ed.tss + (size_t)e*nq;
  const double* g_av = a.ed.av + (size_t)e*2*nq;
  [[maybe_unused]] const double* g_det = DEF ? a.det + (size_t)(e - a.n_car)*nq : nullptr;
  [[maybe_unused]] const double* g_fn = DEF ? a.normals + (size_t)(e - a.n_car)*2*ND*ND*nfq : nullptr;
  {
    constexpr int lines = (nq*8 + 127)/128, fn_lines = (2*ND*ND*nfq*8 + 127)/128 + 1; // 128-byte lines per field / of the face normals
    for (int i = t; i < (DEF ? 4*lines + fn_lines : 3*lines); i += T) {
      const int fld = i/lines, off = (i % lines)*16;
      if (fld < 4) prefetch_l1(fld == 0 ? g_tss + off : fld == 1 ? g_av + off : fld == 2 ? g_av + nq + off : g_det + off);
      else { const int o = (i - 4*lines)*16; prefetch_l1(g_fn + (o < 2*ND*ND*nfq ? o : 2*ND*ND*nfq - 1)); }
    }
  }
  const double nom = a.nom[e];
  const double inv_nom = 1./nom; // reciprocals instead of divisions, see g_local_kernel
  int ld, ll;
  const bool has_line = ns_pad_line<PX>(t, ld, ll);
  const bool vec = t >= 64 && t < 96; // warp 2: contiguous lines, 16-byte accesses
  const int lstride = ld == 0 ? PX : ld == 1 ? RS : 1;
  const int lq0 = ld == 0 ? ll : ld == 1 ? (ll/RS)*PX + ll % RS : (ll/RS)*PX + (ll % RS)*RS;
  // P1 task of every sub-phase (deformed): line l of the current direction, physical component j
  [[maybe_unused]] const bool active = t < 108;
  [[maybe_unused]] const int j = t < 96 ? t/32 : (t - 96)/4;
  [[maybe_unused]] const int slot = t < 96 ? t % 32 : 32 + (t - 96) % 4;
  [[maybe_unused]] double fnv[ND][2]; // the task's face normals, fetched from HBM while the bulk copies are in flight
  if constexpr (DEF) {
    if (active) {
      #pragma unroll
      for (int d = 0; d < ND; ++d) {
        const int l = d == 2 ? ns_pad_diag<PX>(slot) : slot;
        fnv[d][0] = g_fn[((2*d)*ND + j)*nfq + l];
        fnv[d][1] = g_fn[((2*d + 1)*ND + j)*nfq + l];
      }
    }
  }
  mbar_wait(bar, 0);

  /* ---- P1: gradient ---- */
  [[maybe_unused]] double pt_av0[2], pt_av1[2], pt_det[2], pt_tss[2]; // per-point scalars of P2 / P4, fetched one barrier early
  if constexpr (DEF) {
    #pragma unroll
    for (int d = 0; d < ND; ++d) {
      if (active) {
        const int l = d == 2 ? ns_pad_diag<PX>(slot) : slot;
        const int stride = d == 0 ? PX : d == 1 ? RS : 1;
        const int q0 = d == 0 ? l : d == 1 ? (l/RS)*PX + l % RS : (l/RS)*PX + (l % RS)*RS;
        double nk[RS]; // the 1/nominal-size factor of the gradient (Spatial.hpp:398) is folded into the normals once per line
        if (d == 2) {
          #pragma unroll
          for (int k = 0; k < RS; k += 2) ld2(rn + (d*ND + j)*FP + q0 + k, nk[k], nk[k + 1]);
        } else {
          #pragma unroll
          for (int k = 0; k < RS; ++k) nk[k] = rn[(d*ND + j)*FP + q0 + k*stride];
        }
        #pragma unroll
        for (int k = 0; k < RS; ++k) nk[k] *= inv_nom;
        const double fn0 = fnv[d][0]*inv_nom, fn1 = fnv[d][1]*inv_nom;
        #pragma unroll 1 // (code size: the unrolled kernel was 58 KB of SASS and lost 15 % of its issue slots to instruction fetch, profiles/r02k)
        for (int v = 0; v < nv; ++v) {
          double p[RS];
          if (d == 2) {
            #pragma unroll
            for (int k = 0; k < RS; k += 2) ld2(S + v*FP + q0 + k, p[k], p[k + 1]);
          } else {
            #pragma unroll
            for (int k = 0; k < RS; ++k) p[k] = S[v*FP + q0 + k*stride];
          }
          #pragma unroll
          for (int k = 0; k < RS; ++k) p[k] = nk[k]*p[k];
          const double b0 = fn0*fldg[((2*d)*nv + v)*nfq + l], b1 = fn1*fldg[((2*d + 1)*nv + v)*nfq + l];
          double r[RS];
          line_deriv_eo<RS>(ops, p, b0, b1, r);
          double* g = G + (v*ND + j)*FP + q0;
          if (d == 0) {
            #pragma unroll
            for (int i = 0; i < RS; ++i) g[i*stride] = r[i];
          } else if (d == 1) {
            #pragma unroll
            for (int i = 0; i < RS; ++i) g[i*stride] += r[i];
          } else {
            #pragma unroll
            for (int i = 0; i < RS; i += 2) { double g0, g1; ld2(g + i, g0, g1); st2(g + i, g0 + r[i], g1 + r[i + 1]); }
          }
        }
      }
      if (d < ND - 1) __syncthreads();
    }
  } else {
    if (has_line) {
      const int d = ld, l = ll, stride = lstride, q0 = lq0;
      #pragma unroll 1
      for (int v = 0; v < nv; ++v) {
        double p[RS];
        if (vec) {
          #pragma unroll
          for (int k = 0; k < RS; k += 2) ld2(S + v*FP + q0 + k, p[k], p[k + 1]);
        } else {
          #pragma unroll
          for (int k = 0; k < RS; ++k) p[k] = S[v*FP + q0 + k*stride];
        }
        const double b0 = fldg[((2*d)*nv + v)*nfq + l], b1 = fldg[((2*d + 1)*nv + v)*nfq + l];
        double r[RS];
        line_deriv_eo<RS>(ops, p, b0, b1, r);
        #pragma unroll
        for (int i = 0; i < RS; ++i) r[i] *= inv_nom;
        double* g = G + (v*ND + d)*FP + q0;
        if (vec) {
          #pragma unroll
          for (int i = 0; i < RS; i += 2) st2(g + i, r[i], r[i + 1]);
        } else {
          #pragma unroll
          for (int i = 0; i < RS; ++i) g[i*stride] = r[i];
        }
      }
    }
  }
  #pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
    const int q = ns_pad_point<PX>(t + pass*T);
    if (q >= 0) {
      pt_av0[pass] = g_av[q]; pt_av1[pass] = g_av[nq + q];
      if constexpr (DEF) pt_det[pass] = g_det[q];
    }
  }
  __syncthreads();

  /* ---- P2: pointwise fluxes ---- */
  #pragma unroll 1
  for (int pass = 0; pass < 2; ++pass) {
    const int q = ns_pad_point<PX>(t + pass*T);
    if (q >= 0) {
      const int qp = q + (PX - nfq)*(q/nfq);
      typename P::template Comp<ND> comp;
      #pragma unroll
      for (int v = 0; v < nv; ++v) comp.state[v] = S[v*FP + qp];
      comp.state[nv] = pass ? pt_av0[1] : pt_av0[0];
      comp.state[nv + 1] = pass ? pt_av1[1] : pt_av1[0];
      if constexpr (DEF) {
        const double inv_det = 1./(pass ? pt_det[1] : pt_det[0]);
        #pragma unroll
        for (int d = 0; d < ND; ++d)
          #pragma unroll
          for (int j = 0; j < ND; ++j) comp.normal[j][d] = rn[(d*ND + j)*FP + qp];
        #pragma unroll
        for (int v = 0; v < nv; ++v)
          #pragma unroll
          for (int j = 0; j < ND; ++j) comp.gradient[v][j] = G[(v*ND + j)*FP + qp]*inv_det;
      } else {
        #pragma unroll
        for (int v = 0; v < nv; ++v)
          #pragma unroll
          for (int j = 0; j < ND; ++j) comp.gradient[v][j] = G[(v*ND + j)*FP + qp];
      }
      comp.compute_flux_conv(a.pp);
      comp.compute_flux_diff(a.pp);
      #pragma unroll
      for (int d = 0; d < ND; ++d)
        #pragma unroll
        for (int v = 0; v < nv; ++v) {
          F[(d*nv + v)*FP + qp] = comp.flux_conv[v][d];
          G[(d*nv + v)*FP + qp] = comp.flux_diff[v][d]; // all of this point's gradient entries are in registers by now
        }
    }
  }
  __syncthreads();

  /* ---- P3: line derivatives in place, diffusive flux to the LDG faces ---- */
  if (has_line) {
    const int d = ld, l = ll, stride = lstride, q0 = lq0;
    double* fl = a.faces_ldg + (size_t)e*2*ND*wl;
    #pragma unroll 1
    for (int v = 0; v < nv; ++v) {
      double* row = F + (d*nv + v)*FP + q0;
      double f[RS], r[RS];
      if (vec) {
        #pragma unroll
        for (int k = 0; k < RS; k += 2) ld2(row + k, f[k], f[k + 1]);
      } else {
        #pragma unroll
        for (int k = 0; k < RS; ++k) f[k] = row[k*stride];
      }
      const double b0 = fc[((2*d)*nv + v)*nfq + l], b1 = fc[((2*d + 1)*nv + v)*nfq + l];
      line_deriv_eo<RS, true>(ops, f, b0, b1, r);
      if (vec) {
        #pragma unroll
        for (int i = 0; i < RS; i += 2) st2(row + i, r[i], r[i + 1]);
      } else {
        #pragma unroll
        for (int i = 0; i < RS; ++i) row[i*stride] = r[i];
      }
      double* drow = G + (d*nv + v)*FP + q0;
      if (vec) {
        #pragma unroll
        for (int k = 0; k < RS; k += 2) ld2(drow + k, f[k], f[k + 1]);
      } else {
        #pragma unroll
        for (int k = 0; k < RS; ++k) f[k] = drow[k*stride];
      }
      double e0, e1;
      face_extrap_eo<RS>(ops, f, e0, e1);
      fl[(size_t)(2*d)*wl + v*nfq + l] = e0;
      fl[(size_t)(2*d + 1)*wl + v*nfq + l] = e1;
      line_diff_eo<RS, true>(ops, f, r);
      if (vec) {
        #pragma unroll
        for (int i = 0; i < RS; i += 2) st2(drow + i, r[i], r[i + 1]);
      } else {
        #pragma unroll
        for (int i = 0; i < RS; ++i) drow[i*stride] = r[i];
      }
    }
  }
  #pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
    const int q = ns_pad_point<PX>(t + pass*T);
    if (q >= 0) pt_tss[pass] = g_tss[q];
  }
  const double update = (a.dt_dev ? *a.dt_dev*a.update : a.update);
  __syncthreads();

  /* ---- P4: combine and update ---- */
  #pragma unroll 1
  for (int pass = 0; pass < 2; ++pass) {
    const int q = ns_pad_point<PX>(t + pass*T);
    if (q >= 0) {
      const int qp = q + (PX - nfq)*(q/nfq);
      const double tss_q = pass ? pt_tss[1] : pt_tss[0];
      double mult; // update*tss/nom/det with one division (<= 1 ulp)
      if constexpr (DEF) mult = update*tss_q/(nom*(pass ? pt_det[1] : pt_det[0]));
      else mult = update*tss_q/nom;
      #pragma unroll
      for (int v = 0; v < nv; ++v) {
        double r0 = 0., r1 = 0.;
        #pragma unroll
        for (int d = 0; d < ND; ++d) { r0 += F[(d*nv + v)*FP + qp]; r1 += G[(d*nv + v)*FP + qp]; }
        double* cache = a.ed.cache + ((size_t)e*cs + v)*nq + q;
        double u = r0;
        *cache = u;
        u += r1;
        u *= mult;
        if (a.compute_residual) *cache = u;
        else a.ed.state[((size_t)e*nv + v)*nq + q] = S[v*FP + qp] + u;
      }
    }
  }
}

/* ---------------- Navier-Stokes Reconcile_ldg_flux, 3-D, bulk-copy staging ----------------
 * Same arithmetic as g_reconcile_kernel<3, RS, PdeNs<3, RS, true>, DEF> without the modal filter and outside residual mode
 * (reference include/Spatial.hpp:543-594). The kernel is a pure stream (38 KB in+out per element at row size 6 against ~60 flops
 * per point); the generic version reached 2.85 TB/s because each CTA walks load -> barrier -> load -> store -> barrier -> store with
 * per-thread global loads (profiles/r01g_ncu_full_ns.md: long-scoreboard 19.7 cycles per issue). Here the element's state, LDG
 * faces, time-step scale and determinant arrive as four 1-D bulk TMA copies on one mbarrier and ~9 one-shot CTAs per SM keep
 * enough bytes in flight; the new state is written from registers, the faces from the updated shared-memory copy. */
template <int RS, bool DEF>
struct NsRecCfg
{
  static constexpr int ND = 3, nq = RS*RS*RS, nfq = RS*RS, nv = 5;
  static constexpr int threads = ((nq + 31)/32)*32 > 256 ? 256 : ((nq + 31)/32)*32;
  static constexpr int s_state = 0, s_ldg = s_state + nv*nq, s_tss = s_ldg + 2*ND*nv*nfq, s_det = s_tss + nq;
  static constexpr int smem_doubles = s_det + (DEF ? nq : 0);
  static constexpr size_t smem_bytes = sizeof(double)*smem_doubles + 2*sizeof(mbar_t);
};

template <int RS, bool DEF>
__global__ void __launch_bounds__(NsRecCfg<RS, DEF>::threads)
ns_reconcile_bulk_kernel(GArgs a, Ops ops)
{
  using C = NsRecCfg<RS, DEF>;
  constexpr int ND = 3, nq = C::nq, nfq = C::nfq, nv = C::nv, wl = nv*nfq, T = C::threads;
  HB_DYN_SMEM(double, smem);
  double* S = smem + C::s_state;
  const double* fldg = smem + C::s_ldg;
  const double* s_tss = smem + C::s_tss;
  const double* s_det = smem + C::s_det;
  mbar_t* bar = reinterpret_cast<mbar_t*>(smem + C::smem_doubles);
  const int t = threadIdx.x;
  const int e = a.elem_begin + blockIdx.x;
  if (e >= a.elem_end) return;
  if (t == 0) { mbar_init(bar, 1); mbar_init_fence(); }
  __syncthreads();
  if (t == 0) {
    constexpr unsigned b_field = sizeof(double)*nq, b_face = sizeof(double)*2*ND*nv*nfq;
    mbar_arrive_expect_tx(bar, nv*b_field + b_face + b_field + (DEF ? b_field : 0u));
    bulk_g2s(smem + C::s_ldg, a.faces_ldg + (size_t)e*2*ND*wl, b_face, bar);
    bulk_g2s(S, a.ed.state + (size_t)e*nv*nq, nv*b_field, bar);
    bulk_g2s(smem + C::s_tss, a.ed.tss + (size_t)e*nq, b_field, bar);
    if constexpr (DEF) bulk_g2s(smem + C::s_det, a.det + (size_t)(e - a.n_car)*nq, b_field, bar);
  }
  const double nom = a.nom[e];
  mbar_wait(bar, 0);
  for (int q = t; q < nq; q += T) {
    double r[nv];
    #pragma unroll
    for (int v = 0; v < nv; ++v) r[v] = 0.;
    #pragma unroll
    for (int d = 0; d < ND; ++d) {
      const int stride = d == 0 ? RS*RS : d == 1 ? RS : 1;
      const int node = (q/stride) % RS;
      const int fq = (q/(stride*RS))*stride + q % stride;
      const double l0 = ops.lift[node][0], l1 = ops.lift[node][1];
      #pragma unroll
      for (int v = 0; v < nv; ++v) {
        double acc = 0;
        acc += l0*fldg[((2*d)*nv + v)*nfq + fq];
        acc += l1*fldg[((2*d + 1)*nv + v)*nfq + fq];
        r[v] -= acc;
      }
    }
    double mult = (a.dt_dev ? *a.dt_dev*a.update : a.update)*s_tss[q]/nom;
    if constexpr (DEF) mult /= s_det[q];
    #pragma unroll
    for (int v = 0; v < nv; ++v) {
      const double u = S[v*nq + q] + r[v]*mult;
      S[v*nq + q] = u;
      a.ed.state[((size_t)e*nv + v)*nq + q] = u;
    }
  }
  __syncthreads();
  double* dst = a.faces + (size_t)e*2*ND*wl;
  for (int item = t; item < ND*nv*nfq; item += T) {
    const int d = item/(nv*nfq), v = (item/nfq) % nv, fq = item % nfq;
    const int stride = d == 0 ? RS*RS : d == 1 ? RS : 1;
    const int base = (fq/stride)*stride*RS + fq % stride;
    double e0 = 0, e1 = 0;
    #pragma unroll
    for (int k = 0; k < RS; ++k) {
      const double x = S[v*nq + base + k*stride];
      e0 += ops.bnd[0][k]*x;
      e1 += ops.bnd[1][k]*x;
    }
    dst[(size_t)(2*d)*wl + v*nfq + fq] = e0;
    dst[(size_t)(2*d + 1)*wl + v*nfq + fq] = e1;
  }
}

/* 2-D version: a batch of B consecutive elements per CTA (their state / LDG faces / tss / determinant are four contiguous runs) */
template <int RS, bool DEF>
struct NsRec2Cfg
{
  static constexpr int ND = 2, nq = RS*RS, nfq = RS, nv = 4, B = 16, threads = 256;
  static constexpr int s_state = 0, s_ldg = s_state + B*nv*nq, s_tss = s_ldg + B*2*ND*nv*nfq, s_det = s_tss + B*nq;
  static constexpr int smem_doubles = s_det + (DEF ? B*nq : 0);
  static constexpr size_t smem_bytes = sizeof(double)*smem_doubles + 2*sizeof(mbar_t);
};

template <int RS, bool DEF>
__global__ void __launch_bounds__(NsRec2Cfg<RS, DEF>::threads)
ns_reconcile_bulk2d_kernel(GArgs a, Ops ops)
{
  using C = NsRec2Cfg<RS, DEF>;
  constexpr int ND = 2, nq = C::nq, nfq = C::nfq, nv = C::nv, wl = nv*nfq, T = C::threads, B = C::B;
  HB_DYN_SMEM(double, smem);
  double* S = smem + C::s_state;
  const double* fldg = smem + C::s_ldg;
  const double* s_tss = smem + C::s_tss;
  const double* s_det = smem + C::s_det;
  mbar_t* bar = reinterpret_cast<mbar_t*>(smem + C::smem_doubles);
  const int t = threadIdx.x;
  const int e0 = a.elem_begin + blockIdx.x*B;
  if (e0 >= a.elem_end) return;
  const int n = a.elem_end - e0 < B ? a.elem_end - e0 : B;
  if (t == 0) { mbar_init(bar, 1); mbar_init_fence(); }
  __syncthreads();
  if (t == 0) {
    const unsigned b_field = sizeof(double)*nq*n, b_face = sizeof(double)*2*ND*nv*nfq*n;
    mbar_arrive_expect_tx(bar, nv*b_field + b_face + b_field + (DEF ? b_field : 0u));
    bulk_g2s(smem + C::s_ldg, a.faces_ldg + (size_t)e0*2*ND*wl, b_face, bar);
    bulk_g2s(S, a.ed.state + (size_t)e0*nv*nq, nv*b_field, bar);
    bulk_g2s(smem + C::s_tss, a.ed.tss + (size_t)e0*nq, b_field, bar);
    if constexpr (DEF) bulk_g2s(smem + C::s_det, a.det + (size_t)(e0 - a.n_car)*nq, b_field, bar);
  }
  mbar_wait(bar, 0);
  for (int pt = t; pt < n*nq; pt += T) {
    const int pe = pt/nq, q = pt % nq;
    double r[nv];
    #pragma unroll
    for (int v = 0; v < nv; ++v) r[v] = 0.;
    #pragma unroll
    for (int d = 0; d < ND; ++d) {
      const int stride = d == 0 ? RS : 1;
      const int node = (q/stride) % RS;
      const int fq = (q/(stride*RS))*stride + q % stride;
      const double l0 = ops.lift[node][0], l1 = ops.lift[node][1];
      #pragma unroll
      for (int v = 0; v < nv; ++v) {
        double acc = 0;
        acc += l0*fldg[(pe*2*ND + 2*d)*wl + v*nfq + fq];
        acc += l1*fldg[(pe*2*ND + 2*d + 1)*wl + v*nfq + fq];
        r[v] -= acc;
      }
    }
    double mult = (a.dt_dev ? *a.dt_dev*a.update : a.update)*s_tss[pt]/a.nom[e0 + pe];
    if constexpr (DEF) mult /= s_det[pt];
    #pragma unroll
    for (int v = 0; v < nv; ++v) {
      const double u = S[(pe*nv + v)*nq + q] + r[v]*mult;
      S[(pe*nv + v)*nq + q] = u;
      a.ed.state[((size_t)(e0 + pe)*nv + v)*nq + q] = u;
    }
  }
  __syncthreads();
  for (int item = t; item < n*ND*nv*nfq; item += T) {
    const int pe = item/(ND*nv*nfq), rest = item % (ND*nv*nfq);
    const int d = rest/(nv*nfq), v = (rest/nfq) % nv, fq = rest % nfq;
    const int stride = d == 0 ? RS : 1;
    const int base = (fq/stride)*stride*RS + fq % stride;
    double x0 = 0, x1 = 0;
    #pragma unroll
    for (int k = 0; k < RS; ++k) {
      const double x = S[(pe*nv + v)*nq + base + k*stride];
      x0 += ops.bnd[0][k]*x;
      x1 += ops.bnd[1][k]*x;
    }
    double* dst = a.faces + (size_t)(e0 + pe)*2*ND*wl;
    dst[(size_t)(2*d)*wl + v*nfq + fq] = x0;
    dst[(size_t)(2*d + 1)*wl + v*nfq + fq] = x1;
  }
}

/* returns -1 when the combination is not covered and the caller should use g_reconcile_kernel */
template <int ND, int RS>
int launch_ns_reconcile_bulk(hexed_b200_ctx* c, const GArgs& a, int deformed)
{
  if constexpr (ND == 2 && (RS == 2 || RS == 4 || RS == 6 || RS == 8)) {
    if (a.use_filter || a.compute_residual || !c->use_pipe) return -1;
    using C0 = NsRec2Cfg<RS, false>;
    const int grid = (a.elem_end - a.elem_begin + C0::B - 1)/C0::B;
    if (deformed) { using C = NsRec2Cfg<RS, true>; auto k = ns_reconcile_bulk2d_kernel<RS, true>; int r = set_smem(c, k, C::smem_bytes); if (r) return r; HB_LAUNCH(k, grid, C::threads, C::smem_bytes, c->stream, a, c->ops); }
    else { using C = NsRec2Cfg<RS, false>; auto k = ns_reconcile_bulk2d_kernel<RS, false>; int r = set_smem(c, k, C::smem_bytes); if (r) return r; HB_LAUNCH(k, grid, C::threads, C::smem_bytes, c->stream, a, c->ops); }
    return 0;
  } else if constexpr (ND == 3 && (RS == 2 || RS == 4 || RS == 6)) { // bulk copies need 16-byte multiples: nq*8 with even row size
    if (a.use_filter || a.compute_residual || !c->use_pipe) return -1;
    const int grid = a.elem_end - a.elem_begin;
    if (deformed) { using C = NsRecCfg<RS, true>; auto k = ns_reconcile_bulk_kernel<RS, true>; int r = set_smem(c, k, C::smem_bytes); if (r) return r; HB_LAUNCH(k, grid, C::threads, C::smem_bytes, c->stream, a, c->ops); }
    else { using C = NsRecCfg<RS, false>; auto k = ns_reconcile_bulk_kernel<RS, false>; int r = set_smem(c, k, C::smem_bytes); if (r) return r; HB_LAUNCH(k, grid, C::threads, C::smem_bytes, c->stream, a, c->ops); }
    return 0;
  } else {
    return -1;
  }
}

/* ---------------- Navier-Stokes Local, 2-D, batched line-task formulation ----------------
 * The cylinder-class configurations (samples/cylinder, BASELINE config 3) are 2-D: an element is only row_size^2 points, so one CTA
 * takes a BATCH of B consecutive elements and runs the phases of ns_local_line_kernel over all of them at once: (element, line,
 * component) gradient tasks per direction sub-phase, pointwise fluxes, (element, direction, line) derivative tasks, pointwise update.
 * Same arithmetic as g_local_kernel<2, RS, PdeNs<2, RS, true>, DEF> without the modal filter (reference include/Spatial.hpp:326-509).
 * Per element in shared memory: state | flux faces | region A = [LDG faces | reference normals | face normals] | G, fetched by
 * per-element 1-D bulk TMA copies (thread i issues the copies of element i) onto one mbarrier; F aliases region A starting two
 * fields before the normals, so that its fields 2..5 coincide with the four reference normals (see NsCfg). 6.7 KB per element at
 * row size 6 -> B = 10, 67 KB per CTA, three CTAs per SM. The generic point-per-thread kernel it replaces took 5.3 ms per launch at
 * 1 M elements, 3.5x the Euler Local kernel. */
template <int RS, bool DEF>
struct Ns2Cfg
{
  static constexpr int ND = 2, nq = RS*RS, nfq = RS, nv = 4, lines = ND*nfq;
  static constexpr int B = 128/lines; // elements per CTA
  static constexpr int threads = 128;
  static constexpr int s_state = 0, s_fc = s_state + nv*nq, s_ldg = s_fc + 2*ND*nv*nfq;
  static constexpr int s_nrml = s_ldg + 2*ND*nv*nfq, s_fn = s_nrml + (DEF ? ND*ND*nq : 0);
  static constexpr int a_end = s_fn + (DEF ? 2*ND*ND*nfq : 0);
  static constexpr int lead = (2*ND*nv*nfq)/nq; // whole fields that fit in the LDG faces in front of the normals
  static constexpr int s_flux = s_nrml - lead*nq;
  static constexpr int s_grad = (a_end > s_flux + ND*nv*nq) ? a_end : s_flux + ND*nv*nq;
  static constexpr int elem_doubles = s_grad + ND*nv*nq;
  static_assert(elem_doubles % 2 == 0 && s_fc % 2 == 0 && s_ldg % 2 == 0 && s_nrml % 2 == 0 && s_fn % 2 == 0, "16-byte alignment of the bulk copies");
  static constexpr size_t smem_bytes = sizeof(double)*B*elem_doubles + 2*sizeof(mbar_t);
};

template <int RS, bool DEF>
__global__ void __launch_bounds__(Ns2Cfg<RS, DEF>::threads, 3)
ns_local_line2d_kernel(GArgs a, Ops ops)
{
  using C = Ns2Cfg<RS, DEF>;
  using P = PdeNs<2, RS, true>;
  constexpr int ND = 2, nq = C::nq, nfq = C::nfq, nv = C::nv, wl = nv*nfq, T = C::threads, B = C::B;
  constexpr int cs = nv > RS ? nv : RS;
  HB_DYN_SMEM(double, smem);
  mbar_t* bar = reinterpret_cast<mbar_t*>(smem + B*C::elem_doubles);
  const int t = threadIdx.x;
  const int e0 = a.elem_begin + blockIdx.x*B;
  if (e0 >= a.elem_end) return;
  const int n = a.elem_end - e0 < B ? a.elem_end - e0 : B; // elements of this batch
  if (t == 0) { mbar_init(bar, 1); mbar_init_fence(); }
  __syncthreads();
  constexpr unsigned b_field = sizeof(double)*nq, b_face = sizeof(double)*2*ND*nv*nfq;
  constexpr unsigned b_elem = nv*b_field + 2*b_face + (DEF ? ND*ND*b_field + (unsigned)sizeof(double)*2*ND*ND*nfq : 0u);
  if (t == 0) mbar_arrive_expect_tx(bar, b_elem*n);
  __syncthreads(); // the expectation is posted before any copy can complete
  if (t < n) {
    const int e = e0 + t;
    double* base = smem + t*C::elem_doubles;
    bulk_g2s(base + C::s_state, a.ed.state + (size_t)e*nv*nq, nv*b_field, bar);
    bulk_g2s(base + C::s_ldg, a.faces_ldg + (size_t)e*2*ND*wl, b_face, bar);
    bulk_g2s(base + C::s_fc, a.faces + (size_t)e*2*ND*wl, b_face, bar);
    if constexpr (DEF) {
      bulk_g2s(base + C::s_nrml, a.refn + (size_t)(e - a.n_car)*ND*ND*nq, ND*ND*b_field, bar);
      bulk_g2s(base + C::s_fn, a.normals + (size_t)(e - a.n_car)*2*ND*ND*nfq, sizeof(double)*2*ND*ND*nfq, bar);
    }
  }
  // per-point scalars of the batch are contiguous runs in HBM: prefetch, read in P2 / P4
  const double* g_tss = a.ed.tss + (size_t)e0*nq;
  const double* g_av = a.ed.av + (size_t)e0*2*nq;
  [[maybe_unused]] const double* g_det = DEF ? a.det + (size_t)(e0 - a.n_car)*nq : nullptr;
  for (int i = t*16; i < n*nq; i += T*16) { prefetch_l1(g_tss + i); if constexpr (DEF) prefetch_l1(g_det + i); }
  for (int i = t*16; i < n*2*nq; i += T*16) prefetch_l1(g_av + i);
  mbar_wait(bar, 0);

  /* ---- P1: gradient ---- */
  if constexpr (DEF) {
    const int le = t/(nfq*ND), j = (t/nfq) % ND, l = t % nfq; // element of the batch, physical component, line
    const bool on = t < B*nfq*ND && le < n;
    double* eb = smem + le*C::elem_doubles;
    const double inv_nom = on ? 1./a.nom[e0 + le] : 0.;
    auto sub_phase = [&](auto dc) {
      constexpr int d = decltype(dc)::value;
      if (on) {
        constexpr int stride = d == 0 ? RS : 1;
        const int q0 = d == 0 ? l : l*RS;
        const double* S = eb + C::s_state; const double* rn = eb + C::s_nrml; const double* fn = eb + C::s_fn; const double* fldg = eb + C::s_ldg;
        double* G = eb + C::s_grad;
        double nk[RS];
        #pragma unroll
        for (int k = 0; k < RS; ++k) nk[k] = rn[(d*ND + j)*nq + q0 + k*stride]*inv_nom;
        const double fn0 = fn[((2*d)*ND + j)*nfq + l]*inv_nom, fn1 = fn[((2*d + 1)*ND + j)*nfq + l]*inv_nom;
        #pragma unroll
        for (int v = 0; v < nv; ++v) {
          double p[RS];
          #pragma unroll
          for (int k = 0; k < RS; ++k) p[k] = nk[k]*S[v*nq + q0 + k*stride];
          const double b0 = fn0*fldg[((2*d)*nv + v)*nfq + l], b1 = fn1*fldg[((2*d + 1)*nv + v)*nfq + l];
          double r[RS];
          line_deriv_eo<RS>(ops, p, b0, b1, r);
          #pragma unroll
          for (int i = 0; i < RS; ++i) {
            double* g = G + (v*ND + j)*nq + q0 + i*stride;
            if (d == 0) *g = r[i]; else *g += r[i];
          }
        }
      }
      __syncthreads();
    };
    sub_phase(IC<0>{}); sub_phase(IC<1>{});
  } else {
    const int le = t/C::lines, d = (t % C::lines)/nfq, l = t % nfq;
    if (t < B*C::lines && le < n) {
      double* eb = smem + le*C::elem_doubles;
      const double inv_nom = 1./a.nom[e0 + le];
      const int stride = d == 0 ? RS : 1;
      const int q0 = d == 0 ? l : l*RS;
      const double* S = eb + C::s_state; const double* fldg = eb + C::s_ldg;
      double* G = eb + C::s_grad;
      #pragma unroll
      for (int v = 0; v < nv; ++v) {
        double p[RS];
        #pragma unroll
        for (int k = 0; k < RS; ++k) p[k] = S[v*nq + q0 + k*stride];
        const double b0 = fldg[((2*d)*nv + v)*nfq + l], b1 = fldg[((2*d + 1)*nv + v)*nfq + l];
        double r[RS];
        line_deriv_eo<RS>(ops, p, b0, b1, r);
        #pragma unroll
        for (int i = 0; i < RS; ++i) G[(v*ND + d)*nq + q0 + i*stride] = r[i]*inv_nom;
      }
    }
    __syncthreads();
  }

  /* ---- P2: pointwise fluxes ---- */
  for (int pt = t; pt < n*nq; pt += T) {
    const int pe = pt/nq, q = pt % nq;
    double* eb = smem + pe*C::elem_doubles;
    const double* S = eb + C::s_state; [[maybe_unused]] const double* rn = eb + C::s_nrml;
    double* G = eb + C::s_grad; double* F = eb + C::s_flux;
    typename P::template Comp<ND> comp;
    #pragma unroll
    for (int v = 0; v < nv; ++v) comp.state[v] = S[v*nq + q];
    comp.state[nv] = g_av[(size_t)pe*2*nq + q];
    comp.state[nv + 1] = g_av[(size_t)pe*2*nq + nq + q];
    if constexpr (DEF) {
      const double inv_det = 1./g_det[pt];
      #pragma unroll
      for (int d = 0; d < ND; ++d)
        #pragma unroll
        for (int j = 0; j < ND; ++j) comp.normal[j][d] = rn[(d*ND + j)*nq + q];
      #pragma unroll
      for (int v = 0; v < nv; ++v)
        #pragma unroll
        for (int j = 0; j < ND; ++j) comp.gradient[v][j] = G[(v*ND + j)*nq + q]*inv_det;
    } else {
      #pragma unroll
      for (int v = 0; v < nv; ++v)
        #pragma unroll
        for (int j = 0; j < ND; ++j) comp.gradient[v][j] = G[(v*ND + j)*nq + q];
    }
    comp.compute_flux_conv(a.pp);
    comp.compute_flux_diff(a.pp);
    #pragma unroll
    for (int d = 0; d < ND; ++d)
      #pragma unroll
      for (int v = 0; v < nv; ++v) {
        F[(d*nv + v)*nq + q] = comp.flux_conv[v][d]; // fields 2..5 overwrite this point's own normals, read above
        G[(d*nv + v)*nq + q] = comp.flux_diff[v][d];
      }
  }
  __syncthreads();

  /* ---- P3: line derivatives in place, diffusive flux to the LDG faces ---- */
  {
    const int le = t/C::lines, d = (t % C::lines)/nfq, l = t % nfq;
    if (t < B*C::lines && le < n) {
      double* eb = smem + le*C::elem_doubles;
      const double* fc = eb + C::s_fc;
      double* G = eb + C::s_grad; double* F = eb + C::s_flux;
      const int stride = d == 0 ? RS : 1;
      const int q0 = d == 0 ? l : l*RS;
      double* fl = a.faces_ldg + (size_t)(e0 + le)*2*ND*wl;
      #pragma unroll
      for (int v = 0; v < nv; ++v) {
        double* row = F + (d*nv + v)*nq + q0;
        double f[RS];
        #pragma unroll
        for (int k = 0; k < RS; ++k) f[k] = row[k*stride];
        const double b0 = fc[((2*d)*nv + v)*nfq + l], b1 = fc[((2*d + 1)*nv + v)*nfq + l];
        double r[RS];
        line_deriv_eo<RS, true>(ops, f, b0, b1, r);
        #pragma unroll
        for (int i = 0; i < RS; ++i) row[i*stride] = r[i];
        double* drow = G + (d*nv + v)*nq + q0;
        #pragma unroll
        for (int k = 0; k < RS; ++k) f[k] = drow[k*stride];
        double x0, x1;
        face_extrap_eo<RS>(ops, f, x0, x1);
        fl[(size_t)(2*d)*wl + v*nfq + l] = x0;
        fl[(size_t)(2*d + 1)*wl + v*nfq + l] = x1;
        line_diff_eo<RS, true>(ops, f, r);
        #pragma unroll
        for (int i = 0; i < RS; ++i) drow[i*stride] = r[i];
      }
    }
  }
  __syncthreads();

  /* ---- P4: combine and update ---- */
  for (int pt = t; pt < n*nq; pt += T) {
    const int pe = pt/nq, q = pt % nq;
    const int e = e0 + pe;
    const double* eb = smem + pe*C::elem_doubles;
    const double* S = eb + C::s_state; const double* G = eb + C::s_grad; const double* F = eb + C::s_flux;
    const double nom = a.nom[e];
    double mult; // update*tss/nom/det with one division (<= 1 ulp)
    const double update = (a.dt_dev ? *a.dt_dev*a.update : a.update);
    if constexpr (DEF) mult = update*g_tss[pt]/(nom*g_det[pt]);
    else mult = update*g_tss[pt]/nom;
    #pragma unroll
    for (int v = 0; v < nv; ++v) {
      double r0 = 0., r1 = 0.;
      #pragma unroll
      for (int d = 0; d < ND; ++d) { r0 += F[(d*nv + v)*nq + q]; r1 += G[(d*nv + v)*nq + q]; }
      double* cache = a.ed.cache + ((size_t)e*cs + v)*nq + q;
      double u = r0;
      *cache = u;
      u += r1;
      u *= mult;
      if (a.compute_residual) *cache = u;
      else a.ed.state[((size_t)e*nv + v)*nq + q] = S[v*nq + q] + u;
    }
  }
}

/* returns -1 when the combination is not covered and the caller should use g_local_kernel */
template <int ND, int RS>
int launch_ns_local_line(hexed_b200_ctx* c, const GArgs& a, int deformed)
{
  if constexpr (ND == 2 && (RS == 4 || RS == 6 || RS == 8)) {
    if (a.use_filter || !c->use_pipe || !c->ops_symmetric) return -1;
    using C0 = Ns2Cfg<RS, false>;
    const int grid = (a.elem_end - a.elem_begin + C0::B - 1)/C0::B;
    if (deformed) { using C = Ns2Cfg<RS, true>; auto k = ns_local_line2d_kernel<RS, true>; int r = set_smem(c, k, C::smem_bytes); if (r) return r; HB_LAUNCH(k, grid, C::threads, C::smem_bytes, c->stream, a, c->ops); }
    else { using C = Ns2Cfg<RS, false>; auto k = ns_local_line2d_kernel<RS, false>; int r = set_smem(c, k, C::smem_bytes); if (r) return r; HB_LAUNCH(k, grid, C::threads, C::smem_bytes, c->stream, a, c->ops); }
    return 0;
  } else if constexpr (ND == 3 && (RS == 4 || RS == 6)) {
    if (a.use_filter || !c->use_pipe || !c->ops_symmetric) return -1;
    const int grid = a.elem_end - a.elem_begin;
    if constexpr (RS == 6) {
      if (c->ns_layout == 1) { // the padded, bank-conflict-free layout
        if (deformed) { using C = NsPadCfg<true, 38>; auto k = ns_local_pad_kernel<true, 38>; int r = set_smem(c, k, C::smem_bytes); if (r) return r; HB_LAUNCH(k, grid, C::threads, C::smem_bytes, c->stream, a, c->ops); }
        else { using C = NsPadCfg<false, 38>; auto k = ns_local_pad_kernel<false, 38>; int r = set_smem(c, k, C::smem_bytes); if (r) return r; HB_LAUNCH(k, grid, C::threads, C::smem_bytes, c->stream, a, c->ops); }
        return 0;
      }
      if (c->ns_layout == 2) { // dense layout, half-warp-aligned tasks, 16-byte accesses on contiguous lines
        if (deformed) { using C = NsPadCfg<true, 36>; auto k = ns_local_pad_kernel<true, 36>; int r = set_smem(c, k, C::smem_bytes); if (r) return r; HB_LAUNCH(k, grid, C::threads, C::smem_bytes, c->stream, a, c->ops); }
        else { using C = NsPadCfg<false, 36>; auto k = ns_local_pad_kernel<false, 36>; int r = set_smem(c, k, C::smem_bytes); if (r) return r; HB_LAUNCH(k, grid, C::threads, C::smem_bytes, c->stream, a, c->ops); }
        return 0;
      }
    }
    if (deformed) { using C = NsCfg<RS, true>; auto k = ns_local_line_kernel<RS, true>; int r = set_smem(c, k, C::smem_bytes); if (r) return r; HB_LAUNCH(k, grid, C::threads, C::smem_bytes, c->stream, a, c->ops); }
    else { using C = NsCfg<RS, false>; auto k = ns_local_line_kernel<RS, false>; int r = set_smem(c, k, C::smem_bytes); if (r) return r; HB_LAUNCH(k, grid, C::threads, C::smem_bytes, c->stream, a, c->ops); }
    return 0;
  } else {
    return -1;
  }
}
#endif

/* ---------------- Reconcile_ldg_flux ---------------- */
template <int ND, int RS, class P, bool DEF>
__global__ void __launch_bounds__(GCfg<ND, RS, P>::threads)
g_reconcile_kernel(GArgs a, Ops ops, FilterOp filt)
{
  using C = GCfg<ND, RS, P>;
  constexpr int nq = C::nq, nfq = C::nfq, ne = C::ne, nu = C::nu, nv = ND + 2, wl = nv*nfq;
  constexpr int cache_slot0 = nv + 7 + RS;
  HB_DYN_SMEM(double, smem);
  const int t = threadIdx.x;
  const int le = t/nq, q = t % nq;
  const int e = a.elem_begin + blockIdx.x*C::epb + le;
  const bool active = e < a.elem_end;
  double* s_filt = smem + 2*RS*RS; double* s_lift = smem + 3*RS*RS;
  double* sbuf = smem + C::nops + le*C::rec_per_elem;
  double* svface = sbuf + C::rec_buf;
  load_ops<ND, RS, P>(smem, ops, filt, t, C::threads);
  if (active) {
    // the per-point inputs of the update are needed only after the face terms: start fetching them now (the first profile of this
    // kernel showed 20 cycles of long-scoreboard stall per issue on exactly these loads, profiles/r01g_ncu_full_ns.md)
    prefetch_l1(a.ed.tss + (size_t)e*nq + q);
    if constexpr (DEF) prefetch_l1(a.det + (size_t)(e - a.n_car)*nq + q);
    #pragma unroll
    for (int v = 0; v < nu; ++v) prefetch_l1(a.ed.template slot<ND, RS>(e, a.compute_residual ? cache_slot0 + P::update_slot(v) : P::update_slot(v)) + q);
    const double* base = a.faces_ldg + (size_t)e*2*ND*wl;
    for (int i = q; i < 2*ND*nu*nfq; i += nq) svface[i] = base[(size_t)(i/(nu*nfq))*wl + i % (nu*nfq)];
  }
  __syncthreads();
  double r[nu];
  #pragma unroll
  for (int v = 0; v < nu; ++v) r[v] = 0.;
  if (active) {
    #pragma unroll
    for (int d = 0; d < ND; ++d) {
      const Line<ND, RS> ln(d, q);
      const double l0 = s_lift[ln.node*2], l1 = s_lift[ln.node*2 + 1];
      #pragma unroll
      for (int v = 0; v < nu; ++v) {
        double acc = 0;
        acc += l0*svface[((2*d)*nu + v)*nfq + ln.fq];
        acc += l1*svface[((2*d + 1)*nu + v)*nfq + ln.fq];
        r[v] -= acc;
      }
    }
  }
  if (a.use_filter) {
    #pragma unroll
    for (int d = 0; d < ND; ++d) {
      if (active) {
        #pragma unroll
        for (int v = 0; v < nu; ++v) sbuf[v*nq + q] = r[v];
      }
      __syncthreads();
      if (active) {
        const Line<ND, RS> ln(d, q);
        #pragma unroll
        for (int v = 0; v < nu; ++v) {
          double acc = 0;
          for (int k = 0; k < RS; ++k) acc += s_filt[ln.node*RS + k]*sbuf[v*nq + ln.base + k*ln.stride];
          r[v] = acc;
        }
      }
      __syncthreads();
    }
  }
  if (active) {
    const double tss = a.ed.tss[(size_t)e*nq + q];
    double mult = (a.dt_dev ? *a.dt_dev*a.update : a.update)*tss/a.nom[e];
    if constexpr (DEF) mult /= a.det[(size_t)(e - a.n_car)*nq + q];
    double upd[nu]; double* tgt[nu];
    #pragma unroll
    for (int v = 0; v < nu; ++v) {
      upd[v] = r[v]*mult;
      // the reference hands `write_update` the residual cache as base pointer when only the residual is wanted (Spatial.hpp:579,587)
      tgt[v] = a.ed.template slot<ND, RS>(e, a.compute_residual ? cache_slot0 + P::update_slot(v) : P::update_slot(v)) + q;
    }
    P::write_update(a.pp, upd, tgt, tss, true);
    #pragma unroll
    for (int v = 0; v < ne; ++v) sbuf[v*nq + q] = a.ed.template slot<ND, RS>(e, P::extrap_slot(v))[q];
  }
  __syncthreads();
  if (active) extrapolate_faces<ND, RS>(sbuf, ne, a.faces + (size_t)e*2*ND*a.face_width, a.face_width, ops, q);
}

/* ---------------- Write_face ---------------- */
template <int ND, int RS, class P>
__global__ void __launch_bounds__(GCfg<ND, RS, P>::threads)
g_write_face_kernel(GArgs a, Ops ops)
{
  using C = GCfg<ND, RS, P>;
  constexpr int nq = C::nq, ne = C::ne;
  HB_DYN_SMEM(double, smem);
  const int t = threadIdx.x;
  const int le = t/nq, q = t % nq;
  const int e = a.elem_begin + blockIdx.x*C::epb + le;
  const bool active = e < a.elem_end;
  double* sbuf = smem + le*ne*nq;
  if (active) {
    #pragma unroll
    for (int v = 0; v < ne; ++v) sbuf[v*nq + q] = a.ed.template slot<ND, RS>(e, P::extrap_slot(v))[q];
  }
  __syncthreads();
  if (active) extrapolate_faces<ND, RS>(sbuf, ne, a.faces + (size_t)e*2*ND*a.face_width, a.face_width, ops, q);
}

/* ---------------- Max_dt ---------------- */
/* n-linear interpolation of the vertex spacing to point q (reference include/math.hpp:207-218) */
template <int ND, int RS>
__device__ __forceinline__ double g_point_spacing(const GArgs& a, const Ops& ops, int e, int q)
{
  constexpr int n_vert = ipow(2, ND);
  double vals[n_vert];
  #pragma unroll
  for (int i = 0; i < n_vert; ++i) vals[i] = a.vtss[(size_t)e*n_vert + i];
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

/* the local time step 1/scale of one point from its state (already in comp.state) and spacing: Spatial.hpp:808-822 with the five
 * divisions by max_cfl / spacing replaced by one reciprocal of the spacing and the host-side reciprocals of the CFL numbers
 * (<= 1 ulp per use; FP64 division is ~30 instructions and this kernel was issue-bound). ONE function for every caller, so that the
 * screened reduction below returns the very double the unscreened one does. */
template <int ND, class P, class COMP>
__device__ __forceinline__ double g_point_dt(const GArgs& a, COMP& comp, double spacing)
{
  const double inv_spacing = 1./spacing;
  double scale = 0;
  if constexpr (P::has_convection) { comp.compute_char_speed(); scale += comp.char_speed*a.inv_max_cfl_c*inv_spacing; }
  if constexpr (P::has_diffusion) { comp.compute_diffusivity(a.pp); scale += comp.diffusivity*a.inv_max_cfl_d*inv_spacing*inv_spacing; }
  return 1./scale;
}

template <int ND, int RS, class P>
__global__ void __launch_bounds__(256)
g_max_dt_kernel(GArgs a, Ops ops)
{
  constexpr int nq = ipow(RS, ND);
  __shared__ double warp_min[8];
  const long long gid = (long long)blockIdx.x*blockDim.x + threadIdx.x;
  const int e = (int)(gid/nq), q = (int)(gid % nq);
  double val = DBL_MAX;
  if (e < a.elem_end) {
    const double spacing = g_point_spacing<ND, RS>(a, ops, e, q);
    typename P::template Comp<ND> comp;
    #pragma unroll
    for (int i = 0; i < P::n_state; ++i) comp.state[i] = a.ed.template slot<ND, RS>(e, P::state_slot(i))[q];
    const double local_dt = g_point_dt<ND, P>(a, comp, spacing);
    if (a.is_local) a.ed.tss[(size_t)e*nq + q] = local_dt;
    else { if (a.write_tss) a.ed.tss[(size_t)e*nq + q] = 1.; val = local_dt; }
  }
  if (a.is_local) return;
  #pragma unroll
  for (int off = 16; off > 0; off /= 2) val = fmin(val, __shfl_xor_sync(0xffffffffu, val, off));
  if (threadIdx.x % 32 == 0) warp_min[threadIdx.x/32] = val;
  __syncthreads();
  if (threadIdx.x == 0) {
    double m = warp_min[0];
    for (int i = 1; i < (int)blockDim.x/32; ++i) m = fmin(m, warp_min[i]);
    atomicMin(a.global_min, (unsigned long long)__double_as_longlong(m));
  }
}

/* Global time stepping with a running single-precision screen: the scheme of max_dt_euler_screen_kernel (misc_kernels.cu) for the
 * PDEs that provide P::screen_dt (Navier-Stokes). Points whose screen value is within P::screen_margin of min(running, warp minimum)
 * -- or 0 = not trustworthy -- are evaluated with g_point_dt and folded into the result with atomicMin; everything else is skipped. */
template <int ND, int RS, class P>
__global__ void __launch_bounds__(256)
g_max_dt_screen_kernel(GArgs a, Ops ops, int* screen_bits)
{
  constexpr int nq = ipow(RS, ND);
  const unsigned gid = blockIdx.x*256u + threadIdx.x; // the launcher checks that the point count fits 32 bits
  const unsigned e = gid/nq, q = gid - e*nq;
  const bool valid = e < (unsigned)a.elem_end;
  typename P::template Comp<ND> comp;
  double spacing = 1.;
  float ap = 3.0e38f;
  if (valid) {
    spacing = g_point_spacing<ND, RS>(a, ops, (int)e, (int)q);
    #pragma unroll
    for (int i = 0; i < P::n_state; ++i) comp.state[i] = a.ed.template slot<ND, RS>((int)e, P::state_slot(i))[q];
    if (a.write_tss) a.ed.tss[(size_t)e*nq + q] = 1.;
    float s[P::n_state];
    #pragma unroll
    for (int i = 0; i < P::n_state; ++i) s[i] = (float)comp.state[i];
    const float inv_h = __fdividef(1.f, (float)spacing);
    ap = P::screen_dt(s, a.pp, (float)a.inv_max_cfl_c*inv_h, (float)a.inv_max_cfl_d*inv_h*inv_h);
    if (!(ap > 0.f && ap < 3.0e38f)) ap = 0.f;
  }
  float wm = ap > 0.f ? ap : 3.0e38f;
  #pragma unroll
  for (int off = 16; off > 0; off /= 2) wm = fminf(wm, __shfl_xor_sync(0xffffffffu, wm, off));
  float g = 3.0e38f;
  if (threadIdx.x % 32 == 0) {
    g = __int_as_float(__ldca(screen_bits));
    if (wm < g) { atomicMin(screen_bits, __float_as_int(wm)); g = wm; }
  }
  g = __shfl_sync(0xffffffffu, g, 0);
  if (valid && (ap == 0.f || ap <= g*P::screen_margin)) {
    const double exact = g_point_dt<ND, P>(a, comp, spacing);
    const unsigned long long bits = (unsigned long long)__double_as_longlong(exact);
    if (exact > 0. && bits < __ldca(a.global_min)) atomicMin(a.global_min, bits);
  }
}

/* ---------------- host side ---------------- */
template <int ND, int RS> using Pde = typename PdeSelect<HB_PDE, ND, RS>::type;

int fill_args(hexed_b200_ctx* c, GArgs& a, const PdeParams& pp)
{
  const int kind = (HB_PDE == PDE_ADVECTION) ? 2 : 0;
  const size_t nq = c->nq, ne = c->n_elem;
  auto need = [&](double** arr, size_t per_elem) -> int {
    if (*arr) return 0;
    HB_CUDA(c, cudaMalloc(arr, sizeof(double)*(ne ? ne : 1)*per_elem));
    HB_CUDA(c, cudaMemsetAsync(*arr, 0, sizeof(double)*(ne ? ne : 1)*per_elem, c->stream));
    return 0;
  };
  int rc = 0;
  if (HB_PDE == PDE_NAVIER_STOKES) rc = need(&c->av, 2*nq);
  if (!rc && HB_PDE == PDE_SMOOTH_AV) rc = need(&c->forcing, 4*nq);
  if (!rc && HB_PDE == PDE_ADVECTION) rc = need(&c->adv, (size_t)c->rs*nq);
  if (rc) return rc;
  const size_t nslot = c->n_face_slot ? c->n_face_slot : 1;
  if (HB_PDE != PDE_ADVECTION && !c->face_ldg) {
    HB_CUDA(c, cudaMalloc(&c->face_ldg, sizeof(double)*nslot*c->nv*c->nfq));
    HB_CUDA(c, cudaMemsetAsync(c->face_ldg, 0, sizeof(double)*nslot*c->nv*c->nfq, c->stream));
  }
  if (HB_PDE == PDE_ADVECTION && !c->face_wide) {
    HB_CUDA(c, cudaMalloc(&c->face_wide, sizeof(double)*nslot*(c->nd + c->rs)*c->nfq));
    HB_CUDA(c, cudaMemsetAsync(c->face_wide, 0, sizeof(double)*nslot*(c->nd + c->rs)*c->nfq, c->stream));
  }
  a.ed.state = c->state; a.ed.tss = c->tss; a.ed.av = c->av; a.ed.forcing = c->forcing; a.ed.adv = c->adv; a.ed.cache = c->cache;
  a.nom = c->nom; a.refn = c->refn; a.det = c->det; a.normals = c->normals; a.vtss = c->vtss;
  a.faces = kind == 2 ? c->face_wide : c->face_state; a.faces_ldg = c->face_ldg;
  a.face_width = (kind == 2 ? c->nd + c->rs : c->nv)*c->nfq;
  a.con = nullptr; a.perm = c->perm; a.n_con = 0;
  a.elem_begin = 0; a.elem_end = c->n_elem; a.n_car = c->n_car;
  a.update = 0; a.stage = 0; a.compute_residual = 0; a.use_filter = 0; a.dt_dev = c->dt_dev_active;
  a.max_cfl_c = 1; a.max_cfl_d = 1; a.inv_max_cfl_c = 1; a.inv_max_cfl_d = 1; a.is_local = 0; a.global_min = reinterpret_cast<unsigned long long*>(c->d_scalar);
  a.pp = pp;
  return 0;
}

int g_neighbor(hexed_b200_ctx* c, int deformed, const PdeParams& pp, bool reconcile, int first, int count)
{
  const int n_all = deformed ? c->n_def_con : c->n_car_con;
  if (count < 0) count = n_all - first;
  if (first < 0 || first + count > n_all) return fail(c, HEXED_B200_BAD_ARGUMENT, "connection range out of bounds");
  const int n_con = count;
  StatScope scope(c, deformed ? ST_NEIGHBOR_DEF : ST_NEIGHBOR_CAR, n_con);
  GArgs a; int rc = fill_args(c, a, pp); if (rc) return rc; // also allocates the LDG face storage on first use
  if (!n_con) return 0;
  invalidate_admis(c);
  a.con = (deformed ? c->def_con : c->car_con) + (size_t)first*4; a.n_con = n_con;
  return dispatch(c, [&](auto nd, auto rs) {
    constexpr int ND = decltype(nd)::value, RS = decltype(rs)::value;
    using P = Pde<ND, RS>;
    const long long total = (long long)n_con*ipow(RS, ND - 1);
    const int grid = (int)((total + 127)/128);
    if (reconcile) {
      if constexpr (P::has_diffusion) {
        if (deformed) { auto k = g_neighbor_reconcile_kernel<ND, RS, P, true>; HB_LAUNCH(k, grid, 128, 0, c->stream, a); }
        else { auto k = g_neighbor_reconcile_kernel<ND, RS, P, false>; HB_LAUNCH(k, grid, 128, 0, c->stream, a); }
      } else return fail(c, HEXED_B200_BAD_ARGUMENT, "Neighbor_reconcile needs a diffusive PDE");
    } else {
      if (deformed) { auto k = g_neighbor_kernel<ND, RS, P, true>; HB_LAUNCH(k, grid, 128, 0, c->stream, a); }
      else { auto k = g_neighbor_kernel<ND, RS, P, false>; HB_LAUNCH(k, grid, 128, 0, c->stream, a); }
    }
    count_launch(c, deformed ? ST_NEIGHBOR_DEF : ST_NEIGHBOR_CAR);
    HB_CUDA(c, cudaGetLastError());
    return 0;
  });
}

int g_local(hexed_b200_ctx* c, int deformed, hexed_b200_options o, const PdeParams& pp, bool reconcile)
{
  const int begin = deformed ? c->n_car : 0, end = deformed ? c->n_elem : c->n_car;
  StatScope scope(c, reconcile ? (deformed ? ST_RECONCILE_DEF : ST_RECONCILE_CAR) : (deformed ? ST_LOCAL_DEF : ST_LOCAL_CAR), end - begin);
  GArgs a; int rc = fill_args(c, a, pp); if (rc) return rc;
  invalidate_cfl_cache(c); // these kernels may rewrite the flow state
  a.elem_begin = begin; a.elem_end = end;
  a.stage = o.i_stage != 0; a.compute_residual = o.compute_residual; a.use_filter = o.use_filter;
  a.update = (a.stage && !reconcile) ? o.dt*(.5/c->quad_safety) : o.dt; // Spatial.hpp:317 (Local) / :534 (Reconcile_ldg_flux)
  return dispatch(c, [&](auto nd, auto rs) {
    constexpr int ND = decltype(nd)::value, RS = decltype(rs)::value;
    using P = Pde<ND, RS>;
    using C = GCfg<ND, RS, P>;
    // argument checks of the reference constructors (Spatial.hpp:322-323,539-540)
    if (P::has_diffusion && a.stage) return fail(c, HEXED_B200_BAD_ARGUMENT, "two-stage stabilization is not applicable to diffusion equations");
    if (a.stage && a.compute_residual) return fail(c, HEXED_B200_BAD_ARGUMENT, "residual calculation is a single-stage operation");
    if (end == begin) return 0;
    const int grid = (end - begin + C::epb - 1)/C::epb;
    const size_t smem = sizeof(double)*(reconcile ? C::rec_smem_doubles : C::smem_doubles);
    if (reconcile) {
      if constexpr (P::has_diffusion) {
        if (a.compute_residual && HB_PDE == PDE_SMOOTH_AV)
          return fail(c, HEXED_B200_NOT_IMPLEMENTED, "residual-only LDG reconciliation is undefined for Smooth_art_visc (the reference indexes past the residual cache)");
#if HB_PDE == 1 /* PDE_NAVIER_STOKES */
        {
          const int r = launch_ns_reconcile_bulk<ND, RS>(c, a, deformed);
          if (r > 0) return r;
          if (r == 0) { count_launch(c, deformed ? ST_RECONCILE_DEF : ST_RECONCILE_CAR); HB_CUDA(c, cudaGetLastError()); return 0; }
        }
#endif
        if (deformed) { auto k = g_reconcile_kernel<ND, RS, P, true>; int r = set_smem(c, k, smem); if (r) return r; HB_LAUNCH(k, grid, C::threads, smem, c->stream, a, c->ops, c->filt); }
        else { auto k = g_reconcile_kernel<ND, RS, P, false>; int r = set_smem(c, k, smem); if (r) return r; HB_LAUNCH(k, grid, C::threads, smem, c->stream, a, c->ops, c->filt); }
      } else return fail(c, HEXED_B200_BAD_ARGUMENT, "Reconcile_ldg_flux needs a diffusive PDE");
    } else {
#if HB_PDE == 1 /* PDE_NAVIER_STOKES */
      {
        const int r = launch_ns_local_line<ND, RS>(c, a, deformed);
        if (r > 0) return r;
        if (r == 0) { count_launch(c, deformed ? ST_LOCAL_DEF : ST_LOCAL_CAR); HB_CUDA(c, cudaGetLastError()); return 0; }
      }
#endif
      if (deformed) { auto k = g_local_kernel<ND, RS, P, true>; int r = set_smem(c, k, smem); if (r) return r; HB_LAUNCH(k, grid, C::threads, smem, c->stream, a, c->ops, c->filt); }
      else { auto k = g_local_kernel<ND, RS, P, false>; int r = set_smem(c, k, smem); if (r) return r; HB_LAUNCH(k, grid, C::threads, smem, c->stream, a, c->ops, c->filt); }
    }
    count_launch(c, reconcile ? (deformed ? ST_RECONCILE_DEF : ST_RECONCILE_CAR) : (deformed ? ST_LOCAL_DEF : ST_LOCAL_CAR));
    HB_CUDA(c, cudaGetLastError());
    return 0;
  });
}

int g_write_face(hexed_b200_ctx* c, const PdeParams& pp)
{
  StatScope scope(c, ST_WRITE_FACE, c->n_elem);
  invalidate_admis(c);
  GArgs a; int rc = fill_args(c, a, pp); if (rc) return rc;
  if (!c->n_elem) return 0;
  return dispatch(c, [&](auto nd, auto rs) {
    constexpr int ND = decltype(nd)::value, RS = decltype(rs)::value;
    using P = Pde<ND, RS>;
    using C = GCfg<ND, RS, P>;
    const int grid = (c->n_elem + C::epb - 1)/C::epb;
    const size_t smem = sizeof(double)*C::epb*C::ne*C::nq;
    auto k = g_write_face_kernel<ND, RS, P>;
    int r = set_smem(c, k, smem); if (r) return r;
    HB_LAUNCH(k, grid, C::threads, smem, c->stream, a, c->ops);
    count_launch(c, ST_WRITE_FACE);
    HB_CUDA(c, cudaGetLastError());
    return 0;
  });
}

int g_max_dt(hexed_b200_ctx* c, const PdeParams& pp, double safety_conv, double safety_diff, int local_time, double* dt)
{
  StatScope s_car(c, ST_MAX_DT_CAR, c->n_car);
  c->stats[ST_MAX_DT_DEF].work_units += c->n_def;
  if (!c->n_elem) { *dt = local_time ? 1. : DBL_MAX; return 0; }
  GArgs a; int rc = fill_args(c, a, pp); if (rc) return rc;
  a.max_cfl_c = (-2*c->quad_safety/c->min_eig_conv)*safety_conv; // Basis::max_cfl (src/Basis.cpp:6-9), Spatial.hpp:777
  a.max_cfl_d = -2/c->min_eig_diff*safety_diff;                   // Spatial.hpp:778
    a.inv_max_cfl_c = 1./a.max_cfl_c; a.inv_max_cfl_d = 1./a.max_cfl_d;
  a.is_local = local_time;
  a.write_tss = !c->tss_is_one;
  return dispatch(c, [&](auto nd, auto rs) {
    constexpr int ND = decltype(nd)::value, RS = decltype(rs)::value;
    using P = Pde<ND, RS>;
    const long long total = (long long)c->n_elem*ipow(RS, ND);
    const int grid = (int)((total + 255)/256);
    double* result = c->max_dt_device_out ? c->max_dt_device_out : c->d_scalar; // hexed_b200_update_*: leave dt on the device
    a.global_min = reinterpret_cast<unsigned long long*>(result);
    if (!local_time) HB_CUDA(c, cudaMemsetAsync(result, 0x7f, sizeof(double), c->stream));
    if constexpr (P::has_float_screen) {
      if (!local_time && total < (1ll << 32) - 256) {
        int* screen_bits = reinterpret_cast<int*>(c->d_scalar + 1);
        HB_CUDA(c, cudaMemsetAsync(screen_bits, 0x7f, sizeof(int), c->stream)); // 0x7f7f7f7f = 3.39e38
        auto k = g_max_dt_screen_kernel<ND, RS, P>; HB_LAUNCH(k, grid, 256, 0, c->stream, a, c->ops, screen_bits);
      }
      else { auto k = g_max_dt_kernel<ND, RS, P>; HB_LAUNCH(k, grid, 256, 0, c->stream, a, c->ops); }
    }
    else { auto k = g_max_dt_kernel<ND, RS, P>; HB_LAUNCH(k, grid, 256, 0, c->stream, a, c->ops); }
    count_launch(c, ST_MAX_DT_CAR);
    HB_CUDA(c, cudaGetLastError());
    c->tss_is_one = !local_time;
    if (local_time) { *dt = 1.; return 0; }
    if (c->max_dt_device_out) { *dt = 1.; return 0; }
    HB_CUDA(c, cudaMemcpyAsync(c->h_scalar, c->d_scalar, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    HB_CUDA(c, cudaStreamSynchronize(c->stream));
    *dt = *c->h_scalar;
    return 0;
  });
}

} // namespace

#define HB_CAT2(a, b) a##b
#define HB_CAT(a, b) HB_CAT2(a, b)
const GenericOps HB_CAT(generic_ops_pde, HB_PDE) = {g_neighbor, g_local, g_write_face, g_max_dt};

} // namespace hb
