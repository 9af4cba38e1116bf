/* local_euler_pipe.cu -- persistent, TMA-pipelined version of the Euler `Local` kernel for 3-D elements.
 *
 * Same arithmetic as local_euler.cu (reference include/Spatial.hpp:326-509 + the trailing write_face :41-57,507) but organised
 * around the two things the first profile showed to bound that kernel (profiles/r01a_ncu_full.md: 76 % L1/shared-memory
 * pipe, long-scoreboard stalls on the global loads, 53 % occupancy):
 *
 *  1. HBM latency. A persistent CTA (grid = SMs x resident CTAs) walks elements blockIdx.x, +gridDim.x, ... . Every input of an
 *     element is one contiguous run in the element-major layout (state nv*nq, the 2*ND numerical-flux faces, ND*ND normals,
 *     determinant, time-step scale, residual cache), so one elected thread fetches them with 1-D bulk TMA copies
 *     (cp.async.bulk -> UBLKCP) that complete on mbarriers. The (state, faces, normals) set is double buffered: the copies for
 *     element i+2 are in flight while element i is computed; the late inputs (cache, tss, det) are single buffered and
 *     fetched one phase ahead. No thread ever stalls on a global load.
 *  2. Shared-memory traffic. Work is split into LINE tasks (one thread owns the row_size points of one line of one dimension)
 *     instead of one thread per point: a line task reads each flux value once and produces row_size derivative values from
 *     registers, with the 1-D operator entries as constant-bank operands. 2.6x fewer shared-memory bytes per element than the
 *     point-per-thread kernel.
 *
 * Phases per element: A (line tasks: pointwise flux on the line, derivative + lifted face flux -> R_d[v][q]),
 * B (point tasks: r = R_0 + R_1 + R_2, two-stage update, new state -> HBM and in place in shared memory),
 * C (line tasks: extrapolate the new state to both faces of the line -> HBM).
 */
#include "euler.cuh"

namespace hb {

/* LEAN (deformed elements): only the state is double buffered; the numerical-flux faces and reference normals are dead after phase A
 * and live in ONE buffer that is refilled right after it; time-step scale, determinant and residual cache are read once, straight from
 * HBM after an L2 prefetch issued a phase earlier. 67 KB instead of 105 KB per CTA at row size 6: three resident CTAs per SM instead
 * of two (the same restructuring took the 2-D kernel from 0.72 to 0.88 of the HBM bandwidth, local_euler_pipe2d.cu). */
template <int RS, bool DEF, bool LEAN = false>
struct PipeCfg
{
  static constexpr int ND = 3, nq = RS*RS*RS, nfq = RS*RS, nv = 5;
  static constexpr int n_line = ND*nfq;
  static constexpr int threads = ((n_line + 31)/32)*32;
  static_assert(RS % 2 == 0, "the line tasks process points in pairs");
  static constexpr int cs = nv > RS ? nv : RS;
  // double-buffered stage: state | numerical flux faces | reference level normals (LEAN: the state only)
  static constexpr int st_state = 0, st_face = nv*nq, st_nrml = st_face + 2*ND*nv*nfq;
  static constexpr int stage_doubles = LEAN ? nv*nq : st_nrml + (DEF ? ND*ND*nq : 0);
  // single-buffered late inputs: residual cache | tss | det (LEAN: this region holds faces | normals instead)
  static constexpr int lt_cache = 0, lt_tss = nv*nq, lt_det = lt_tss + nq;
  static constexpr int lt_vtss = lt_det + (DEF ? nq : 0);
  static constexpr int fn_face = 0, fn_nrml = 2*ND*nv*nfq;
  static constexpr int late_doubles = LEAN ? fn_nrml + (DEF ? ND*ND*nq : 0) : lt_vtss + 8; // + the 2^3 vertex time-step scales (used by the CFL screen)
  static constexpr int r_doubles = ND*nv*nq;
  static constexpr int n_iter = (nq + threads - 1)/threads; // point tasks per thread
  static constexpr int smem_doubles = 2*stage_doubles + late_doubles + r_doubles;
  static constexpr size_t smem_bytes = sizeof(double)*smem_doubles + 4*sizeof(mbar_t) + 9*sizeof(double); // + mbarriers + CFL screen scratch (4 per-warp minima, 8 vertex spacings, floats) + the element's nominal size
};

struct PipeArgs
{
  double* state; const double* tss /* null: time_step_scale is known to hold 1 everywhere (ctx::tss_is_one), not read */; double* cache; const double* nom; const double* refn; const double* det; double* faces;
  int elem_begin, elem_end, n_car;
  double update; int stage; int compute_residual;
  const double* dt_dev; // non-null: the time step lives on the device and multiplies `update`
  int end_barrier; // 1 (default): end every iteration with a CTA barrier; 0 (HEXED_B200_OPT_PIPELINED_LOCAL value 4): mbarrier hand-over of the stage buffer
  int* record; // non-null: leave Element::record-style admissibility bits of the NEW state and faces per element (bit 0 inadmissible, bit 1 non-finite)
  const double* vtss; float nodef[MAX_RS]; float* cfl_approx; // CFL instantiation: single-precision min_q spacing/char_speed of the NEW state per element
};

template <int RS, bool DEF>
__device__ __forceinline__ void pipe_issue_stage(const PipeArgs& a, int e, double* buf, mbar_t* bar)
{
  using C = PipeCfg<RS, DEF>;
  constexpr unsigned b_state = sizeof(double)*C::nv*C::nq, b_face = sizeof(double)*2*C::ND*C::nv*C::nfq, b_nrml = sizeof(double)*C::ND*C::ND*C::nq;
  mbar_arrive_expect_tx(bar, b_state + b_face + (DEF ? b_nrml : 0u));
  bulk_g2s(buf + C::st_state, a.state + (size_t)e*C::nv*C::nq, b_state, bar);
  bulk_g2s(buf + C::st_face, a.faces + (size_t)e*2*C::ND*C::nv*C::nfq, b_face, bar);
  if constexpr (DEF) bulk_g2s(buf + C::st_nrml, a.refn + (size_t)(e - a.n_car)*C::ND*C::ND*C::nq, b_nrml, bar);
}

template <int RS, bool DEF>
__device__ __forceinline__ void pipe_issue_state(const PipeArgs& a, int e, double* buf, mbar_t* bar)
{
  using C = PipeCfg<RS, DEF, true>;
  constexpr unsigned b_state = sizeof(double)*C::nv*C::nq;
  mbar_arrive_expect_tx(bar, b_state);
  bulk_g2s(buf, a.state + (size_t)e*C::nv*C::nq, b_state, bar);
}

template <int RS, bool DEF>
__device__ __forceinline__ void pipe_issue_fn(const PipeArgs& a, int e, double* buf, mbar_t* bar)
{
  using C = PipeCfg<RS, DEF, true>;
  constexpr unsigned b_face = sizeof(double)*2*C::ND*C::nv*C::nfq, b_nrml = sizeof(double)*C::ND*C::ND*C::nq;
  mbar_arrive_expect_tx(bar, b_face + (DEF ? b_nrml : 0u));
  bulk_g2s(buf + C::fn_face, a.faces + (size_t)e*2*C::ND*C::nv*C::nfq, b_face, bar);
  if constexpr (DEF) bulk_g2s(buf + C::fn_nrml, a.refn + (size_t)(e - a.n_car)*C::ND*C::ND*C::nq, b_nrml, bar);
}

/* the late inputs of element e towards L2: one 128-byte line per call (the arrays are only 64-byte aligned per element: the last
 * line is touched explicitly) */
template <int RS, bool DEF>
__device__ __forceinline__ void pipe_prefetch_late(const PipeArgs& a, int e, int t)
{
  using C = PipeCfg<RS, DEF, true>;
  if (a.tss) {
    const double* tss = a.tss + (size_t)e*C::nq;
    for (int i = t*16; i < C::nq; i += C::threads*16) prefetch_l2(tss + i);
    if (t == 0) prefetch_l2(tss + C::nq - 1);
  }
  if constexpr (DEF) {
    const double* det = a.det + (size_t)(e - a.n_car)*C::nq;
    for (int i = t*16; i < C::nq; i += C::threads*16) prefetch_l2(det + i);
    if (t == 1) prefetch_l2(det + C::nq - 1);
  }
  if (a.stage) {
    const double* cache = a.cache + (size_t)e*C::cs*C::nq;
    for (int i = t*16; i < C::nv*C::nq; i += C::threads*16) prefetch_l2(cache + i);
    if (t == 2) prefetch_l2(cache + C::nv*C::nq - 1);
  }
}

template <int RS, bool DEF, bool CFL>
__device__ __forceinline__ void pipe_issue_late(const PipeArgs& a, int e, double* buf, mbar_t* bar)
{
  using C = PipeCfg<RS, DEF>;
  constexpr unsigned b_cache = sizeof(double)*C::nv*C::nq, b_pt = sizeof(double)*C::nq;
  mbar_arrive_expect_tx(bar, (a.stage ? b_cache : 0u) + (a.tss ? b_pt : 0u) + (DEF ? b_pt : 0u) + (CFL ? 64u : 0u));
  if constexpr (CFL) bulk_g2s(buf + C::lt_vtss, a.vtss + (size_t)e*8, 64u, bar);
  if (a.stage) bulk_g2s(buf + C::lt_cache, a.cache + (size_t)e*C::cs*C::nq, b_cache, bar);
  if (a.tss) bulk_g2s(buf + C::lt_tss, a.tss + (size_t)e*C::nq, b_pt, bar);
  if constexpr (DEF) bulk_g2s(buf + C::lt_det, a.det + (size_t)(e - a.n_car)*C::nq, b_pt, bar);
}

template <int RS, bool DEF, bool CFL, bool LEAN>
__device__ __forceinline__ void pipe_body(const PipeArgs& a, const Ops& ops)
{
  using C = PipeCfg<RS, DEF, LEAN>;
  static_assert(!(LEAN && CFL), "the CFL-screen instantiation keeps the classic layout");
  constexpr int ND = 3, nq = C::nq, nfq = C::nfq, nv = C::nv;
  HB_DYN_SMEM(double, smem);
  double* late = smem + 2*C::stage_doubles;
  double* R = late + C::late_doubles;
  mbar_t* bars = reinterpret_cast<mbar_t*>(R + C::r_doubles); // [0],[1]: stage buffers; [2]: late inputs
  [[maybe_unused]] float* warp_cfl = reinterpret_cast<float*>(bars + 4); // per-warp minima of the single-precision CFL screen
  double* s_nom = reinterpret_cast<double*>(bars + 4) + 8;                // the element's nominal size, broadcast by thread 0
  const int t = threadIdx.x;
  const int stride_e = gridDim.x;
  int e = a.elem_begin + blockIdx.x;
  if (e >= a.elem_end) return;

  if (t == 0) {
    mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_init(&bars[2], 1);
    mbar_init(&bars[3], C::threads); // "stage buffer consumed": every thread arrives after phase C, only the refilling thread waits
    mbar_init_fence();
  }
  __syncthreads();
  if constexpr (LEAN) {
    if (t == 0) {
      pipe_issue_state<RS, DEF>(a, e, smem, &bars[0]);
      pipe_issue_fn<RS, DEF>(a, e, late, &bars[2]);
      if (e + stride_e < a.elem_end) pipe_issue_state<RS, DEF>(a, e + stride_e, smem + C::stage_doubles, &bars[1]);
    }
    pipe_prefetch_late<RS, DEF>(a, e, t);
  } else if (t == 0) {
    pipe_issue_stage<RS, DEF>(a, e, smem, &bars[0]);
    if (e + stride_e < a.elem_end) pipe_issue_stage<RS, DEF>(a, e + stride_e, smem + C::stage_doubles, &bars[1]);
    pipe_issue_late<RS, DEF, CFL>(a, e, late, &bars[2]);
  }

  // line task of this thread: dimension d, line l (its face quadrature point), points q0 + k*stride
  int d, l;
  using Map = LineMap<RS, !DEF>; // measured: pays off for the Cartesian kernels only (see common.cuh)
  const bool has_line = Map::get(t, d, l);
  const bool vec = Map::vec2 && d == 2; // this thread's line is contiguous: 16-byte accesses
  const int stride = d == 0 ? RS*RS : d == 1 ? RS : 1;
  const int q0 = d == 0 ? l : d == 1 ? (l/RS)*RS*RS + l % RS : l*RS;

  // the time step (when it lives on the device) is the same for every element. The nominal size of an element is fetched by thread 0 at the top
  // of its iteration and handed to the others through shared memory just before the barrier that ends phase A, so that the HBM latency hides
  // behind phase A: read where it is used it cost 12 % of all stall samples (profiles/r02k_ncu_full_euler_car.md: the reciprocal in phase B), and
  // read by every thread at the top the compiler moved it to a uniform register at once, with the same wait (profiles/r02t_ncu_full_euler_car.md)
  const double update = a.dt_dev ? *a.dt_dev*a.update : a.update;
  for (int it = 0; e < a.elem_end; ++it, e += stride_e) {
    double nom_t0 = 0.;
    if (t == 0) nom_t0 = a.nom[e];
    const int s = it & 1;
    const unsigned par = (it >> 1) & 1;
    double* const stage_buf = smem + s*C::stage_doubles;
    double* S = stage_buf + C::st_state;
    const double* F = LEAN ? late + C::fn_face : stage_buf + C::st_face;
    const double* N = LEAN ? late + C::fn_nrml : stage_buf + C::st_nrml;
    mbar_wait(&bars[s], par);
    if constexpr (LEAN) mbar_wait(&bars[2], it & 1);
    int bad = 0; // thermodynamic admissibility of what this thread writes (Solver::is_admissible, fused: see misc_kernels.cu)

    /* ---- phase A: flux on the line, then D(flux, face flux) -> R_d ---- */
    if constexpr (Map::vec2) {
      if (has_line) {
        double f[nv][RS]; // the state on the line, replaced point by point by the flux through reference direction d
        if (vec) {
          #pragma unroll
          for (int v = 0; v < nv; ++v)
            #pragma unroll
            for (int k = 0; k + 1 < RS; k += 2) ld2(S + v*nq + q0 + k, f[v][k], f[v][k + 1]);
        } else {
          #pragma unroll
          for (int v = 0; v < nv; ++v)
            #pragma unroll
            for (int k = 0; k < RS; ++k) f[v][k] = S[v*nq + q0 + k*stride];
        }
        #pragma unroll
        for (int k2 = 0; k2 < RS; k2 += 2) {
          [[maybe_unused]] double n[2][ND];
          if constexpr (DEF) {
            // normal[j] of reference direction d at these points: refn[d][j][q] (reference Spatial.hpp:411)
            if (vec) {
              #pragma unroll
              for (int j = 0; j < ND; ++j) ld2(N + (d*ND + j)*nq + q0 + k2, n[0][j], n[1][j]);
            } else {
              #pragma unroll
              for (int j = 0; j < ND; ++j) { n[0][j] = N[(d*ND + j)*nq + q0 + k2*stride]; n[1][j] = N[(d*ND + j)*nq + q0 + (k2 + 1)*stride]; }
            }
          }
          #pragma unroll
          for (int kk = 0; kk < 2; ++kk) {
            const int k = k2 + kk;
            EulerPoint<ND> p;
            #pragma unroll
            for (int v = 0; v < nv; ++v) p.s[v] = f[v][k];
            p.scalars();
            double fl[nv];
            if constexpr (DEF) p.flux(n[kk], fl);
            else {
              // unit normal e_d; selects instead of a dynamically indexed register array
              const double mass_flux = d == 0 ? p.s[0] : d == 1 ? p.s[1] : p.s[2];
              const double vol_flux = mass_flux*p.inv_mass;
              fl[ND] = mass_flux;
              fl[ND + 1] = (p.s[ND + 1] + p.pressure)*vol_flux;
              #pragma unroll
              for (int j = 0; j < ND; ++j) fl[j] = p.s[j]*vol_flux + (j == d ? p.pressure : 0.);
            }
            #pragma unroll
            for (int v = 0; v < nv; ++v) f[v][k] = fl[v];
          }
        }
        #pragma unroll
        for (int v = 0; v < nv; ++v) {
          const double b0 = F[((2*d)*nv + v)*nfq + l], b1 = F[((2*d + 1)*nv + v)*nfq + l];
          double r[RS];
          line_deriv_eo<RS, true>(ops, f[v], b0, b1, r);
          double* row = R + (d*nv + v)*nq + q0;
          if (vec) {
            #pragma unroll
            for (int i = 0; i + 1 < RS; i += 2) st2(row + i, r[i], r[i + 1]);
          } else {
            #pragma unroll
            for (int i = 0; i < RS; ++i) row[i*stride] = r[i];
          }
        }
      }
    } else {
      if (has_line) {
        double f[nv][RS];
        #pragma unroll
        for (int k = 0; k < RS; ++k) {
          EulerPoint<ND> p;
          #pragma unroll
          for (int v = 0; v < nv; ++v) p.s[v] = S[v*nq + q0 + k*stride];
          p.scalars();
          double fl[nv];
          if constexpr (DEF) {
            double n[ND];
            // normal[j] of reference direction d at this point: refn[d][j][q] (reference Spatial.hpp:411)
            #pragma unroll
            for (int j = 0; j < ND; ++j) n[j] = N[(d*ND + j)*nq + q0 + k*stride];
            p.flux(n, fl);
          } else {
            // unit normal e_d; selects instead of a dynamically indexed register array
            const double mass_flux = d == 0 ? p.s[0] : d == 1 ? p.s[1] : p.s[2];
            const double vol_flux = mass_flux*p.inv_mass;
            fl[ND] = mass_flux;
            fl[ND + 1] = (p.s[ND + 1] + p.pressure)*vol_flux;
            #pragma unroll
            for (int j = 0; j < ND; ++j) fl[j] = p.s[j]*vol_flux + (j == d ? p.pressure : 0.);
          }
          #pragma unroll
          for (int v = 0; v < nv; ++v) f[v][k] = fl[v];
        }
        #pragma unroll
        for (int v = 0; v < nv; ++v) {
          const double b0 = F[((2*d)*nv + v)*nfq + l], b1 = F[((2*d + 1)*nv + v)*nfq + l];
          double r[RS];
          line_deriv_eo<RS, true>(ops, f[v], b0, b1, r);
          #pragma unroll
          for (int i = 0; i < RS; ++i) R[(d*nv + v)*nq + q0 + i*stride] = r[i];
        }
      }
    }
    /* (lean) the late inputs of this thread's points are loaded from HBM (L2 hits after the prefetch of a phase ago) BEFORE the barrier that
     * ends phase A: they depend on nothing phase A computes, and their latency then overlaps the wait for the slowest warp
     * (profiles/r02n_ncu_full_euler.md: 11 % of all stall samples sat on the first use of these loads at the top of phase B) */
    [[maybe_unused]] double l_tss[C::n_iter], l_cache[C::n_iter][nv], l_det[C::n_iter];
    if constexpr (LEAN) {
      #pragma unroll
      for (int k = 0; k < C::n_iter; ++k) {
        const int q = t + k*C::threads;
        if (q < nq) {
          l_tss[k] = a.tss ? a.tss[(size_t)e*nq + q] : 1.;
          if constexpr (DEF) l_det[k] = a.det[(size_t)(e - a.n_car)*nq + q];
          if (a.stage) {
            #pragma unroll
            for (int v = 0; v < nv; ++v) l_cache[k][v] = a.cache[((size_t)e*C::cs + v)*nq + q];
          }
        }
      }
    }
    if (t == 0) *s_nom = nom_t0;
    __syncthreads(); // R complete; faces / normals of this stage buffer are dead, the state is still needed
    const double nom = *s_nom;

    if constexpr (LEAN) {
      // faces / normals buffer is dead: refill it for the next element, and pull that element's late inputs towards L2
      if (e + stride_e < a.elem_end) {
        if (t == 0) { fence_proxy_async(); pipe_issue_fn<RS, DEF>(a, e + stride_e, late, &bars[2]); }
        pipe_prefetch_late<RS, DEF>(a, e + stride_e, t);
      }
      /* ---- phase B (lean) ---- */
      #pragma unroll
      for (int k = 0; k < C::n_iter; ++k) {
        const int q = t + k*C::threads;
        if (q < nq) {
          double mult; // update*tss/nom/det (reference Spatial.hpp:484-487) with one division instead of two (<= 1 ulp)
          if constexpr (DEF) mult = update*l_tss[k]/(nom*l_det[k]);
          else mult = update*l_tss[k]/nom;
          #pragma unroll
          for (int v = 0; v < nv; ++v) {
            double u = R[(0*nv + v)*nq + q];
            u += R[(1*nv + v)*nq + q];
            u += R[(2*nv + v)*nq + q];
            double* cache = a.cache + ((size_t)e*C::cs + v)*nq + q;
            if (a.stage) u -= l_cache[k][v];
            else if (!a.compute_residual) *cache = u;
            u *= mult;
            if (a.compute_residual) *cache = u;
            else {
              const double xv = S[v*nq + q] + u;
              S[v*nq + q] = xv;
              a.state[((size_t)e*nv + v)*nq + q] = xv;
              if (a.record) bad |= (isfinite(xv) ? 0 : 2) | ((v >= ND && !(xv > 0.)) ? 1 : 0);
            }
          }
        }
      }
    } else {
    /* ---- phase B: combine, two-stage update (reference Spatial.hpp:484-503) ---- */
    mbar_wait(&bars[2], it & 1);
    {
      [[maybe_unused]] float cfl_min = 3.0e38f;
      [[maybe_unused]] float vt_f[8];
      if constexpr (CFL) {
        #pragma unroll
        for (int i = 0; i < 8; ++i) vt_f[i] = (float)late[C::lt_vtss + i];
      }
      for (int q = t; q < nq; q += C::threads) {
        // update*tss/nom/det (reference Spatial.hpp:484-487) with one division instead of two (<= 1 ulp)
        double mult;
        const double tss_q = a.tss ? late[C::lt_tss + q] : 1.;
        if constexpr (DEF) mult = update*tss_q/(nom*late[C::lt_det + q]);
        else mult = update*tss_q/nom;
        [[maybe_unused]] double x[nv];
        #pragma unroll
        for (int v = 0; v < nv; ++v) {
          double u = R[(0*nv + v)*nq + q];
          u += R[(1*nv + v)*nq + q];
          u += R[(2*nv + v)*nq + q];
          double* cache = a.cache + ((size_t)e*C::cs + v)*nq + q;
          if (a.stage) u -= late[C::lt_cache + v*nq + q];
          else if (!a.compute_residual) *cache = u;
          u *= mult;
          if (a.compute_residual) *cache = u;
          else {
            const double xv = S[v*nq + q] + u;
            S[v*nq + q] = xv;
            a.state[((size_t)e*nv + v)*nq + q] = xv;
            if (a.record) bad |= (isfinite(xv) ? 0 : 2) | ((v >= ND && !(xv > 0.)) ? 1 : 0);
            if constexpr (CFL) x[v] = xv;
          }
        }
        if constexpr (CFL) {
          /* CFL screen in single precision while the new state is in registers: spacing/char_speed (reference Spatial.hpp:808-822)
           * to ~1e-6 relative with MUFU-assisted float instructions. The next max_dt_euler reduces these and re-evaluates in
           * FP64 only the elements within 1e-5 of the global minimum (misc_kernels.cu), so the time step itself is exact. A value
           * that is not a positive finite float is stored as 0 = "always re-evaluate this element". */
          const float rho = (float)x[ND], en = (float)x[ND + 1];
          const float m2 = (float)(x[0]*x[0] + x[1]*x[1] + x[2]*x[2]);
          const float inv = __fdividef(1.f, rho);
          const float a2 = 0.56f*en*inv;
          const float cs = a2*rsqrtf(a2) + m2*rsqrtf(m2 + 1e-30f)*inv; // sqrt(x) = x*rsqrt(x): one MUFU each
          float sv[8];
          #pragma unroll
          for (int i = 0; i < 8; ++i) sv[i] = vt_f[i];
          int str = 8;
          #pragma unroll
          for (int d = 0; d < ND; ++d) {
            const float coord = a.nodef[(q/ipow(RS, ND - 1 - d)) % RS];
            str /= 2;
            #pragma unroll
            for (int i = 0; i < 4; ++i) if (i < str) sv[i] += coord*(sv[i + str] - sv[i]);
          }
          float r = __fdividef(sv[0], cs);
          if (!(r > 0.f && r < 3.0e38f)) r = 0.f;
          cfl_min = fminf(cfl_min, r);
        }
      }
      if constexpr (CFL) {
        #pragma unroll
        for (int off = 16; off > 0; off /= 2) cfl_min = fminf(cfl_min, __shfl_xor_sync(0xffffffffu, cfl_min, off));
        if (t % 32 == 0) warp_cfl[t/32] = cfl_min;
      }
    }
    }
    __syncthreads(); // new state complete in S; late buffer free
    if constexpr (CFL) {
      if (t == 0) {
        float m = warp_cfl[0];
        #pragma unroll
        for (int i = 1; i < C::threads/32; ++i) m = fminf(m, warp_cfl[i]);
        a.cfl_approx[e] = m;
      }
    }
    if constexpr (!LEAN) {
      if (t == 0 && e + stride_e < a.elem_end) {
        fence_proxy_async();
        pipe_issue_late<RS, DEF, CFL>(a, e + stride_e, late, &bars[2]);
      }
    }

    /* ---- phase C: write_face from the updated state (reference Spatial.hpp:41-57) ---- */
    if constexpr (Map::vec2) {
      if (has_line) {
        double* fout = a.faces + (size_t)e*2*ND*nv*nfq;
        #pragma unroll
        for (int v = 0; v < nv; ++v) {
          double x[RS];
          if (vec) {
            #pragma unroll
            for (int k = 0; k + 1 < RS; k += 2) ld2(S + v*nq + q0 + k, x[k], x[k + 1]);
          } else {
            #pragma unroll
            for (int k = 0; k < RS; ++k) x[k] = S[v*nq + q0 + k*stride];
          }
          double e0, e1;
          face_extrap_eo<RS>(ops, x, e0, e1);
          fout[((2*d)*nv + v)*nfq + l] = e0;
          fout[((2*d + 1)*nv + v)*nfq + l] = e1;
          if (a.record) bad |= ((isfinite(e0) && isfinite(e1)) ? 0 : 2) | ((v >= ND && !(e0 > 0. && e1 > 0.)) ? 1 : 0);
        }
      }
    } else {
      if (has_line) {
        double* fout = a.faces + (size_t)e*2*ND*nv*nfq;
        #pragma unroll
        for (int v = 0; v < nv; ++v) {
          double x[RS];
          #pragma unroll
          for (int k = 0; k < RS; ++k) x[k] = S[v*nq + q0 + k*stride];
          double e0, e1;
          face_extrap_eo<RS>(ops, x, e0, e1);
          fout[((2*d)*nv + v)*nfq + l] = e0;
          fout[((2*d + 1)*nv + v)*nfq + l] = e1;
          if (a.record) bad |= ((isfinite(e0) && isfinite(e1)) ? 0 : 2) | ((v >= ND && !(e0 > 0. && e1 > 0.)) ? 1 : 0);
        }
      }
    }
    if (a.record) { // uniform across the CTA; the two votes also are the barrier that frees stage buffer s
      const int inadmissible = __syncthreads_or(bad & 1), nonfinite = __syncthreads_or(bad & 2);
      if (t == 0) a.record[e] = (inadmissible ? 1 : 0) | (nonfinite ? 2 : 0);
    } else if (a.end_barrier) __syncthreads();
    else {
      // Stage buffer s is handed back without stopping the CTA: every thread arrives on the "consumed" mbarrier when its phase C is done
      // and goes straight on to the next element (whose inputs sit in the other stage buffer; R, the late buffer and the scalars are
      // protected by the barriers after phases A and B); only the thread that refills the buffer waits for all arrivals.
      // (10 % of the kernel's stall samples were barrier waits, a third of them here, profiles/r02t_ncu_full_euler_car.md -- and yet the step
      // time did not move: 25.91 against 25.96 ms Cartesian, 30.50 against 30.57 ms deformed on one box, visit r02v. Kept as an option.)
      mbar_arrive(&bars[3]);
      if (t == 0 && e + 2*stride_e < a.elem_end) mbar_wait(&bars[3], it & 1);
    }
    if (t == 0 && e + 2*stride_e < a.elem_end) {
      fence_proxy_async();
      if constexpr (LEAN) pipe_issue_state<RS, DEF>(a, e + 2*stride_e, stage_buf, &bars[s]);
      else pipe_issue_stage<RS, DEF>(a, e + 2*stride_e, stage_buf, &bars[s]);
    }
  }
}

template <int RS, bool DEF, bool CFL, bool LEAN = false>
__global__ void __launch_bounds__(PipeCfg<RS, DEF>::threads)
local_euler_pipe_kernel(PipeArgs a, Ops ops) { pipe_body<RS, DEF, CFL, LEAN>(a, ops); }

/* Cartesian elements in the lean layout need 52 KB: four CTAs fit one SM if the kernel keeps to 128 registers */
template <int RS>
__global__ void __launch_bounds__(PipeCfg<RS, false>::threads, 4)
local_euler_pipe4_kernel(PipeArgs a, Ops ops) { pipe_body<RS, false, false, true>(a, ops); }

template <int RS, bool DEF, bool CFL, bool LEAN = false>
static int launch_pipe(hexed_b200_ctx* c, const PipeArgs& a)
{
  using C = PipeCfg<RS, DEF, LEAN>;
  void (*k)(PipeArgs, Ops) = local_euler_pipe_kernel<RS, DEF, CFL, LEAN>;
  if constexpr (LEAN && !DEF) k = local_euler_pipe4_kernel<RS>;
  // function attributes and occupancy are per DEVICE (one host process may drive several: hexed_b200_group_*): cached per instantiation and device
  static int blocks_per_sm_of [64] = {};
  int& blocks_per_sm = blocks_per_sm_of[c->device & 63];
  if (!blocks_per_sm) {
    HB_CUDA(c, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::smem_bytes));
    HB_CUDA(c, cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    int n = 0;
    HB_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k, C::threads, C::smem_bytes));
    if (n < 1) return fail(c, HEXED_B200_CUDA_ERROR, "pipelined local kernel does not fit on this device");
    blocks_per_sm = n;
  }
  int sms = 0;
  HB_CUDA(c, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device));
  const int n_elem = a.elem_end - a.elem_begin;
  int grid = sms*blocks_per_sm;
  if (grid > n_elem) grid = n_elem;
  HB_LAUNCH(k, grid, C::threads, C::smem_bytes, c->stream, a, c->ops);
  HB_CUDA(c, cudaGetLastError());
  return 0;
}

/* returns -1 if this (n_dim, row_size, options) combination is not covered and the caller should use the general kernel */
int launch_local_euler_pipe(hexed_b200_ctx* c, int deformed, hexed_b200_options o, int begin, int end)
{
  if (c->nd != 3 || (c->rs != 4 && c->rs != 6) || o.use_filter || !c->use_pipe || !c->ops_symmetric) return -1;
  PipeArgs a;
  a.state = c->state; a.tss = c->tss_is_one ? nullptr : c->tss; a.cache = c->cache; a.nom = c->nom; a.refn = c->refn; a.det = c->det; a.faces = c->face_state;
  a.elem_begin = begin; a.elem_end = end; a.n_car = c->n_car;
  a.update = o.i_stage ? o.dt*(.5/c->quad_safety) : o.dt;
  a.dt_dev = c->dt_dev_active;
  a.end_barrier = c->pipe_end_barrier;
  a.stage = o.i_stage != 0; a.compute_residual = o.compute_residual;
  a.vtss = c->vtss; a.cfl_approx = nullptr;
  a.record = nullptr;
  c->admis_valid[deformed ? 1 : 0] = false;
  const bool leave_admis = c->use_fused_admis && !a.compute_residual;
  if (leave_admis) {
    if (!c->record) HB_CUDA(c, cudaMalloc(&c->record, sizeof(int)*(c->n_elem ? c->n_elem : 1)));
    a.record = c->record;
  }
  for (int i = 0; i < MAX_RS; ++i) a.nodef[i] = (float)c->ops.node[i];
  c->cfl_valid[deformed ? 1 : 0] = false; // this launch rewrites the state of the set
  const bool leave_cfl = c->use_cfl_cache && a.stage && !a.compute_residual;
  if (leave_cfl) {
    if (!c->cfl_approx) HB_CUDA(c, cudaMalloc(&c->cfl_approx, sizeof(float)*c->n_elem));
    a.cfl_approx = c->cfl_approx;
  }
  int rc;
  if (leave_cfl) {
    if (c->rs == 6) rc = deformed ? launch_pipe<6, true, true>(c, a) : launch_pipe<6, false, true>(c, a);
    else rc = deformed ? launch_pipe<4, true, true>(c, a) : launch_pipe<4, false, true>(c, a);
  }
  else if (deformed && c->pipe_lean) rc = c->rs == 6 ? launch_pipe<6, true, false, true>(c, a) : launch_pipe<4, true, false, true>(c, a);
  else if (!deformed && c->pipe_lean4) rc = c->rs == 6 ? launch_pipe<6, false, false, true>(c, a) : launch_pipe<4, false, false, true>(c, a);
  else if (c->rs == 6) rc = deformed ? launch_pipe<6, true, false>(c, a) : launch_pipe<6, false, false>(c, a);
  else rc = deformed ? launch_pipe<4, true, false>(c, a) : launch_pipe<4, false, false>(c, a);
  if (rc == 0 && leave_cfl) c->cfl_valid[deformed ? 1 : 0] = true;
  if (rc == 0 && leave_admis) c->admis_valid[deformed ? 1 : 0] = true;
  return rc;
}

} // namespace hb
