/* local_euler_pipe2d.cu -- persistent, TMA-pipelined Euler `Local` kernel for 2-D elements (the vortex / naca0012 / cylinder
 * configurations of BASELINE.json are 2-D, row size 6).
 *
 * Same arithmetic as local_euler.cu (reference include/Spatial.hpp:326-509 + the trailing write_face :41-57,507) and the same
 * organisation as the 3-D kernel of local_euler_pipe.cu, with one difference forced by the element size: a 2-D element is only
 * row_size^2 points (36 at row size 6, 1.1 KB of state), so the unit of work of a persistent CTA is a BATCH of B consecutive
 * elements. Because the layout is element-major, the state / numerical-flux faces / reference normals / time-step scale /
 * determinant of a batch are each ONE contiguous run: a batch is fetched with a handful of 1-D bulk TMA copies
 * (cp.async.bulk -> UBLKCP) onto mbarriers, double buffered two batches ahead (late inputs one phase ahead), exactly like one
 * 3-D element. The point-per-thread kernel it replaces reached 47 % (Cartesian) / 62 % (deformed) of the measured HBM bandwidth
 * (profiles/r01g_bench_lines.jsonl): 144-thread CTAs with per-thread global loads between barriers.
 *
 * Shared memory holds only what is reused: the state (double buffered, it is read in phases A, B and C), the numerical-flux faces and
 * reference normals (single buffered: dead after phase A, so the copies of the NEXT batch are issued right after it) and R. The
 * per-point late inputs (time-step scale, determinant, residual cache) are read once, straight from HBM after an L2 prefetch issued a
 * phase earlier. 65 KB per CTA at row size 6 instead of 103 KB: three resident CTAs per SM instead of two (the kernel is latency-bound:
 * 12 % occupancy, 27 % issue utilisation in profiles/r01l_ncu_full_2d.md).
 *
 * Shared-memory banks (profiles/r01o_ncu_full_2d.md: with three CTAs resident the kernel became LSU-bound, l1tex 88 % busy, 56 % of the
 * shared wavefronts bank conflicts): the per-element strides of every staged array are padded by row_size doubles (so a batch is
 * fetched with one bulk copy per element and array instead of one per array), warps are homogeneous in the line direction, the lines
 * that are contiguous in memory (dimension 1) are moved with 128-bit accesses, and the two directions enumerate their tasks
 * differently (dimension 0 element-major, dimension 1 line-major) so that a half / quarter warp touches distinct banks.
 *
 * Phases per batch: A (line tasks (element, dimension, line): pointwise flux on the line, derivative + lifted face flux -> R_d),
 * B (point tasks: r = R_0 + R_1, two-stage update, new state -> HBM and in place in shared memory),
 * C (line tasks: extrapolate the new state to both ends of the line -> HBM).
 */
#include "euler.cuh"

namespace hb {

template <int RS, bool DEF>
struct Pipe2Cfg
{
  static constexpr int ND = 2, nq = RS*RS, nfq = RS, nv = 4;
  static constexpr int lines_per_elem = ND*nfq;
  static constexpr int B = 128/lines_per_elem;            // elements per batch: 10 at row size 6
  static constexpr int n_line = B*lines_per_elem;
  static constexpr int threads = ((n_line + 31)/32)*32;
  static constexpr int cs = nv > RS ? nv : RS;            // slots per element of the residual cache array
  // per-element doubles of each staged array
  static constexpr int e_state = nv*nq, e_face = 2*ND*nv*nfq, e_nrml = DEF ? ND*ND*nq : 0;
  // per-element strides in shared memory: padded by row_size doubles (16-byte multiples are kept: row_size is even)
  static constexpr int pad = RS;
  static constexpr int p_state = e_state + pad, p_face = e_face + pad, p_nrml = DEF ? e_nrml + pad : 0, p_r = ND*nv*nq + pad;
  static constexpr int half = threads/2;                                // threads [0, half): lines in dimension 0, [half, threads): dimension 1
  static_assert(RS % 2 == 0 && B*nfq <= half && B <= 32, "task layout");
  static constexpr int state_doubles = B*p_state;                       // one of the two state buffers
  static constexpr int fn_face = 0, fn_nrml = B*p_face, fn_doubles = B*(p_face + p_nrml); // faces | normals, single buffered
  static constexpr int r_doubles = B*p_r;
  static constexpr int smem_doubles = 2*state_doubles + fn_doubles + r_doubles;
  static constexpr int n_iter = (B*nq + threads - 1)/threads;           // point tasks per thread and batch
  static constexpr size_t smem_bytes = sizeof(double)*smem_doubles + 4*sizeof(mbar_t) + 16*sizeof(int); // + per-element admissibility bits of the batch
};

struct Pipe2Args
{
  double* state; const double* tss; double* cache; const double* nom; const double* refn; const double* det; double* faces;
  int elem_begin, elem_end, n_car;
  double update; int stage; int compute_residual;
  const double* dt_dev; // non-null: the time step lives on the device and multiplies `update`
  int* record; // non-null: admissibility bits of the new state and faces per element (see local_euler_pipe.cu)
};

/* both issue functions are called by every lane of warp 0: lane 0 posts the byte count, lane i < n copies element i of the batch */
template <int RS, bool DEF>
__device__ __forceinline__ void pipe2_issue_state(const Pipe2Args& a, int e0, int n, double* buf, mbar_t* bar, int lane)
{
  using C = Pipe2Cfg<RS, DEF>;
  constexpr unsigned b_state = sizeof(double)*C::e_state;
  if (lane == 0) mbar_arrive_expect_tx(bar, b_state*n);
  __syncwarp();
  if (lane < n) bulk_g2s(buf + lane*C::p_state, a.state + (size_t)(e0 + lane)*C::e_state, b_state, bar);
}

template <int RS, bool DEF>
__device__ __forceinline__ void pipe2_issue_fn(const Pipe2Args& a, int e0, int n, double* buf, mbar_t* bar, int lane)
{
  using C = Pipe2Cfg<RS, DEF>;
  constexpr unsigned b_face = sizeof(double)*C::e_face, b_nrml = sizeof(double)*C::e_nrml;
  if (lane == 0) mbar_arrive_expect_tx(bar, (b_face + b_nrml)*n);
  __syncwarp();
  if (lane < n) {
    bulk_g2s(buf + C::fn_face + lane*C::p_face, a.faces + (size_t)(e0 + lane)*C::e_face, b_face, bar);
    if constexpr (DEF) bulk_g2s(buf + C::fn_nrml + lane*C::p_nrml, a.refn + (size_t)(e0 + lane - a.n_car)*C::e_nrml, b_nrml, bar);
  }
}

/* the late inputs of batch e0 towards L2: one 128-byte line per call */
template <int RS, bool DEF>
__device__ __forceinline__ void pipe2_prefetch_late(const Pipe2Args& a, int e0, int n, int t, int n_threads)
{
  using C = Pipe2Cfg<RS, DEF>;
  for (int i = t*16; i < n*C::nq; i += n_threads*16) {
    if (a.tss) prefetch_l2(a.tss + (size_t)e0*C::nq + i);
    if constexpr (DEF) prefetch_l2(a.det + (size_t)(e0 - a.n_car)*C::nq + i);
  }
  if (a.stage) { // the cache array keeps max(nv, row_size) slots per element; only the first nv are read
    constexpr int lines = (C::e_state + 15)/16;
    for (int i = t; i < n*lines; i += n_threads) prefetch_l2(a.cache + ((size_t)(e0 + i/lines)*C::cs)*C::nq + (i % lines)*16);
  }
}

template <int RS, bool DEF>
__global__ void __launch_bounds__(Pipe2Cfg<RS, DEF>::threads)
local_euler_pipe2d_kernel(Pipe2Args a, Ops ops)
{
  using C = Pipe2Cfg<RS, DEF>;
  constexpr int ND = 2, nq = C::nq, nfq = C::nfq, nv = C::nv, B = C::B;
  HB_DYN_SMEM(double, smem);
  double* FN = smem + 2*C::state_doubles;
  double* R = FN + C::fn_doubles;
  mbar_t* bars = reinterpret_cast<mbar_t*>(R + C::r_doubles); // [0],[1]: state buffers; [2]: faces + normals
  int* s_bad = reinterpret_cast<int*>(bars + 4); // [B] admissibility bits of the batch's elements
  static_assert(B <= 16, "s_bad holds 16 entries");
  const int t = threadIdx.x;
  const int stride_e = gridDim.x*B;
  int e0 = a.elem_begin + blockIdx.x*B;
  if (e0 >= a.elem_end) return;
  auto count = [&](int first) { const int n = a.elem_end - first; return n < B ? n : B; };

  if (t == 0) {
    mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_init(&bars[2], 1);
    mbar_init_fence();
  }
  __syncthreads();
  if (t < 32) {
    pipe2_issue_state<RS, DEF>(a, e0, count(e0), smem, &bars[0], t);
    pipe2_issue_fn<RS, DEF>(a, e0, count(e0), FN, &bars[2], t);
    if (e0 + stride_e < a.elem_end) pipe2_issue_state<RS, DEF>(a, e0 + stride_e, count(e0 + stride_e), smem + C::state_doubles, &bars[1], t);
  }
  pipe2_prefetch_late<RS, DEF>(a, e0, count(e0), t, C::threads);

  // line task of this thread: element le of the batch, dimension d (uniform per warp), line l; points q0 + k*(d == 0 ? row_size : 1).
  // dimension 0 enumerates element-major (consecutive threads: consecutive doubles), dimension 1 line-major (consecutive threads:
  // the same line of consecutive elements, 16-byte chunks 3*(le + l) mod 8 apart at row size 6)
  const int d = t >= C::half, ti = t - d*C::half;
  const int le = d == 0 ? ti/nfq : ti % B, l = d == 0 ? ti % nfq : ti/B;
  const int q0 = d == 0 ? l : l*RS;
  // points k, k + 1 (k even) of this thread's line in the field starting at `field`
  auto load_pair = [&](const double* field, int k, double& x, double& y) {
    if (d == 0) { x = field[q0 + k*RS]; y = field[q0 + (k + 1)*RS]; }
    else ld2(field + q0 + k, x, y);
  };
  auto store_pair = [&](double* field, int k, double x, double y) {
    if (d == 0) { field[q0 + k*RS] = x; field[q0 + (k + 1)*RS] = y; }
    else st2(field + q0 + k, x, y);
  };

  for (int it = 0; e0 < a.elem_end; ++it, e0 += stride_e) {
    const int s = it & 1;
    const unsigned par = (it >> 1) & 1;
    const int n = count(e0);
    const bool has_line = ti < B*nfq && le < n;
    double* S = smem + s*C::state_doubles;
    const double* F = FN + C::fn_face;
    const double* N = FN + C::fn_nrml;
    mbar_wait(&bars[s], par);
    mbar_wait(&bars[2], it & 1);
    if (a.record && t < B) s_bad[t] = 0; // ordered before the first atomicOr by the barrier after phase A

    /* ---- phase A: flux on the line, then D(flux, face flux) -> R_d ---- */
    if (has_line) {
      const double* Se = S + le*C::p_state;
      double f[nv][RS];
      #pragma unroll
      for (int k = 0; k < RS; k += 2) {
        EulerPoint<ND> p[2];
        #pragma unroll
        for (int v = 0; v < nv; ++v) load_pair(Se + v*nq, k, p[0].s[v], p[1].s[v]);
        [[maybe_unused]] double nr[2][ND];
        if constexpr (DEF) {
          #pragma unroll
          for (int j = 0; j < ND; ++j) load_pair(N + le*C::p_nrml + (d*ND + j)*nq, k, nr[0][j], nr[1][j]);
        }
        #pragma unroll
        for (int h = 0; h < 2; ++h) {
          p[h].scalars();
          double fl[nv];
          if constexpr (DEF) p[h].flux(nr[h], fl);
          else {
            const double mass_flux = d == 0 ? p[h].s[0] : p[h].s[1];
            const double vol_flux = mass_flux*p[h].inv_mass;
            fl[ND] = mass_flux;
            fl[ND + 1] = (p[h].s[ND + 1] + p[h].pressure)*vol_flux;
            #pragma unroll
            for (int j = 0; j < ND; ++j) fl[j] = p[h].s[j]*vol_flux + (j == d ? p[h].pressure : 0.);
          }
          #pragma unroll
          for (int v = 0; v < nv; ++v) f[v][k + h] = fl[v];
        }
      }
      const double* Fe = F + le*C::p_face;
      double* Re = R + le*C::p_r;
      #pragma unroll
      for (int v = 0; v < nv; ++v) {
        const double b0 = Fe[((2*d)*nv + v)*nfq + l], b1 = Fe[((2*d + 1)*nv + v)*nfq + l];
        double r[RS];
        line_deriv_eo<RS, true>(ops, f[v], b0, b1, r);
        #pragma unroll
        for (int i = 0; i < RS; i += 2) store_pair(Re + (d*nv + v)*nq, i, r[i], r[i + 1]);
      }
    }
    // every late input of this thread's points is loaded before the barrier that ends phase A (nothing phase A computes is needed for them:
    // their latency overlaps the wait for the slowest warp) and before the thread's first store (a store could alias a later load and would
    // chain the memory round trips)
    double l_tss[C::n_iter], l_cache[C::n_iter][nv];
    [[maybe_unused]] double l_det[C::n_iter];
    double l_nom[C::n_iter];
    #pragma unroll
    for (int k = 0; k < C::n_iter; ++k) {
      const int pt = t + k*C::threads;
      if (pt < n*nq) {
        const int pe = pt/nq, q = pt % nq;
        const int e = e0 + pe;
        l_tss[k] = a.tss ? a.tss[(size_t)e*nq + q] : 1.;
        l_nom[k] = a.nom[e];
        if constexpr (DEF) l_det[k] = a.det[(size_t)(e - a.n_car)*nq + q];
        if (a.stage) {
          #pragma unroll
          for (int v = 0; v < nv; ++v) l_cache[k][v] = a.cache[((size_t)e*C::cs + v)*nq + q];
        }
      }
    }
    __syncthreads(); // R complete; the faces / normals buffer is dead, the state is still needed
    if (t < 32 && e0 + stride_e < a.elem_end) {
      fence_proxy_async();
      pipe2_issue_fn<RS, DEF>(a, e0 + stride_e, count(e0 + stride_e), FN, &bars[2], t);
    }
    if (e0 + stride_e < a.elem_end) pipe2_prefetch_late<RS, DEF>(a, e0 + stride_e, count(e0 + stride_e), t, C::threads);

    /* ---- phase B: combine, two-stage update (reference Spatial.hpp:484-503) ---- */
    {
      const double update = a.dt_dev ? *a.dt_dev*a.update : a.update;
      #pragma unroll
      for (int k = 0; k < C::n_iter; ++k) {
        const int pt = t + k*C::threads;
        if (pt < n*nq) {
          const int pe = pt/nq, q = pt % nq;
          const int e = e0 + pe;
          double mult; // update*tss/nom/det with one division (<= 1 ulp)
          if constexpr (DEF) mult = update*l_tss[k]/(l_nom[k]*l_det[k]);
          else mult = update*l_tss[k]/l_nom[k];
          #pragma unroll
          for (int v = 0; v < nv; ++v) {
            double u = R[pe*C::p_r + (0*nv + v)*nq + q];
            u += R[pe*C::p_r + (1*nv + v)*nq + q];
            double* cache = a.cache + ((size_t)e*C::cs + v)*nq + q;
            if (a.stage) u -= l_cache[k][v];
            else if (!a.compute_residual) *cache = u;
            u *= mult;
            if (a.compute_residual) *cache = u;
            else {
              const double xv = S[pe*C::p_state + v*nq + q] + u;
              S[pe*C::p_state + v*nq + q] = xv;
              a.state[(size_t)e*C::e_state + v*nq + q] = xv;
              if (a.record) {
                const int bad = (isfinite(xv) ? 0 : 2) | ((v >= ND && !(xv > 0.)) ? 1 : 0);
                if (bad) atomicOr(&s_bad[pe], bad);
              }
            }
          }
        }
      }
    }
    __syncthreads(); // new state complete in S

    /* ---- phase C: write_face from the updated state (reference Spatial.hpp:41-57) ---- */
    if (has_line) {
      const double* Se = S + le*C::p_state;
      double* fout = a.faces + (size_t)(e0 + le)*C::e_face;
      #pragma unroll
      for (int v = 0; v < nv; ++v) {
        double x[RS];
        #pragma unroll
        for (int k = 0; k < RS; k += 2) load_pair(Se + v*nq, k, x[k], x[k + 1]);
        double x0, x1;
        face_extrap_eo<RS>(ops, x, x0, x1);
        fout[((2*d)*nv + v)*nfq + l] = x0;
        fout[((2*d + 1)*nv + v)*nfq + l] = x1;
        if (a.record) {
          const int bad = ((isfinite(x0) && isfinite(x1)) ? 0 : 2) | ((v >= ND && !(x0 > 0. && x1 > 0.)) ? 1 : 0);
          if (bad) atomicOr(&s_bad[le], bad);
        }
      }
    }
    __syncthreads(); // stage buffer s free
    if (a.record && t < n) a.record[e0 + t] = s_bad[t];
    if (t < 32 && e0 + 2*stride_e < a.elem_end) {
      fence_proxy_async();
      pipe2_issue_state<RS, DEF>(a, e0 + 2*stride_e, count(e0 + 2*stride_e), S, &bars[s], t);
    }
  }
}

template <int RS, bool DEF>
static int launch_pipe2(hexed_b200_ctx* c, const Pipe2Args& a)
{
  using C = Pipe2Cfg<RS, DEF>;
  auto k = local_euler_pipe2d_kernel<RS, DEF>;
  static int blocks_per_sm_of [64] = {}; // per instantiation and device (function attributes are per device)
  int& blocks_per_sm = blocks_per_sm_of[c->device & 63];
  if (!blocks_per_sm) {
    HB_CUDA(c, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::smem_bytes));
    int n = 0;
    HB_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k, C::threads, C::smem_bytes));
    if (n < 1) return fail(c, HEXED_B200_CUDA_ERROR, "pipelined 2-D local kernel does not fit on this device");
    blocks_per_sm = n;
  }
  int sms = 0;
  HB_CUDA(c, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device));
  const int n_batch = (a.elem_end - a.elem_begin + C::B - 1)/C::B;
  int grid = sms*blocks_per_sm;
  if (grid > n_batch) grid = n_batch;
  HB_LAUNCH(k, grid, C::threads, C::smem_bytes, c->stream, a, c->ops);
  HB_CUDA(c, cudaGetLastError());
  return 0;
}

/* returns -1 if this (n_dim, row_size, options) combination is not covered and the caller should use the general kernel.
 * Bulk copies need 16-byte multiples: row_size^2*8 bytes per field -> even row sizes. */
int launch_local_euler_pipe2d(hexed_b200_ctx* c, int deformed, hexed_b200_options o, int begin, int end)
{
  if (c->nd != 2 || (c->rs != 4 && c->rs != 6 && c->rs != 8) || o.use_filter || !c->use_pipe || !c->ops_symmetric) return -1;
  Pipe2Args a;
  a.state = c->state; a.tss = c->tss_is_one ? nullptr : c->tss /* null: time_step_scale known to hold 1, not read */; a.cache = c->cache; a.nom = c->nom; a.refn = c->refn; a.det = c->det; a.faces = c->face_state;
  a.elem_begin = begin; a.elem_end = end; a.n_car = c->n_car;
  a.update = o.i_stage ? o.dt*(.5/c->quad_safety) : o.dt;
  a.dt_dev = c->dt_dev_active;
  a.stage = o.i_stage != 0; a.compute_residual = o.compute_residual;
  c->cfl_valid[deformed ? 1 : 0] = false;
  a.record = nullptr;
  c->admis_valid[deformed ? 1 : 0] = false;
  const bool leave_admis = c->use_fused_admis && !a.compute_residual;
  if (leave_admis) {
    if (!c->record) HB_CUDA(c, cudaMalloc(&c->record, sizeof(int)*(c->n_elem ? c->n_elem : 1)));
    a.record = c->record;
  }
  int rc;
  if (c->rs == 6) rc = deformed ? launch_pipe2<6, true>(c, a) : launch_pipe2<6, false>(c, a);
  else if (c->rs == 4) rc = deformed ? launch_pipe2<4, true>(c, a) : launch_pipe2<4, false>(c, a);
  else rc = deformed ? launch_pipe2<8, true>(c, a) : launch_pipe2<8, false>(c, a);
  if (rc == 0 && leave_admis) c->admis_valid[deformed ? 1 : 0] = true;
  return rc;
}

} // namespace hb
