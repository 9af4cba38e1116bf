/* misc_kernels.cu -- CFL time step, hanging-node transfer, boundary ghost fill, face gather/scatter. */
#include "euler.cuh"
#include "characteristics.cuh"

namespace hb {

/* ---------------- Max_dt: reference include/Spatial.hpp:784-828, include/math.hpp:207-218 ----------------
 * One thread per quadrature point: n-linear interpolation of the vertex spacing, characteristic speed, local
 * time-step scale written to tss (or 1.) and, for global time stepping, a warp-shuffle + shared-memory block minimum
 * followed by a one-block final reduction. The Cartesian and deformed instantiations of the reference are the same
 * arithmetic, so one launch covers every element. */
struct MaxDtArgs
{
  const double* state; double* tss; const double* vtss; int n_elem; double max_cfl_c; int is_local; unsigned long long* global_min;
  int write_tss; // global time stepping: 0 = time_step_scale already holds 1 everywhere (ctx::tss_is_one), skip the 8 bytes per point
};

constexpr int max_dt_ppt = 1; // points per thread. Measured with 2 (both points' loads in flight before either is used; 70 registers, 3 CTAs
                              // of 256 per SM instead of 6): 3.0 ms against 2.54 ms at 1 M 3-D elements -- occupancy hides the load latency
                              // better than per-thread memory-level parallelism does; the arithmetic (two square roots, a reciprocal and a
                              // division per point, ~256 instructions) keeps the kernel at 4.1 TB/s (profiles/r01l_ncu_full_euler.md).
                              // Also measured: a single-precision screen with a block-wide minimum so that only the points within 2e-5 of it
                              // take the FP64 path (dt stays bit-identical): 2.97 ms -- the extra barrier serialises the block's loads.
template <int ND, int RS>
__global__ void __launch_bounds__(256)
max_dt_euler_kernel(MaxDtArgs a, Ops ops)
{
  constexpr int nq = ipow(RS, ND), nv = ND + 2, n_vert = ipow(2, ND), PPT = max_dt_ppt;
  __shared__ double warp_min[8];
  const long long first = (long long)blockIdx.x*(256*PPT) + threadIdx.x;
  double st[PPT][nv], vt[PPT][n_vert];
  int elem[PPT], pt[PPT];
  #pragma unroll
  for (int r = 0; r < PPT; ++r) {
    const long long gid = first + r*256;
    elem[r] = (int)(gid/nq); pt[r] = (int)(gid % nq);
    if (elem[r] < a.n_elem) {
      #pragma unroll
      for (int v = 0; v < nv; ++v) st[r][v] = a.state[((size_t)elem[r]*nv + v)*nq + pt[r]];
      #pragma unroll
      for (int i = 0; i < n_vert; ++i) vt[r][i] = a.vtss[(size_t)elem[r]*n_vert + i];
    }
  }
  double val = DBL_MAX;
  #pragma unroll
  for (int r = 0; r < PPT; ++r) {
    if (elem[r] < a.n_elem) {
      const double spacing = interp_vertex_spacing<ND, RS>(vt[r], ops, pt[r]);
      EulerPoint<ND> p;
      #pragma unroll
      for (int v = 0; v < nv; ++v) p.s[v] = st[r][v];
      // 1/scale with scale = char_speed/max_cfl/spacing (Spatial.hpp:808-822), rearranged to a single division
      const double local_dt = p.cfl_time_step(a.max_cfl_c*spacing);
      if (a.is_local) a.tss[(size_t)elem[r]*nq + pt[r]] = local_dt;
      else { if (a.write_tss) a.tss[(size_t)elem[r]*nq + pt[r]] = 1.; val = fmin(val, local_dt); }
    }
  }
  if (a.is_local) return;
  #pragma unroll
  for (int off = 16; off > 0; off /= 2) val = fmin(val, __shfl_xor_sync(0xffffffffu, val, off));
  if (threadIdx.x % 32 == 0) warp_min[threadIdx.x/32] = val;
  __syncthreads();
  if (threadIdx.x == 0) {
    double m = warp_min[0];
    for (int i = 1; i < (int)blockDim.x/32; ++i) m = fmin(m, warp_min[i]);
    // time steps are positive, so the IEEE bit patterns order like unsigned integers
    atomicMin(a.global_min, (unsigned long long)__double_as_longlong(m));
  }
}

/* Global time stepping with a RUNNING single-precision screen (no barrier, bit-identical result). max_dt_euler_kernel issues ~257
 * instructions per point (two FP64 square roots and a division) and is issue-bound at 3.9 TB/s (profiles/r01q_ncu_full_maxdt.md), yet
 * only the point that holds the minimum matters. Here every point evaluates the time step in single precision (good to ~1e-6),
 * a warp takes the minimum with shuffles and lane 0 folds it into a device-wide running minimum (read through L1: a stale value only
 * makes the screen less sharp; the atomic is issued only when it lowers the value). Points within 2e-5 of min(running, warp) -- and
 * points whose screen value is not a positive finite float -- are evaluated with the exact FP64 arithmetic and folded into the result
 * with atomicMin (again guarded by a cached read). The point holding the true minimum always passes the screen (its screen value is
 * within 2e-6 of the smallest screen value there will ever be), so the result is the same double as the full reduction's.
 * A uniform flow degenerates to screen + exact for every point (~1.2x the plain kernel). */
constexpr bool max_dt_running_screen = true;

template <int ND, int RS>
__global__ void __launch_bounds__(256)
max_dt_euler_screen_kernel(MaxDtArgs a, Ops ops, int* screen_bits)
{
  constexpr int nq = ipow(RS, ND), nv = ND + 2, n_vert = ipow(2, ND);
  const unsigned gid = blockIdx.x*256u + threadIdx.x; // the launcher checks that the point count fits 32 bits (64-bit division by nq
  const unsigned elem = gid/nq, pt = gid - elem*nq;    // is ~15 instructions and this kernel is issue-bound)
  const bool valid = elem < (unsigned)a.n_elem;
  EulerPoint<ND> p;
  double c_spacing = 0;
  float ap = 3.0e38f; // screen value; 0 = not representable, evaluate exactly
  if (valid) {
    const double* sp = a.state + (size_t)elem*(nv*nq) + pt;
    #pragma unroll
    for (int v = 0; v < nv; ++v) p.s[v] = sp[v*nq];
    c_spacing = a.max_cfl_c*interp_vertex_spacing<ND, RS>(a.vtss + (size_t)elem*n_vert, ops, pt);
    if (a.write_tss) a.tss[(size_t)elem*nq + pt] = 1.; // Spatial.hpp:823-825
    const float rho = (float)p.s[ND], en = (float)p.s[ND + 1];
    float sq = 0;
    #pragma unroll
    for (int i = 0; i < ND; ++i) { const float m = (float)p.s[i]; sq += m*m; }
    ap = __fdividef((float)c_spacing*rho, __fsqrt_rn((float)(heat_rat*(heat_rat - 1))*en*rho) + __fsqrt_rn(sq));
    if (!(ap > 0.f && ap < 3.0e38f)) ap = 0.f;
  }
  float wm = ap > 0.f ? ap : 3.0e38f;
  #pragma unroll
  for (int off = 16; off > 0; off /= 2) wm = fminf(wm, __shfl_xor_sync(0xffffffffu, wm, off));
  float g = 3.0e38f;
  if (threadIdx.x % 32 == 0) {
    g = __int_as_float(__ldca(screen_bits));
    if (wm < g) { atomicMin(screen_bits, __float_as_int(wm)); g = wm; } // positive floats order like their bit patterns
  }
  g = __shfl_sync(0xffffffffu, g, 0);
  if (valid && (ap == 0.f || ap <= g*1.00002f)) {
    const double exact = p.cfl_time_step(c_spacing);
    const unsigned long long bits = (unsigned long long)__double_as_longlong(exact);
    // time steps are positive, so the IEEE bit patterns order like unsigned integers (a NaN or negative value never lowers the minimum
    // of the plain kernel either: fmin drops NaN; negative values are an inadmissible state and the caller's problem in both)
    if (exact > 0. && bits < __ldca(a.global_min)) atomicMin(a.global_min, bits);
  }
}

/* Global time step from the single-precision CFL screen that the stage-1 Local kernel leaves behind (local_euler_pipe.cu):
 *   1. cfl_screen_min_kernel: minimum of the positive screen values (bit pattern of a positive float orders like an int);
 *   2. cfl_exact_kernel: one warp per element; an element whose screen value is 0 (not representable) or within 1e-5 of that
 *      minimum (the screen is good to ~1e-6) is re-evaluated from its state with the exact FP64 arithmetic of max_dt_euler_kernel.
 * Typically a handful of elements are touched instead of the whole state; a uniform flow degenerates to the full pass. */
__global__ void __launch_bounds__(256)
cfl_screen_min_kernel(const float* approx, int n, int* global_min_bits)
{
  __shared__ float warp_min[8];
  float val = 3.0e38f;
  for (long long i = (long long)blockIdx.x*blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x*blockDim.x) {
    const float a = approx[i];
    if (a > 0.f) val = fminf(val, a);
  }
  #pragma unroll
  for (int off = 16; off > 0; off /= 2) val = fminf(val, __shfl_xor_sync(0xffffffffu, val, off));
  if (threadIdx.x % 32 == 0) warp_min[threadIdx.x/32] = val;
  __syncthreads();
  if (threadIdx.x == 0) {
    float m = warp_min[0];
    for (int i = 1; i < (int)blockDim.x/32; ++i) m = fminf(m, warp_min[i]);
    atomicMin(global_min_bits, __float_as_int(m));
  }
}

template <int ND, int RS>
__global__ void __launch_bounds__(256)
cfl_exact_kernel(MaxDtArgs a, Ops ops, const float* approx, const int* screen_min_bits)
{
  constexpr int nq = ipow(RS, ND), nv = ND + 2, n_vert = ipow(2, ND);
  const float thr = __int_as_float(*screen_min_bits)*1.00001f;
  const int lane = threadIdx.x % 32;
  const long long warp0 = ((long long)blockIdx.x*blockDim.x + threadIdx.x)/32, n_warp = (long long)gridDim.x*blockDim.x/32;
  double val = DBL_MAX;
  for (long long e = warp0; e < a.n_elem; e += n_warp) {
    const float ap = approx[e];
    if (ap > thr) continue;
    for (int q = lane; q < nq; q += 32) {
      const double spacing = interp_vertex_spacing<ND, RS>(a.vtss + (size_t)e*n_vert, ops, q);
      EulerPoint<ND> p;
      #pragma unroll
      for (int v = 0; v < nv; ++v) p.s[v] = a.state[((size_t)e*nv + v)*nq + q];
      val = fmin(val, p.cfl_time_step(a.max_cfl_c*spacing));
    }
  }
  #pragma unroll
  for (int off = 16; off > 0; off /= 2) val = fmin(val, __shfl_xor_sync(0xffffffffu, val, off));
  if (lane == 0 && val < DBL_MAX) atomicMin(a.global_min, (unsigned long long)__double_as_longlong(val));
}

__global__ void __launch_bounds__(256)
fill_kernel(double* dst, long long n, double value)
{
  for (long long i = (long long)blockIdx.x*blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x*blockDim.x) dst[i] = value;
}

static int max_dt_from_cfl_screen(hexed_b200_ctx* c, double max_cfl_c, double* dt)
{
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
  if (!c->tss_is_one) { // the reference's Max_dt writes time_step_scale = 1 for global time stepping (Spatial.hpp:823-825)
    HB_LAUNCH(fill_kernel, sms*8, 256, 0, c->stream, c->tss, (long long)c->n_elem*c->nq, 1.);
    count_launch(c, ST_MAX_DT_CAR);
  }
  HB_CUDA(c, cudaMemsetAsync(c->d_scalar, 0x7f, 2*sizeof(double), c->stream)); // slot 0: exact minimum (1.4e306), slot 1: screen minimum (3.4e38f)
  int* screen_bits = reinterpret_cast<int*>(c->d_scalar + 1);
  int grid = (c->n_elem + 255)/256;
  if (grid > sms*8) grid = sms*8;
  HB_LAUNCH(cfl_screen_min_kernel, grid, 256, 0, c->stream, c->cfl_approx, c->n_elem, screen_bits);
  count_launch(c, ST_MAX_DT_CAR);
  int rc = dispatch(c, [&](auto nd, auto rs) {
    constexpr int ND = decltype(nd)::value, RS = decltype(rs)::value;
    MaxDtArgs a;
    a.state = c->state; a.tss = c->tss; a.vtss = c->vtss; a.n_elem = c->n_elem; a.max_cfl_c = max_cfl_c; a.is_local = 0;
    a.global_min = reinterpret_cast<unsigned long long*>(c->d_scalar);
    long long blocks = ((long long)c->n_elem*32 + 255)/256;
    if (blocks > sms*16) blocks = sms*16;
    auto k = cfl_exact_kernel<ND, RS>;
    HB_LAUNCH(k, (int)blocks, 256, 0, c->stream, a, c->ops, (const float*)c->cfl_approx, (const int*)screen_bits);
    return 0;
  });
  if (rc) return rc;
  count_launch(c, ST_MAX_DT_CAR);
  HB_CUDA(c, cudaGetLastError());
  HB_CUDA(c, cudaMemcpyAsync(c->h_scalar, c->d_scalar, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  HB_CUDA(c, cudaStreamSynchronize(c->stream));
  *dt = *c->h_scalar;
  c->tss_is_one = true;
  return 0;
}

int launch_max_dt_euler(hexed_b200_ctx* c, double safety_conv, int local_time, double* dt)
{
  if (!c->have_mesh) return fail(c, HEXED_B200_NO_MESH, "no mesh uploaded");
  StatScope s_car(c, ST_MAX_DT_CAR, c->n_car);
  c->stats[ST_MAX_DT_DEF].work_units += c->n_def;
  if (!c->n_elem) { *dt = local_time ? 1. : DBL_MAX; return 0; }
  if (!local_time && c->use_cfl_cache && c->cfl_approx && (c->n_car == 0 || c->cfl_valid[0]) && (c->n_def == 0 || c->cfl_valid[1]))
    return max_dt_from_cfl_screen(c, (-2*c->quad_safety/c->min_eig_conv)*safety_conv, dt);
  return dispatch(c, [&](auto nd, auto rs) {
    constexpr int ND = decltype(nd)::value, RS = decltype(rs)::value;
    const long long total = (long long)c->n_elem*ipow(RS, ND);
    const int grid = (int)((total + 256*max_dt_ppt - 1)/(256*max_dt_ppt));
    MaxDtArgs a;
    a.state = c->state; a.tss = c->tss; a.vtss = c->vtss; a.n_elem = c->n_elem;
    a.max_cfl_c = (-2*c->quad_safety/c->min_eig_conv)*safety_conv; // Basis::max_cfl (src/Basis.cpp:6-9) * safety (Spatial.hpp:777)
    a.is_local = local_time; a.global_min = reinterpret_cast<unsigned long long*>(c->d_scalar);
    a.write_tss = !c->tss_is_one;
    if (!local_time) HB_CUDA(c, cudaMemsetAsync(c->d_scalar, 0x7f, sizeof(double), c->stream)); // 0x7f7f... = 1.4e306
    if (!local_time && max_dt_running_screen && total < (1ll << 32) - 256) {
      int* screen_bits = reinterpret_cast<int*>(c->d_scalar + 1);
      HB_CUDA(c, cudaMemsetAsync(screen_bits, 0x7f, sizeof(int), c->stream)); // 0x7f7f7f7f = 3.39e38
      static_assert(max_dt_ppt == 1, "max_dt_euler_screen_kernel handles exactly one point per thread: its grid must cover `total` threads");
      auto k = max_dt_euler_screen_kernel<ND, RS>; HB_LAUNCH(k, grid, 256, 0, c->stream, a, c->ops, screen_bits);
    }
    else { auto k = max_dt_euler_kernel<ND, RS>; HB_LAUNCH(k, grid, 256, 0, c->stream, a, c->ops); }
    count_launch(c, ST_MAX_DT_CAR);
    HB_CUDA(c, cudaGetLastError());
    c->tss_is_one = !local_time;
    if (local_time) { *dt = 1.; return 0; }
    HB_CUDA(c, cudaMemcpyAsync(c->h_scalar, c->d_scalar, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    HB_CUDA(c, cudaStreamSynchronize(c->stream));
    *dt = *c->h_scalar;
    return 0;
  });
}

/* the same reduction with the result left on the device (hexed_b200_update_euler): no read-back, no synchronisation, graph-capturable */
int launch_max_dt_euler_device(hexed_b200_ctx* c, double safety_conv, double* d_dt)
{
  if (!c->have_mesh) return fail(c, HEXED_B200_NO_MESH, "no mesh uploaded");
  StatScope s_car(c, ST_MAX_DT_CAR, c->n_car);
  c->stats[ST_MAX_DT_DEF].work_units += c->n_def;
  return dispatch(c, [&](auto nd, auto rs) {
    constexpr int ND = decltype(nd)::value, RS = decltype(rs)::value;
    const long long total = (long long)c->n_elem*ipow(RS, ND);
    const int grid = (int)((total + 256*max_dt_ppt - 1)/(256*max_dt_ppt));
    MaxDtArgs a;
    a.state = c->state; a.tss = c->tss; a.vtss = c->vtss; a.n_elem = c->n_elem;
    a.max_cfl_c = (-2*c->quad_safety/c->min_eig_conv)*safety_conv;
    a.is_local = 0; a.global_min = reinterpret_cast<unsigned long long*>(d_dt);
    a.write_tss = !c->tss_is_one;
    HB_CUDA(c, cudaMemsetAsync(d_dt, 0x7f, sizeof(double), c->stream));
    if (grid && max_dt_running_screen && total < (1ll << 32) - 256) {
      int* screen_bits = reinterpret_cast<int*>(c->d_scalar + 1);
      HB_CUDA(c, cudaMemsetAsync(screen_bits, 0x7f, sizeof(int), c->stream));
      auto k = max_dt_euler_screen_kernel<ND, RS>; HB_LAUNCH(k, grid, 256, 0, c->stream, a, c->ops, screen_bits); count_launch(c, ST_MAX_DT_CAR);
    }
    else if (grid) { auto k = max_dt_euler_kernel<ND, RS>; HB_LAUNCH(k, grid, 256, 0, c->stream, a, c->ops); count_launch(c, ST_MAX_DT_CAR); }
    HB_CUDA(c, cudaGetLastError());
    c->tss_is_one = true;
    return 0;
  });
}

/* step = {nominal dt from the reduction, accumulated flow time, dt of this step = nominal * Chebyshev factor} */
__global__ void scale_dt_kernel(double* step, double factor) { step[2] = step[0]*factor; }
__global__ void accumulate_time_kernel(double* step) { step[1] += step[2]; }

int launch_scale_dt(hexed_b200_ctx* c, double* d_step, double factor)
{
  HB_LAUNCH(scale_dt_kernel, 1, 1, 0, c->stream, d_step, factor);
  ++c->launches;
  HB_CUDA(c, cudaGetLastError());
  return 0;
}

int launch_accumulate_time(hexed_b200_ctx* c, double* d_step)
{
  HB_LAUNCH(accumulate_time_kernel, 1, 1, 0, c->stream, d_step);
  ++c->launches;
  HB_CUDA(c, cudaGetLastError());
  return 0;
}

/* ---------------- hanging-node transfer: reference include/Spatial.hpp:153-206 (prolong), :230-285 (restrict) ----------------
 * One CTA per refined face; the fine faces are processed one after another in shared memory with the same
 * dimension-by-dimension in-place sweeps (and the same accumulation order into the coarse face) as the reference.
 * Restrict mutates the fine (mortar) data exactly like the reference does. */
template <int ND, int RS>
__device__ void transfer_sweeps(double* buf, int n_var, const double (*mat)[MAX_RS][MAX_RS], const int* str, int i_face, double mult, bool divide)
{
  constexpr int nfq = ipow(RS, ND - 1);
  for (int d = 0; d < ND - 1; ++d) {
    if (str[d]) {
      for (int i = threadIdx.x; i < n_var*nfq; i += blockDim.x) buf[i] = divide ? buf[i]/mult : buf[i]*mult;
    } else {
      const int pw = ND - 2 - d;
      const int face_stride = str[ND - 2 >= 0 ? ND - 2 : 0] ? 1 : ipow(2, pw);
      const int qstride = ipow(RS, pw);
      const int i_half = (i_face/face_stride) % 2;
      const int lines_per_var = nfq/RS;
      for (int line = threadIdx.x; line < n_var*lines_per_var; line += blockDim.x) {
        const int v = line/lines_per_var, l = line % lines_per_var;
        const int o = l/qstride, in = l % qstride;
        double* p = buf + v*nfq + o*RS*qstride + in;
        double row[RS], res[RS];
        #pragma unroll
        for (int k = 0; k < RS; ++k) row[k] = p[k*qstride];
        #pragma unroll
        for (int i = 0; i < RS; ++i) {
          double s = 0;
          #pragma unroll
          for (int k = 0; k < RS; ++k) s += mat[i_half][i][k]*row[k];
          res[i] = s;
        }
        #pragma unroll
        for (int k = 0; k < RS; ++k) p[k*qstride] = res[k];
      }
    }
    __syncthreads();
  }
}

template <int ND, int RS>
__global__ void __launch_bounds__(128)
prolong_kernel(double* faces, int width, int n_var, const int* ref_face, TransferOps ops, int scl, const int* ref_index)
{
  constexpr int nfq = ipow(RS, ND - 1);
  HB_DYN_SMEM(double, buf);
  const int* rf = ref_face + (size_t)(ref_index ? ref_index[blockIdx.x] : (int)blockIdx.x)*8;
  const int str[2] = {rf[5], rf[6]};
  int nf = ipow(2, ND - 1);
  for (int d = 0; d < ND - 1; ++d) nf /= 1 + str[d];
  const double* coarse = faces + (size_t)rf[0]*width;
  for (int f = 0; f < nf; ++f) {
    for (int i = threadIdx.x; i < n_var*nfq; i += blockDim.x) buf[i] = coarse[i];
    __syncthreads();
    transfer_sweeps<ND, RS>(buf, n_var, ops.prolong, str, f, 1. + scl, false);
    double* fine = faces + (size_t)rf[1 + f]*width;
    for (int i = threadIdx.x; i < n_var*nfq; i += blockDim.x) fine[i] = buf[i];
    __syncthreads();
  }
}

template <int ND, int RS>
__global__ void __launch_bounds__(128)
restrict_kernel(double* faces, int width, int n_var, const int* ref_face, TransferOps ops, int scl)
{
  constexpr int nfq = ipow(RS, ND - 1);
  HB_DYN_SMEM(double, buf);
  double* acc = buf + n_var*nfq;
  const int* rf = ref_face + (size_t)blockIdx.x*8;
  const int str[2] = {rf[5], rf[6]};
  int nf = ipow(2, ND - 1);
  for (int d = 0; d < ND - 1; ++d) nf /= 1 + str[d];
  for (int i = threadIdx.x; i < n_var*nfq; i += blockDim.x) acc[i] = 0.;
  for (int f = 0; f < nf; ++f) {
    double* fine = faces + (size_t)rf[1 + f]*width;
    for (int i = threadIdx.x; i < n_var*nfq; i += blockDim.x) buf[i] = fine[i];
    __syncthreads();
    transfer_sweeps<ND, RS>(buf, n_var, ops.restrict_, str, f, 1. + scl, true);
    for (int i = threadIdx.x; i < n_var*nfq; i += blockDim.x) { fine[i] = buf[i]; acc[i] += buf[i]; }
    __syncthreads();
  }
  double* coarse = faces + (size_t)rf[0]*width;
  for (int i = threadIdx.x; i < n_var*nfq; i += blockDim.x) coarse[i] = acc[i];
}

static double* face_array(hexed_b200_ctx* c, int kind) { return kind == 0 ? c->face_state : kind == 1 ? c->face_ldg : c->face_wide; }
static int face_width(hexed_b200_ctx* c, int kind) { return (kind == 2 ? c->nd + c->rs : c->nv)*c->nfq; }

template <bool PROLONG>
static int launch_transfer(hexed_b200_ctx* c, int kind, int n_var, int scale, const int* ref_index = nullptr, int n_index = 0)
{
  if (!c->have_mesh) return fail(c, HEXED_B200_NO_MESH, "no mesh uploaded");
  const int n_ref = ref_index ? n_index : c->n_ref;
  StatScope scope(c, ST_PR, n_ref);
  if (!n_ref) return 0;
  double* faces = face_array(c, kind);
  if (!faces) return fail(c, HEXED_B200_BAD_ARGUMENT, "face storage of the requested kind has not been allocated");
  if (!PROLONG && kind == 0) invalidate_admis(c); // restriction rewrites coarse ELEMENT faces (prolongation only mortar faces, which is_admissible always scans)
  const int width = face_width(c, kind);
  return dispatch(c, [&](auto nd, auto rs) {
    constexpr int ND = decltype(nd)::value, RS = decltype(rs)::value;
    const size_t smem = sizeof(double)*n_var*ipow(RS, ND - 1)*(PROLONG ? 1 : 2);
    if (PROLONG) { auto k = prolong_kernel<ND, RS>; HB_LAUNCH(k, n_ref, 128, smem, c->stream, faces, width, n_var, c->ref_face, c->transfer, scale, ref_index); }
    else { auto k = restrict_kernel<ND, RS>; HB_LAUNCH(k, n_ref, 128, smem, c->stream, faces, width, n_var, c->ref_face, c->transfer, scale); }
    count_launch(c, ST_PR);
    HB_CUDA(c, cudaGetLastError());
    return 0;
  });
}
int launch_prolong(hexed_b200_ctx* c, int kind, int n_var, int scale, const int* ref_index, int n_index) { return launch_transfer<true>(c, kind, n_var, scale, ref_index, n_index); }
int launch_restrict(hexed_b200_ctx* c, int kind, int n_var, int scale) { return launch_transfer<false>(c, kind, n_var, scale); }

/* ---------------- thermodynamic admissibility: Solver::is_admissible (reference src/Solver.cpp:921-958, src/thermo.cpp:6-18) ----------
 * Runs after EVERY stage of Solver::update when fix_admis is on, over the whole state and every face: left on the host it would pull
 * the full state across PCIe each stage. One warp per element scans the element's state and its 2*n_dim faces (mass > 0 and energy > 0
 * at every point; every variable finite, the reference's HEXED_ASSERT), writes Element::record (1 = inadmissible) and ORs two global
 * flags; the warps after the last element scan the fine mortar faces of the refined faces (:946-954), which set the flag but no record. */
__global__ void __launch_bounds__(256)
admissible_kernel(const double* state, const double* faces, const int* ref_face, int n_elem, int n_ref, int nd, int nq, int nfq, int* record, int* flags)
{
  const int warp = (int)(((long long)blockIdx.x*blockDim.x + threadIdx.x)/32), lane = threadIdx.x % 32;
  const bool in_range = warp < n_elem + n_ref; // no early return: the warp votes below want every lane
  const int nv = nd + 2;
  bool inadmissible = false, nonfinite = false;
  auto scan = [&](const double* data, int n_point) {
    for (int i = lane; i < nv*n_point; i += 32) {
      const double x = data[i];
      if (!isfinite(x)) nonfinite = true;
      if (i >= nd*n_point && !(x > 0.)) inadmissible = true;
    }
  };
  if (!in_range) {}
  else if (warp < n_elem) {
    scan(state + (size_t)warp*nv*nq, nq);
    for (int f = 0; f < 2*nd; ++f) scan(faces + ((size_t)warp*2*nd + f)*nv*nfq, nfq);
  }
  else {
    const int* rf = ref_face + (size_t)(warp - n_elem)*8;
    int n_fine = 1 << (nd - 1);
    for (int i = 0; i < nd - 1; ++i) n_fine /= 1 + rf[5 + i];
    for (int i = 0; i < n_fine; ++i) scan(faces + (size_t)rf[1 + i]*nv*nfq, nfq);
  }
  inadmissible = __any_sync(0xffffffffu, inadmissible);
  nonfinite = __any_sync(0xffffffffu, nonfinite);
  if (lane == 0 && in_range) {
    if (warp < n_elem) record[warp] = inadmissible ? 1 : 0;
    if (inadmissible) atomicOr(&flags[0], 1);
    if (nonfinite) atomicOr(&flags[1], 1);
  }
}

/* fused path: OR of the bits the pipelined Local kernels left; record[] is normalised to the reference's 0 / 1 on the way */
__global__ void __launch_bounds__(256)
record_reduce_kernel(int* record, int n_elem, int* flags)
{
  const int e = blockIdx.x*blockDim.x + threadIdx.x;
  const int bits = e < n_elem ? record[e] : 0;
  if (e < n_elem) record[e] = bits & 1;
  const int inadmissible = __any_sync(0xffffffffu, bits & 1), nonfinite = __any_sync(0xffffffffu, bits & 2);
  if (threadIdx.x % 32 == 0) {
    if (inadmissible) atomicOr(&flags[0], 1);
    if (nonfinite) atomicOr(&flags[1], 1);
  }
}

/* enqueue the check and the asynchronous read-back of its two flags; launch_is_admissible_finish waits for them. Split so that one host thread
 * driving several devices (hexed_b200_group_*) starts the check everywhere before it waits anywhere. */
int launch_is_admissible_begin(hexed_b200_ctx* c)
{
  if (!c->have_mesh) return fail(c, HEXED_B200_NO_MESH, "no mesh uploaded");
  if (!c->record) { HB_CUDA(c, cudaMalloc(&c->record, sizeof(int)*(c->n_elem ? c->n_elem : 1))); }
  if (!c->d_flags) { HB_CUDA(c, cudaMalloc(&c->d_flags, 2*sizeof(int))); HB_CUDA(c, cudaMallocHost(&c->h_flags, 2*sizeof(int))); }
  StatScope scope(c, ST_ADMIS, c->n_elem);
  HB_CUDA(c, cudaMemsetAsync(c->d_flags, 0, 2*sizeof(int), c->stream));
  const bool fused = c->use_fused_admis && (c->n_car == 0 || c->admis_valid[0]) && (c->n_def == 0 || c->admis_valid[1]);
  if (fused) {
    // the pipelined Local kernels left the bits of the state and element faces they wrote (nothing has touched either since): reduce
    // 4 bytes per element instead of scanning 17 KB, then scan only the fine mortar faces
    if (c->n_elem) { HB_LAUNCH(record_reduce_kernel, (c->n_elem + 255)/256, 256, 0, c->stream, c->record, c->n_elem, c->d_flags); count_launch(c, ST_ADMIS); }
    if (c->n_ref) {
      HB_LAUNCH(admissible_kernel, (int)(((long long)c->n_ref*32 + 255)/256), 256, 0, c->stream, c->state, c->face_state, c->ref_face, 0, c->n_ref,
                c->nd, c->nq, c->nfq, c->record, c->d_flags);
      count_launch(c, ST_ADMIS);
    }
    HB_CUDA(c, cudaGetLastError());
  }
  const long long warps = fused ? 0 : (long long)c->n_elem + c->n_ref;
  if (warps) {
    HB_LAUNCH(admissible_kernel, (int)((warps*32 + 255)/256), 256, 0, c->stream, c->state, c->face_state, c->ref_face, c->n_elem, c->n_ref,
              c->nd, c->nq, c->nfq, c->record, c->d_flags);
    count_launch(c, ST_ADMIS);
    HB_CUDA(c, cudaGetLastError());
  }
  HB_CUDA(c, cudaMemcpyAsync(c->h_flags, c->d_flags, 2*sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  return 0;
}

int launch_is_admissible_finish(hexed_b200_ctx* c, int* admissible)
{
  if (!c->h_flags) return fail(c, HEXED_B200_BAD_ARGUMENT, "hexed_b200_is_admissible_finish without _begin");
  HB_CUDA(c, cudaStreamSynchronize(c->stream));
  if (c->h_flags[1]) return fail(c, HEXED_B200_NOT_FINITE, "state is not finite");
  *admissible = c->h_flags[0] ? 0 : 1;
  return 0;
}

int launch_is_admissible(hexed_b200_ctx* c, int* admissible)
{
  int rc = launch_is_admissible_begin(c);
  return rc ? rc : launch_is_admissible_finish(c, admissible);
}

/* ---------------- ghost-state boundary conditions (reference src/Boundary_condition.cpp) ---------------- */
constexpr double specific_gas_air_c = 287.05287; // include/constants.hpp:44

/* Riemann_invariants for one face point (src/Boundary_condition.cpp:97-182): `in`/`gh`/`sc`/`nr` point at this point's first
 * variable (stride nfq). sign = velocity sign of incoming characteristics = 1 - 2*inside_face_sign. */
template <int ND>
__device__ void riemann_state_point(const double* in, double* gh, double* sc, const double* nr, const double* fs, int sign, int nfq)
{
  constexpr int NV = ND + 2;
  double inside[NV], outside[NV], n[ND], st[NV];
  #pragma unroll
  for (int v = 0; v < NV; ++v) { inside[v] = in[v*nfq]; outside[v] = fs[v]; }
  #pragma unroll
  for (int d = 0; d < ND; ++d) n[d] = nr[d*nfq];
  apply_char<ND>(inside, n, sign, inside, outside, st); // incoming characteristics from the freestream, outgoing left alone
  // limit the ghost state to keep it thermodynamically admissible (:125-128)
  st[ND] = fmax(st[ND], inside[ND]/2);
  double gsq = 0., isq = 0.;
  #pragma unroll
  for (int d = 0; d < ND; ++d) { gsq += st[d]*st[d]; isq += inside[d]*inside[d]; }
  const double kin_ener = .5*gsq/st[ND];
  const double inside_kin_ener = .5*isq/inside[ND];
  st[ND + 1] = fmax(kin_ener + fmax(st[ND + 1] - kin_ener, (inside[ND + 1] - inside_kin_ener)/2), 0.);
  #pragma unroll
  for (int v = 0; v < NV; ++v) { gh[v*nfq] = st[v]; sc[v*nfq] = inside[v]; } // state cache primed with the inside state (:134-138)
}

template <int ND>
__device__ void riemann_flux_point(const double* in, double* gh, const double* sc, const double* nr, int sign, int nfq)
{
  constexpr int NV = ND + 2;
  double flux[NV], zero[NV], cache[NV], n[ND], st[NV];
  #pragma unroll
  for (int v = 0; v < NV; ++v) { flux[v] = in[v*nfq]; cache[v] = sc[v*nfq]; zero[v] = 0.; }
  #pragma unroll
  for (int d = 0; d < ND; ++d) n[d] = nr[d*nfq];
  apply_char<ND>(cache, n, sign, zero, flux, st); // outgoing characteristic flux zero, incoming left alone (:165-173)
  #pragma unroll
  for (int v = 0; v < NV; ++v) gh[v*nfq] = st[v];
}
__global__ void __launch_bounds__(256)
bc_kernel(int kind, int n, const int* inside, const int* ghost, const int* normal, const double* params, double* cache,
          double* faces, double* faces_ldg, const double* normals, int nd, int nfq)
{
  const int nv = nd + 2, w = nv*nfq;
  const long long gid = (long long)blockIdx.x*blockDim.x + threadIdx.x;
  const int i = (int)(gid/nfq), q = (int)(gid % nfq);
  if (i >= n) return;
  double* gh = faces + (size_t)ghost[i]*w + q;
  if (kind == HEXED_B200_BC_FREESTREAM) { // Freestream::apply_state :66-76
    for (int v = 0; v < nv; ++v) gh[v*nfq] = params[v];
    return;
  }
  const double* in = faces + (size_t)inside[i]*w + q;
  if (kind == HEXED_B200_BC_COPY || kind == HEXED_B200_BC_OUTFLOW) { // copy_state :12-23 copies both halves of the face storage (Outflow::apply_state :465-468)
    for (int v = 0; v < nv; ++v) gh[v*nfq] = in[v*nfq];
    if (faces_ldg) for (int v = 0; v < nv; ++v) faces_ldg[(size_t)ghost[i]*w + q + v*nfq] = faces_ldg[(size_t)inside[i]*w + q + v*nfq];
    return;
  }
  if (kind == HEXED_B200_BC_NO_SLIP) { // No_slip::apply_state :367-385
    const double mass = in[nd*nfq], ener = in[(nd + 1)*nfq];
    for (int d = 0; d < nd; ++d) gh[d*nfq] = -in[d*nfq];
    gh[nd*nfq] = mass;
    // Thermal_bc::ghost_energy: Prescribed_energy -> energy_per_mass*mass, otherwise the inside energy (include/Boundary_condition.hpp:154-186)
    const double ge = ((int)params[0] == 1) ? params[1]*mass : ener;
    gh[(nd + 1)*nfq] = ge*ge/ener; // math::pow(x, 2)/state(last)
    double* sc = cache + (size_t)i*w + q; // state cache = average of ghost and inside state, used by the thermal flux condition
    for (int v = 0; v < nv; ++v) sc[v*nfq] = (gh[v*nfq] + in[v*nfq])/2;
    return;
  }
  const double* nr = normals + (size_t)normal[i]*nd*nfq + q;
  if (kind == HEXED_B200_BC_RIEMANN_INVARIANTS) { // Riemann_invariants::apply_state :97-139
    const int sign = 1 - 2*((inside[i] % (2*nd)) % 2);
    double* sc = cache + (size_t)i*w + q;
    if (nd == 3) riemann_state_point<3>(in, gh, sc, nr, params, sign, nfq);
    else if (nd == 2) riemann_state_point<2>(in, gh, sc, nr, params, sign, nfq);
    else riemann_state_point<1>(in, gh, sc, nr, params, sign, nfq);
    return;
  }
  if (kind == HEXED_B200_BC_PRESSURE_OUTFLOW) { // Pressure_outflow::apply_state :184-211
    const int sign = 2*((inside[i] % (2*nd)) % 2) - 1;
    const double mass = in[nd*nfq];
    double dot = 0., nsq = 0., msq = 0.;
    for (int d = 0; d < nd; ++d) { const double nn = nr[d*nfq], m = in[d*nfq]; dot += m*nn; nsq += nn*nn; msq += m*m; }
    const double nrml_veloc = dot/mass/sqrt(nsq);
    const double kin_ener = .5*msq/mass;
    const double pres = fmax(.4*(in[(nd + 1)*nfq] - kin_ener), 0.);
    const double sound_speed = sqrt(1.4*pres/mass);
    for (int v = 0; v < nv; ++v) gh[v*nfq] = in[v*nfq];
    if (nrml_veloc*sign < sound_speed) gh[(nd + 1)*nfq] = params[0]/.4 + kin_ener; // subsonic: impose the specified pressure
    return;
  }
  // Nonpenetration::apply_state -> reflect_momentum / reflect_normal :301-327
  double dot = 0., nsq = 0.;
  for (int d = 0; d < nd; ++d) { const double nn = nr[d*nfq]; dot += in[d*nfq]*nn; nsq += nn*nn; }
  for (int d = 0; d < nd; ++d) gh[d*nfq] = in[d*nfq] - 2*dot*nr[d*nfq]/nsq;
  gh[nd*nfq] = in[nd*nfq];
  gh[(nd + 1)*nfq] = in[(nd + 1)*nfq];
}

int launch_bcs(hexed_b200_ctx* c)
{
  if (!c->have_mesh) return fail(c, HEXED_B200_NO_MESH, "no mesh uploaded");
  long long total_faces = 0;
  for (auto& b : c->bcs) total_faces += b.n;
  StatScope scope(c, ST_BC, total_faces);
  for (auto& b : c->bcs) {
    if (!b.n) continue;
    const long long total = (long long)b.n*c->nfq;
    HB_LAUNCH(bc_kernel, (int)((total + 255)/256), 256, 0, c->stream, b.kind, b.n, b.inside, b.ghost, b.normal, b.params, b.cache,
              c->face_state, c->face_ldg, c->normals, c->nd, c->nfq);
    count_launch(c, ST_BC);
    HB_CUDA(c, cudaGetLastError());
  }
  return 0;
}

/* boundary loops of the artificial-viscosity and admissibility pipelines. mode:
 *   HEXED_B200_BC_MODE_ADVECTION  Flow_bc::apply_advection on the wide (n_dim + row_size variable) faces, Solver.cpp:505-510:
 *       default (src/Boundary_condition.cpp:24-41) velocity = inside, advected scalars = 2 - inside; Nonpenetration (:346-369) reflected
 *       velocity, scalars copied; No_slip (:429-448) negated velocity, scalars copied; Copy (:460-463) = copy_state, which copies the
 *       first 2*(n_dim + 2)*nfq doubles of the face storage
 *   HEXED_B200_BC_MODE_COPY_STATE  Flow_bc::apply_diffusion (:43-52) for every kind (Solver::apply_avc_diff_bcs, Solver.cpp:83-91) and
 *       the ghost copy inside fix_admissibility (Solver.cpp:1063-1068)
 *   HEXED_B200_BC_MODE_NEGATE_FLUX  Flow_bc::flux_diffusion (:54-60) for every kind (Solver::apply_avc_diff_flux_bcs, Solver.cpp:93-101)
 *       and Solver::apply_fta_flux_bcs (Solver.cpp:103-115) */
__global__ void __launch_bounds__(256)
aux_bc_kernel(int mode, int kind, int n, const int* inside, const int* ghost, const int* normal,
              double* faces, double* faces_ldg, double* faces_wide, const double* normals, int nd, int rs, int nfq)
{
  const int nv = nd + 2, w = nv*nfq, ww = (nd + rs)*nfq;
  const long long gid = (long long)blockIdx.x*blockDim.x + threadIdx.x;
  const int i = (int)(gid/nfq), q = (int)(gid % nfq);
  if (i >= n) return;
  if (mode == HEXED_B200_BC_MODE_COPY_STATE) {
    for (int v = 0; v < nv; ++v) faces[(size_t)ghost[i]*w + v*nfq + q] = faces[(size_t)inside[i]*w + v*nfq + q];
    return;
  }
  if (mode == HEXED_B200_BC_MODE_NEGATE_FLUX) {
    for (int v = 0; v < nv; ++v) faces_ldg[(size_t)ghost[i]*w + v*nfq + q] = -faces_ldg[(size_t)inside[i]*w + v*nfq + q];
    return;
  }
  const double* in = faces_wide + (size_t)inside[i]*ww + q;
  double* gh = faces_wide + (size_t)ghost[i]*ww + q;
  if (kind == HEXED_B200_BC_COPY) {
    const int n_copy = 2*nv < nd + rs ? 2*nv : nd + rs; // copy_state stops after 2*(n_dim + 2)*nfq doubles
    for (int v = 0; v < n_copy; ++v) gh[v*nfq] = in[v*nfq];
    return;
  }
  if (kind == HEXED_B200_BC_NONPENETRATION) {
    const double* nr = normals + (size_t)normal[i]*nd*nfq + q;
    double dot = 0., nsq = 0.;
    for (int d = 0; d < nd; ++d) { const double nn = nr[d*nfq]; dot += in[d*nfq]*nn; nsq += nn*nn; }
    for (int d = 0; d < nd; ++d) gh[d*nfq] = in[d*nfq] - 2*dot*nr[d*nfq]/nsq;
    for (int a = 0; a < rs; ++a) gh[(nd + a)*nfq] = in[(nd + a)*nfq];
    return;
  }
  if (kind == HEXED_B200_BC_NO_SLIP) {
    for (int d = 0; d < nd; ++d) gh[d*nfq] = -in[d*nfq];
    for (int a = 0; a < rs; ++a) gh[(nd + a)*nfq] = in[(nd + a)*nfq];
    return;
  }
  for (int d = 0; d < nd; ++d) gh[d*nfq] = in[d*nfq];
  for (int a = 0; a < rs; ++a) gh[(nd + a)*nfq] = 2. - in[(nd + a)*nfq];
}

int launch_aux_bcs(hexed_b200_ctx* c, int mode)
{
  if (!c->have_mesh) return fail(c, HEXED_B200_NO_MESH, "no mesh uploaded");
  if (mode < HEXED_B200_BC_MODE_ADVECTION || mode > HEXED_B200_BC_MODE_NEGATE_FLUX) return fail(c, HEXED_B200_BAD_ARGUMENT, "unknown boundary mode");
  if (mode == HEXED_B200_BC_MODE_ADVECTION && !c->face_wide) return fail(c, HEXED_B200_BAD_ARGUMENT, "advection faces have not been allocated");
  if (mode == HEXED_B200_BC_MODE_NEGATE_FLUX && !c->face_ldg) return fail(c, HEXED_B200_BAD_ARGUMENT, "LDG faces have not been allocated");
  long long total_faces = 0;
  for (auto& b : c->bcs) total_faces += b.n;
  StatScope scope(c, ST_BC, total_faces);
  for (auto& b : c->bcs) {
    if (!b.n) continue;
    const long long total = (long long)b.n*c->nfq;
    HB_LAUNCH(aux_bc_kernel, (int)((total + 255)/256), 256, 0, c->stream, mode, b.kind, b.n, b.inside, b.ghost, b.normal,
              c->face_state, c->face_ldg, c->face_wide, c->normals, c->nd, c->rs, c->nfq);
    count_launch(c, ST_BC);
    HB_CUDA(c, cudaGetLastError());
  }
  return 0;
}

/* flux boundary conditions: Solver::apply_flux_bcs (reference src/Solver.cpp:69-81) */
__global__ void __launch_bounds__(256)
flux_bc_kernel(int kind, int n, const int* inside, const int* ghost, const int* normal, const double* params, const double* cache,
               double* faces, double* faces_ldg, const double* normals, int nd, int nfq)
{
  const int nv = nd + 2, w = nv*nfq;
  const long long gid = (long long)blockIdx.x*blockDim.x + threadIdx.x;
  const int i = (int)(gid/nfq), q = (int)(gid % nfq);
  if (i >= n) return;
  double* gh = faces_ldg + (size_t)ghost[i]*w + q;
  const double* in = faces_ldg + (size_t)inside[i]*w + q;
  if (kind == HEXED_B200_BC_FREESTREAM || kind == HEXED_B200_BC_COPY) { // copy_state: both halves (src/Boundary_condition.cpp:12-23)
    for (int v = 0; v < nv; ++v) gh[v*nfq] = in[v*nfq];
    for (int v = 0; v < nv; ++v) faces[(size_t)ghost[i]*w + q + v*nfq] = faces[(size_t)inside[i]*w + q + v*nfq];
    return;
  }
  if (kind == HEXED_B200_BC_OUTFLOW || kind == HEXED_B200_BC_PRESSURE_OUTFLOW) { // negative of the inside flux :213-221,470-477
    for (int v = 0; v < nv; ++v) gh[v*nfq] = -in[v*nfq];
    return;
  }
  const double* nr = normals + (size_t)normal[i]*nd*nfq + q;
  if (kind == HEXED_B200_BC_RIEMANN_INVARIANTS) { // Riemann_invariants::apply_flux :141-182
    const int sign = 1 - 2*((inside[i] % (2*nd)) % 2);
    const double* sc = cache + (size_t)i*w + q;
    if (nd == 3) riemann_flux_point<3>(in, gh, sc, nr, sign, nfq);
    else if (nd == 2) riemann_flux_point<2>(in, gh, sc, nr, sign, nfq);
    else riemann_flux_point<1>(in, gh, sc, nr, sign, nfq);
    return;
  }
  if (kind == HEXED_B200_BC_NO_SLIP) { // No_slip::apply_flux :387-418
    for (int d = 0; d < nd; ++d) gh[d*nfq] = in[d*nfq];
    gh[nd*nfq] = -in[nd*nfq];
    double nsq = 0.;
    for (int d = 0; d < nd; ++d) { const double nn = nr[d*nfq]; nsq += nn*nn; }
    const double nrm = sqrt(nsq);
    const int flux_sign = 2*((inside[i] % (2*nd)) % 2) - 1;
    const double in_e = in[(nd + 1)*nfq];
    const int thermal = (int)params[0];
    double ghf; // Thermal_bc::ghost_heat_flux(state cache, inside heat flux per area)
    if (thermal == 0) ghf = params[1];
    else if (thermal == 1) ghf = in_e*flux_sign/nrm;
    else { // Thermal_equilibrium::ghost_heat_flux :359-365
      const double* sc = cache + (size_t)i*w + q;
      const double temp = sc[(nd + 1)*nfq]*.4/sc[nd*nfq]/specific_gas_air_c;
      const double radiative = params[1]*params[5]*(temp*temp*temp*temp);
      const double conductive = params[2]*(temp - params[3]);
      ghf = radiative + conductive;
    }
    gh[(nd + 1)*nfq] = params[4]*(nrm*flux_sign*ghf - in_e) + in_e;
    return;
  }
  // Nonpenetration::apply_flux (src/Boundary_condition.cpp:329-341): negate, then un-invert the normal momentum flux
  for (int v = 0; v < nv; ++v) gh[v*nfq] = -in[v*nfq];
  double dot = 0., nsq = 0.;
  for (int d = 0; d < nd; ++d) { const double nn = nr[d*nfq]; dot += gh[d*nfq]*nn; nsq += nn*nn; }
  for (int d = 0; d < nd; ++d) gh[d*nfq] -= 2*dot*nr[d*nfq]/nsq;
}

int launch_flux_bcs(hexed_b200_ctx* c)
{
  if (!c->have_mesh) return fail(c, HEXED_B200_NO_MESH, "no mesh uploaded");
  if (!c->face_ldg) return fail(c, HEXED_B200_BAD_ARGUMENT, "flux boundary conditions need the LDG face storage (run a viscous kernel first)");
  long long total_faces = 0;
  for (auto& b : c->bcs) total_faces += b.n;
  StatScope scope(c, ST_BC, total_faces);
  for (auto& b : c->bcs) {
    if (!b.n) continue;
    const long long total = (long long)b.n*c->nfq;
    HB_LAUNCH(flux_bc_kernel, (int)((total + 255)/256), 256, 0, c->stream, b.kind, b.n, b.inside, b.ghost, b.normal, b.params, b.cache,
              c->face_state, c->face_ldg, c->normals, c->nd, c->nfq);
    count_launch(c, ST_BC);
    HB_CUDA(c, cudaGetLastError());
  }
  return 0;
}

/* ---------------- smoothness indicator: reference src/stabilizing_art_visc.cpp:8-66 ----------------
 * One CTA per element. Pointwise and per-line terms are computed in parallel into shared memory; the two sums are then
 * accumulated by one thread in the reference's loop order, so the result does not depend on the launch geometry. */
struct StabArgs { const double* state; const double* nom; double* uncert; int n_elem; double char_speed; double proj[MAX_RS]; double weight[MAX_RS]; };

template <int ND, int RS>
__global__ void __launch_bounds__(ipow(RS, ND) < 32 ? 32 : ipow(RS, ND))
stab_art_visc_kernel(StabArgs a)
{
  constexpr int nq = ipow(RS, ND), nfq = nq/RS, nv = ND + 2;
  __shared__ double ind[nq], t1[nq], t2[ND*nfq];
  const int e = blockIdx.x, q = threadIdx.x;
  if (q < nq) {
    const double x = 1./a.state[((size_t)e*nv + ND)*nq + q];
    double w = 1;
    #pragma unroll
    for (int d = 0; d < ND; ++d) w *= a.weight[(q/ipow(RS, ND - 1 - d)) % RS];
    ind[q] = x;
    t1[q] = x*x*w;
  }
  __syncthreads();
  for (int item = q; item < ND*nfq; item += blockDim.x) {
    const int d = item/nfq, fq = item % nfq;
    const int stride = ipow(RS, ND - 1 - d);
    const int base = (fq/stride)*stride*RS + fq % stride;
    double dot = 0;
    #pragma unroll
    for (int k = 0; k < RS; ++k) dot += ind[base + k*stride]*a.proj[k];
    double fw = 1;
    #pragma unroll
    for (int dd = 0; dd < ND - 1; ++dd) fw *= a.weight[(fq/ipow(RS, ND - 2 - dd)) % RS];
    t2[item] = dot*dot*fw;
  }
  __syncthreads();
  if (q == 0) {
    double norm_sq = 0, nonsmooth = 0;
    for (int i = 0; i < nq; ++i) norm_sq += t1[i];
    for (int i = 0; i < ND*nfq; ++i) nonsmooth += t2[i];
    nonsmooth /= norm_sq*ND;
    const double ramp_center = -4.25*log((double)(RS - 1))/log(10.), half_width = 0.5;
    double indicator = log(nonsmooth)/log(10.);
    if (indicator <= ramp_center - half_width) indicator = 0;
    else if (indicator < ramp_center + half_width) indicator = .5*(1 + sin(3.14159265358979323846*(indicator - ramp_center)/2/half_width));
    else indicator = 1;
    a.uncert[e] = (RS - 1)*a.char_speed*a.nom[e]*indicator;
  }
}

int launch_stab_art_visc(hexed_b200_ctx* c, double char_speed)
{
  if (!c->have_mesh) return fail(c, HEXED_B200_NO_MESH, "no mesh uploaded");
  if (!c->n_elem) return 0;
  StabArgs a;
  a.state = c->state; a.nom = c->nom; a.uncert = c->uncert; a.n_elem = c->n_elem; a.char_speed = char_speed;
  for (int i = 0; i < MAX_RS; ++i) { a.weight[i] = i < c->rs ? c->weight[i] : 0.; a.proj[i] = i < c->rs ? c->orthogonal[c->rs - 1][i]*c->weight[i] : 0.; }
  return dispatch(c, [&](auto nd, auto rs) {
    constexpr int ND = decltype(nd)::value, RS = decltype(rs)::value;
    constexpr int threads = ipow(RS, ND) < 32 ? 32 : ipow(RS, ND);
    auto k = stab_art_visc_kernel<ND, RS>;
    HB_LAUNCH(k, c->n_elem, threads, 0, c->stream, a);
    ++c->launches;
    HB_CUDA(c, cudaGetLastError());
    return 0;
  });
}

/* ---------------- packed gather / scatter of face slots (boundary faces crossing PCIe) ---------------- */
__global__ void __launch_bounds__(256) gather_kernel(const double* src, int width, const int* slots, int n, double* dst)
{
  const long long gid = (long long)blockIdx.x*blockDim.x + threadIdx.x;
  const long long i = gid/width; const int k = (int)(gid % width);
  if (i < n) dst[i*width + k] = src[(size_t)slots[i]*width + k];
}
__global__ void __launch_bounds__(256) scatter_kernel(double* dst, int width, const int* slots, int n, const double* src)
{
  const long long gid = (long long)blockIdx.x*blockDim.x + threadIdx.x;
  const long long i = gid/width; const int k = (int)(gid % width);
  if (i < n) dst[(size_t)slots[i]*width + k] = src[i*width + k];
}
int launch_gather_faces(hexed_b200_ctx* c, const double* src, int width, const int* d_slots, int n, double* dst)
{
  const long long total = (long long)n*width;
  if (!total) return 0;
  HB_LAUNCH(gather_kernel, (int)((total + 255)/256), 256, 0, c->stream, src, width, d_slots, n, dst);
  ++c->launches;
  HB_CUDA(c, cudaGetLastError());
  return 0;
}
int launch_scatter_faces(hexed_b200_ctx* c, double* dst, int width, const int* d_slots, int n, const double* src)
{
  if (dst == c->face_state) invalidate_admis(c);
  const long long total = (long long)n*width;
  if (!total) return 0;
  HB_LAUNCH(scatter_kernel, (int)((total + 255)/256), 256, 0, c->stream, dst, width, d_slots, n, src);
  ++c->launches;
  HB_CUDA(c, cudaGetLastError());
  return 0;
}

/* ---------------- Face_permutation on one face (reference include/Spatial.hpp:73-131) ---------------- */
__global__ void permute_face_kernel(double* data, double* tmp, int n_var, int nfq, const int* table, int restore)
{
  const int total = n_var*nfq;
  for (int i = threadIdx.x; i < total; i += blockDim.x) tmp[i] = data[i];
  __syncthreads();
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int v = i/nfq, p = i % nfq;
    if (restore) data[v*nfq + table[p]] = tmp[i];   // inverse permutation
    else data[i] = tmp[v*nfq + table[p]];           // matched[p] = original[table[p]]
  }
}
int launch_permute_face(hexed_b200_ctx* c, double* d_data, int n_var, int code, int restore)
{
  HB_LAUNCH(permute_face_kernel, 1, 128, 0, c->stream, d_data, d_data + n_var*c->nfq, n_var, c->nfq, c->perm + code*c->nfq, restore);
  ++c->launches;
  HB_CUDA(c, cudaGetLastError());
  return 0;
}

} // namespace hb
