/* common.cuh -- context, device-side constants and dispatch helpers shared by all translation units. */
#ifndef HB_COMMON_CUH_
#define HB_COMMON_CUH_

#include "hb_rt.cuh"
#include "../../include/hexed_b200.h"
#include <cfloat>
#include <cstdio>
#include <string>
#include <vector>

namespace hb {

constexpr int MAX_RS = 8;

__host__ __device__ constexpr int ipow(int b, int e) { int r = 1; for (int i = 0; i < e; ++i) r *= b; return r; }

/* 1-D operators of the basis, passed by value to kernels (kernel parameters live in the constant bank).
 * `dfull = diff_mat - lift*boundary` folds the interior part of the DG derivative
 *   D(q, b) = diff_mat q + lift (b - boundary q)        (reference include/Derivative.hpp:51-55)
 * into one matrix so that D(q, b) = dfull q + lift b. */
struct Ops
{
  double diff[MAX_RS][MAX_RS];
  double dfull[MAX_RS][MAX_RS];
  double bnd[2][MAX_RS];
  double lift[MAX_RS][2];
  double node[MAX_RS];
  /* Even-odd halves of the same operators (H = row_size/2, the factor 1/2 folded in), valid when the nodes are symmetric about 1/2 (Gauss-Legendre and
   * Gauss-Lobatto are): then dfull[RS-1-i][RS-1-k] = -dfull[i][k], lift[RS-1-i][0] = -lift[i][1], bnd[1][k] = bnd[0][RS-1-k], and a line
   * derivative costs RS*(RS/2 + 1) multiply-adds on the sums and differences of mirrored points instead of RS*(RS + 2) (line_deriv_eo).
   *   eo_a = (dfull[i][k] - dfull[i][RS-1-k])/2, eo_s = (... + ...)/2, eo_la = (lift[i][0] - lift[i][1])/2, eo_ls = (... + ...)/2,
   *   eo_da / eo_ds likewise for diff, eo_ba = (bnd[0][k] - bnd[0][RS-1-k])/2, eo_bs = (... + ...)/2 */
  double eo_a[MAX_RS/2][MAX_RS/2], eo_s[MAX_RS/2][MAX_RS/2], eo_la[MAX_RS/2], eo_ls[MAX_RS/2];
  double eo_da[MAX_RS/2][MAX_RS/2], eo_ds[MAX_RS/2][MAX_RS/2];
  double eo_ba[MAX_RS/2], eo_bs[MAX_RS/2];
};

#if defined(__CUDACC__) || defined(HB_EMULATE)
/* r = dfull f + lift (b0, b1) -- the DG line derivative D(q, b) of include/Derivative.hpp:51-55 -- through the even-odd halves:
 * with a[k] = f[k] - f[RS-1-k], s[k] = f[k] + f[RS-1-k]:  r[i] = P + Q, r[RS-1-i] = P - Q,
 * P = sum_k eo_a[i][k] a[k] + eo_la[i] (b0 - b1), Q = sum_k eo_s[i][k] s[k] + eo_ls[i] (b0 + b1).  NEG: returns -r (the residual sign). */
template <int RS, bool NEG = false>
__device__ __forceinline__ void line_deriv_eo(const Ops& ops, const double (&f)[RS], double b0, double b1, double (&r)[RS])
{
  constexpr int H = RS/2;
  static_assert(RS % 2 == 0, "even row sizes only");
  double a[H], s[H];
  #pragma unroll
  for (int k = 0; k < H; ++k) { a[k] = f[k] - f[RS - 1 - k]; s[k] = f[k] + f[RS - 1 - k]; }
  const double bd = b0 - b1, bs = b0 + b1;
  #pragma unroll
  for (int i = 0; i < H; ++i) {
    double p = 0, q = 0;
    #pragma unroll
    for (int k = 0; k < H; ++k) { p += ops.eo_a[i][k]*a[k]; q += ops.eo_s[i][k]*s[k]; }
    p += ops.eo_la[i]*bd;
    q += ops.eo_ls[i]*bs;
    if constexpr (NEG) { r[i] = -p - q; r[RS - 1 - i] = q - p; }
    else { r[i] = p + q; r[RS - 1 - i] = p - q; }
  }
}
/* r = diff_mat f (no boundary term) */
template <int RS, bool NEG = false>
__device__ __forceinline__ void line_diff_eo(const Ops& ops, const double (&f)[RS], double (&r)[RS])
{
  constexpr int H = RS/2;
  double a[H], s[H];
  #pragma unroll
  for (int k = 0; k < H; ++k) { a[k] = f[k] - f[RS - 1 - k]; s[k] = f[k] + f[RS - 1 - k]; }
  #pragma unroll
  for (int i = 0; i < H; ++i) {
    double p = 0, q = 0;
    #pragma unroll
    for (int k = 0; k < H; ++k) { p += ops.eo_da[i][k]*a[k]; q += ops.eo_ds[i][k]*s[k]; }
    if constexpr (NEG) { r[i] = -p - q; r[RS - 1 - i] = q - p; }
    else { r[i] = p + q; r[RS - 1 - i] = p - q; }
  }
}
/* e0 = bnd[0] . x, e1 = bnd[1] . x (extrapolation of a line to its two faces, include/Spatial.hpp:41-57) */
template <int RS>
__device__ __forceinline__ void face_extrap_eo(const Ops& ops, const double (&x)[RS], double& e0, double& e1)
{
  constexpr int H = RS/2;
  double p = 0, q = 0;
  #pragma unroll
  for (int k = 0; k < H; ++k) { p += ops.eo_bs[k]*(x[k] + x[RS - 1 - k]); q += ops.eo_ba[k]*(x[k] - x[RS - 1 - k]); }
  e0 = p + q; e1 = p - q;
}
#endif

struct FilterOp { double filter[MAX_RS][MAX_RS]; };
struct TransferOps { double prolong[2][MAX_RS][MAX_RS]; double restrict_[2][MAX_RS][MAX_RS]; };

struct Stat
{
  const char* name; int deformed; long long work_units = 0; long long launches = 0; double seconds = 0;
};
enum { ST_NEIGHBOR_CAR, ST_NEIGHBOR_DEF, ST_LOCAL_CAR, ST_LOCAL_DEF, ST_MAX_DT_CAR, ST_MAX_DT_DEF, ST_PR, ST_BC, ST_WRITE_FACE,
       ST_RECONCILE_CAR, ST_RECONCILE_DEF, ST_ADMIS, ST_COUNT };

/* scalar parameters of every PDE, passed by value to the generic kernels (pde.cuh) */
struct PdeParams
{
  hexed_b200_transport visc, cond; // Navier-Stokes
  double visc_inv_sqrt_ref, cond_inv_sqrt_ref; // 1/sqrt_ref_temp of the two models (one division per point and model less on the device)
  double p0, p1;                   // advection: advect_length | smooth_av: diff_time, chebyshev_step
  double adv_nodes[MAX_RS];        // advection: Gauss-Legendre nodes mapped to [-1, 1] (reference include/pde.hpp:281)
};

struct FaceList
{
  int n = 0; int* d_slots = nullptr; double* d_buf = nullptr; size_t buf_doubles = 0;
  // asynchronous host traffic of host-applied boundary conditions (hexed_b200_face_list_prefetch / _upload_deferred): pinned host
  // buffers, a second device buffer for the upload so that it cannot collide with a download in flight, events on the copy stream
  double* h_down = nullptr; double* h_up = nullptr; double* d_up = nullptr;
  cudaEvent_t ev_down = nullptr, ev_up = nullptr;
  int down_kind = -1; // face kind of the prefetch in flight (-1: none)
  int up_kind = -1;   // face kind of the deferred upload waiting to be scattered (-1: none)
};
struct Bc { int kind; int n; int *inside = nullptr, *ghost = nullptr, *normal = nullptr; double* params = nullptr; int n_params = 0;
            double* cache = nullptr; /* No_slip: Boundary_face::state_cache(), [n][nv*nfq] */ };

} // namespace hb

struct hexed_b200_ctx
{
  int device = 0;
  int nd = 0, rs = 0, nq = 0, nfq = 0, nv = 0, n_vert = 0;
  hb::Ops ops;
  hb::FilterOp filt;
  hb::TransferOps transfer;
  double weight[hb::MAX_RS];
  double gl_node[hb::MAX_RS]; // nodes of Gauss_legendre(row_size), used by pde::Advection whatever the solution basis
  double orthogonal[hb::MAX_RS][hb::MAX_RS];
  double min_eig_conv = 0, min_eig_diff = 0, quad_safety = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr; // PCIe traffic of host-applied boundary conditions, overlapped with kernels on `stream`
  cudaEvent_t ev_copy = nullptr;
  int n_pending_uploads = 0;          // face lists whose deferred upload has not been scattered yet
  // mesh
  bool have_mesh = false;
  int n_car = 0, n_def = 0, n_elem = 0, n_face_slot = 0, n_normal_slot = 0, n_car_con = 0, n_def_con = 0, n_ref = 0;
  double *state = nullptr, *tss = nullptr, *cache = nullptr, *av = nullptr, *forcing = nullptr, *adv = nullptr;
  double *nom = nullptr, *vtss = nullptr, *uncert = nullptr, *refn = nullptr, *det = nullptr;
  double *face_state = nullptr, *face_ldg = nullptr, *face_wide = nullptr, *normals = nullptr;
  int *car_con = nullptr;  // [n][4]: slot0, slot1, i_dim, 0
  int *def_con = nullptr;  // [n][4]: slot0, slot1, dir code, normal slot
  int *ref_face = nullptr; // [n][8]: coarse, fine0..3, stretch0, stretch1, 0
  int *perm = nullptr;     // [36][nfq] face permutation tables, indexed by dir code
  // domain decomposition (hexed_b200_set_partition): cut connections sit at the end of the tables
  int n_cut_car = 0, n_cut_def = 0, n_pre_prolong = 0;
  int *pre_prolong = nullptr; // indices into ref_face whose coarse face is a halo slot
  std::vector<int> h_perm;
  // scratch
  double* d_scalar = nullptr; double* h_scalar = nullptr;
  double* d_face_scratch = nullptr;
  std::vector<hb::FaceList> lists;
  std::vector<hb::Bc> bcs;
  // stats
  hb::Stat stats[hb::ST_COUNT];
  bool timing = false;
  bool ops_symmetric = false; // the basis nodes are symmetric about 1/2: the even-odd operator halves (Ops::eo_*) are valid, which the line-task kernels require
  bool use_pipe = true; // TMA-pipelined Local kernel where it applies (hexed_b200_set_option)
  bool pipe_lean4 = false; // 3-D Cartesian Euler: lean layout with four resident CTAs (option value 3; experiment)
  int ns_layout = 2; // 3-D row-size-6 Navier-Stokes Local kernel variant, option HEXED_B200_OPT_NS_LOCAL_LAYOUT (include/hexed_b200.h)
  bool pipe_end_barrier = true; // 3-D Euler pipelined kernels: CTA barrier at the end of an iteration; false (option value 4) = mbarrier hand-over of the stage buffer, measured equal
  bool pipe_lean = true; // 3-D deformed Euler: the 67 KB / three-CTA layout of the pipelined kernel (option value 2 = the classic 105 KB one)
  // CFL screen: single-precision min over the element's points of spacing/char_speed of the state the stage-1 Local kernel has
  // just written (0 = not representable, always re-evaluate); cfl_valid[0|1] = every Cartesian | deformed element's entry belongs
  // to the current state. Anything else that writes the state or the vertex spacing clears the flags (invalidate_cfl_cache), and
  // max_dt_euler then runs its full kernel. With valid entries it re-evaluates in FP64 only the elements within 1e-5 of the minimum.
  bool use_cfl_cache = false; // measured break-even on B200 (profiles/r01g_ncu_full_euler.md): off unless asked for
  float* cfl_approx = nullptr;
  bool cfl_valid[2] = {false, false};
  // fused admissibility (HEXED_B200_OPT_FUSED_ADMIS): the pipelined Local kernels leave Element::record-style bits of the state and
  // faces they have just written; admis_valid[0|1] = the bits of every Cartesian | deformed element describe the CURRENT state and
  // element faces. Every other writer of element state or element faces clears the flags (invalidate_admis), and is_admissible then
  // runs its full scan.
  bool use_fused_admis = false;
  bool admis_valid[2] = {false, false};
  // hexed_b200_update_euler: the time step stays on the device. d_step = {dt of the current step, accumulated flow time}; while
  // dt_dev_active is non-null the Euler Local launchers hand it to their kernels, which multiply their `update` factor by it.
  double* d_step = nullptr; const double* dt_dev_active = nullptr;
  double* max_dt_device_out = nullptr; // non-null: the generic max_dt leaves its result there instead of reading it back
  bool tss_is_one = false; // time_step_scale is known to hold 1. everywhere (written by a global-time-step max_dt)
  // vertex topology of the epoch (hexed_b200_vertex_topology) and per-element-vertex scratch (vertex_fix_admis_coef / vertex_elwise_av)
  int* elem_vertex = nullptr; int n_vertex = 0; int* matchers = nullptr; int n_match = 0;
  double* vertex_vals = nullptr; double* vertex_scratch = nullptr;
  int* record = nullptr;  // Element::record as left by is_admissible: 1 = thermodynamically inadmissible element
  int* d_flags = nullptr; int* h_flags = nullptr; // {inadmissible, non-finite} found by the last is_admissible
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  long long launches = 0;
  std::string err;
};

namespace hb {

/* direction code used to index the permutation tables: i_dim0 + 3*i_dim1 + 9*sign0 + 18*sign1 */
__host__ __device__ inline int dir_code(int d0, int d1, int s0, int s1) { return d0 + 3*d1 + 9*s0 + 18*s1; }

int fail(hexed_b200_ctx* c, int code, const std::string& msg);
int flush_pending_uploads(hexed_b200_ctx* c);
int check(hexed_b200_ctx* c, cudaError_t e, const char* what);
#define HB_CUDA(c, call) do { int hb_rc_ = hb::check(c, (call), #call); if (hb_rc_) return hb_rc_; } while (0)
/* every C ABI entry point makes its context's device current: one host thread may drive several contexts on different devices
 * (hexed_b200_group_*, the single-process multi-GPU path behind the C++ adapter) */
#define HB_ENTER_KEEP(c) do { cudaSetDevice((c)->device); } while (0)
/* ... and, unless the entry point places them itself (the stage drivers scatter them after their interior Neighbor kernels), completes
 * deferred ghost-face uploads first, so that nothing ever reads a face that is still on its way */
#define HB_ENTER(c) do { cudaSetDevice((c)->device); if ((c)->n_pending_uploads) { int hb_rc_ = hb::flush_pending_uploads(c); if (hb_rc_) return hb_rc_; } } while (0)

struct StatScope
{
  hexed_b200_ctx* c; int id;
  StatScope(hexed_b200_ctx* ctx, int stat_id, long long work) : c{ctx}, id{stat_id}
  {
    c->stats[id].work_units += work;
    if (c->timing) cudaEventRecord(c->ev0, c->stream);
  }
  ~StatScope()
  {
    if (c->timing) {
      cudaEventRecord(c->ev1, c->stream);
      cudaEventSynchronize(c->ev1);
      float ms = 0; cudaEventElapsedTime(&ms, c->ev0, c->ev1);
      c->stats[id].seconds += ms*1e-3;
    }
  }
};
/* scratch device allocation released on every exit path */
template <class T> struct DevScratch
{
  T* p = nullptr;
  ~DevScratch() { if (p) cudaFree(p); }
  cudaError_t alloc(size_t n) { return cudaMalloc(&p, sizeof(T)*(n ? n : 1)); }
};

inline void invalidate_admis(hexed_b200_ctx* c) { c->admis_valid[0] = c->admis_valid[1] = false; }
/* called by everything that rewrites the flow state (or the vertex spacing) outside the pipelined Local kernels */
inline void invalidate_cfl_cache(hexed_b200_ctx* c) { c->cfl_valid[0] = c->cfl_valid[1] = false; invalidate_admis(c); }
inline void count_launch(hexed_b200_ctx* c, int stat_id) { ++c->launches; ++c->stats[stat_id].launches; }

/* launchers implemented in the kernel translation units; return a HEXED_B200_* code */
int launch_neighbor_euler(hexed_b200_ctx* c, int deformed, int first = 0, int count = -1);
int launch_local_euler(hexed_b200_ctx* c, int deformed, hexed_b200_options o);
int launch_local_euler_pipe(hexed_b200_ctx* c, int deformed, hexed_b200_options o, int begin, int end);
int launch_local_euler_pipe2d(hexed_b200_ctx* c, int deformed, hexed_b200_options o, int begin, int end);
int launch_write_face(hexed_b200_ctx* c);
int launch_max_dt_euler(hexed_b200_ctx* c, double safety_conv, int local_time, double* dt);
int launch_max_dt_euler_device(hexed_b200_ctx* c, double safety_conv, double* d_dt); // global time step, result left at d_dt (no read-back)
int launch_accumulate_time(hexed_b200_ctx* c, double* d_step);
int launch_scale_dt(hexed_b200_ctx* c, double* d_step, double factor);
int launch_prolong(hexed_b200_ctx* c, int kind, int n_var, int scale, const int* ref_index = nullptr, int n_index = 0);
int launch_restrict(hexed_b200_ctx* c, int kind, int n_var, int scale);
int launch_bcs(hexed_b200_ctx* c);
int launch_gather_faces(hexed_b200_ctx* c, const double* src, int width, const int* d_slots, int n, double* dst);
int launch_scatter_faces(hexed_b200_ctx* c, double* dst, int width, const int* d_slots, int n, const double* src);
int launch_permute_face(hexed_b200_ctx* c, double* d_data, int n_var, int code, int restore);

/* per-PDE launchers of the generic kernels (generic_part.cu, one translation unit per PDE) */
struct GenericOps
{
  int (*neighbor)(hexed_b200_ctx*, int deformed, const PdeParams&, bool reconcile, int first, int count); // count < 0: to the end
  int (*local)(hexed_b200_ctx*, int deformed, hexed_b200_options, const PdeParams&, bool reconcile);
  int (*write_face)(hexed_b200_ctx*, const PdeParams&);
  int (*max_dt)(hexed_b200_ctx*, const PdeParams&, double safety_conv, double safety_diff, int local_time, double* dt);
};
extern const GenericOps generic_ops_pde1, generic_ops_pde2, generic_ops_pde3, generic_ops_pde4;
int launch_stab_art_visc(hexed_b200_ctx* c, double char_speed);
int launch_flux_bcs(hexed_b200_ctx* c);
int launch_is_admissible(hexed_b200_ctx* c, int* admissible);
int launch_is_admissible_begin(hexed_b200_ctx* c);
int launch_is_admissible_finish(hexed_b200_ctx* c, int* admissible);
int launch_set_jacobian(hexed_b200_ctx* c, const double* d_vert, const double* d_node_adj);
int launch_shared_normals(hexed_b200_ctx* c);
int launch_av_scale_velocity(hexed_b200_ctx* c, int restore);
int launch_av_project_forcing(hexed_b200_ctx* c, const double* weights, const double* orth);
int launch_av_finish(hexed_b200_ctx* c, double mult, double us_max, int n_real, const double* node_weights, double* resid_sq);
int launch_interp_vertices(hexed_b200_ctx* c, int target, const double* d_vert, const double* d_interp);
int launch_av_swap(hexed_b200_ctx* c);
int launch_aux_bcs(hexed_b200_ctx* c, int mode);
int launch_share_vertex_data(hexed_b200_ctx* c, double* elem_vals, int is_max);
int launch_fix_admis_spread(hexed_b200_ctx* c, const double* d_interp);
int launch_av_elwise_ramp(hexed_b200_ctx* c, double scale);
int launch_av_elwise_forcing(hexed_b200_ctx* c, int dir);
int launch_av_elwise_vertices(hexed_b200_ctx* c, const double* d_interp);

/* Thread -> line-task map. In the dense [i][j][k] field layout that the bulk copies deliver, lines of dimension 0 (stride RS^2) are
 * conflict-free for consecutive lanes, but with 8-byte accesses consecutive lines of dimension 1 (stride RS) and 2 (stride 1) hit every
 * shared-memory bank twice (profiles/r01g_ncu_full_euler.md: 217 M of 537 M shared wavefronts were bank conflicts, L1/shared pipe 73 %
 * busy). At row size 6 the map therefore
 *   - gives dimension-2 lines 16-byte accesses (two points per LDS.128/STS.128): 8 consecutive lines then cover 8 distinct 16-byte
 *     bank groups ((2i + 3j) mod 8 advances by 3 per line), conflict-free;
 *   - places dimension-1 lines half-warp by half-warp as rows i = (0,2), (1,3), (4,5) with 6 active lanes of every 8: rows 0/2 and 1/3
 *     sit 8 banks apart (conflict-free), only the (4,5) pair still shares two banks.
 * 120 of 128 threads carry a line. Other row sizes keep the plain map.
 * Measured (profiles/r01i_ncu_full_euler.md against r01g): shared wavefronts -19 %, conflicts 217 M -> 67 M, L1/shared pipe 73 % -> 50 %
 * busy, and yet the deformed Euler kernel got 3 % SLOWER (104 -> 168 registers, +11 % instructions, fixed-latency "wait" stalls 1.4 ->
 * 2.4 per issue): with 8 warps per SM the kernel is latency-bound, not LSU-bound. The Cartesian kernel, which has no normals to stage
 * and half the arithmetic per line, gained 7 %. The map is therefore used by the Cartesian instantiation only (SPARSE = !DEF), and the
 * Navier-Stokes line kernel (8.2 -> 8.5 ms with it) keeps the plain map. */
template <int RS, bool SPARSE = true> struct LineMap
{
  static constexpr bool vec2 = false; // 16-byte accesses along dimension 2
  __device__ static __forceinline__ bool get(int t, int& d, int& l)
  {
    d = t/(RS*RS); l = t % (RS*RS);
    return t < 3*RS*RS;
  }
};
template <> struct LineMap<6, true>
{
  static constexpr bool vec2 = true;
  __device__ static __forceinline__ bool get(int t, int& d, int& l)
  {
    if (t < 36) { d = 0; l = t; return true; }
    if (t < 72) { d = 2; l = t - 36; return true; }
    if (t < 80) { d = 0; l = 0; return false; }
    d = 1;
    const int u = t - 80, h = u/16, w = u % 16, slot = w/8, k = w % 8;
    const int i = h < 2 ? h + 2*slot : 4 + slot;
    l = i*6 + (k < 6 ? k : 0);
    return k < 6;
  }
};

__device__ __forceinline__ void ld2(const double* p, double& a, double& b)
{ const double2 v = *reinterpret_cast<const double2*>(p); a = v.x; b = v.y; }
__device__ __forceinline__ void st2(double* p, double a, double b)
{ double2 v; v.x = a; v.y = b; *reinterpret_cast<double2*>(p) = v; }

/* run-time (n_dim, row_size) -> compile-time dispatch; the analogue of the reference's kernel_factory
 * (include/kernel_factory.hpp:64-119). F is a generic lambda taking two integral_constants. */
template <int V> struct IC { static constexpr int value = V; };
template <class F>
int dispatch(hexed_b200_ctx* c, F&& f)
{
  #define HB_CASE(ND, RS) if (c->nd == ND && c->rs == RS) return f(IC<ND>{}, IC<RS>{});
  #define HB_ROW(ND) HB_CASE(ND, 2) HB_CASE(ND, 3) HB_CASE(ND, 4) HB_CASE(ND, 5) HB_CASE(ND, 6) HB_CASE(ND, 7) HB_CASE(ND, 8)
  HB_ROW(1) HB_ROW(2) HB_ROW(3)
  #undef HB_ROW
  #undef HB_CASE
  return fail(c, HEXED_B200_INVALID_KERNEL, "demand for invalid kernel");
}

} // namespace hb
#endif
