/* av_glue.cu -- the pointwise loops Solver wraps around the artificial-viscosity kernels.
 *
 * Solver::update_art_visc_smoothness (reference src/Solver.cpp:457-581) alternates the advection / smoothing kernels of this library
 * with four loops over every quadrature point of the host objects; Solver::fix_admissibility (:1021-1031), set_art_visc_admis
 * (:636-658) and update_art_visc_elwise (:625-633) interpolate per-vertex values to the quadrature points. With the state resident on
 * the device each of them would be a full download + upload, so they are provided here:
 *   av_scale_velocity      :467-478 / :567-571   momentum /= or *= sqrt(2*mass*energy)
 *   av_project_forcing     :527-541   forcing[0] = (sum_i adv_i w_i orth_i)^2 * 2*energy/mass
 *   av_finish              :551-572   bulk AV coefficient from forcing[n_real], squared residual, velocity restored
 *   interp_vertices        math::hypercube_matvec(interp, vertex values) -> laplacian AV coefficient (or any element slot group)
 *   av_swap                :1032-1038,1080-1086   swap bulk and laplacian AV coefficients
 */
#include "common.cuh"

namespace hb {

__global__ void __launch_bounds__(256)
av_scale_velocity_kernel(double* state, long long n_point_total, int nd, int nq, int restore)
{
  const long long gid = (long long)blockIdx.x*blockDim.x + threadIdx.x;
  if (gid >= n_point_total) return;
  const long long e = gid/nq; const int q = (int)(gid % nq);
  double* s = state + (size_t)e*(nd + 2)*nq + q;
  const double scale = sqrt(2*s[nd*nq]*s[(nd + 1)*nq]);
  for (int d = 0; d < nd; ++d) { if (restore) s[d*nq] *= scale; else s[d*nq] /= scale; }
}

struct ProjArgs { double w[MAX_RS], orth[MAX_RS]; };

__global__ void __launch_bounds__(256)
av_project_forcing_kernel(const double* state, const double* adv, double* forcing, long long n_point_total, int nd, int nq, int rs, ProjArgs p)
{
  const long long gid = (long long)blockIdx.x*blockDim.x + threadIdx.x;
  if (gid >= n_point_total) return;
  const long long e = gid/nq; const int q = (int)(gid % nq);
  double proj = 0;
  for (int i = 0; i < rs; ++i) proj += adv[((size_t)e*rs + i)*nq + q]*p.w[i]*p.orth[i];
  const double* s = state + (size_t)e*(nd + 2)*nq + q;
  forcing[(size_t)e*4*nq + q] = proj*proj*2*s[(nd + 1)*nq]/s[nd*nq];
}

/* one CTA per element: new AV coefficient, velocity restored, and the element's share of the squared residual reduced in a fixed
 * order (shuffle tree, then warp 0), so the residual is reproducible run to run */
template <int ND, int RS>
__global__ void __launch_bounds__(256)
av_finish_kernel(double* state, double* av, const double* forcing, const double* nom, int n_elem, double mult, double us_max, int n_real,
                 ProjArgs p, double* partial)
{
  constexpr int nq = ipow(RS, ND);
  __shared__ double warp_sum[8];
  const int e = blockIdx.x, t = threadIdx.x;
  double vol = 1;
  for (int d = 0; d < ND; ++d) vol *= nom[e];
  double local = 0;
  for (int q = t; q < nq; q += blockDim.x) {
    double* s = state + (size_t)e*(ND + 2)*nq + q;
    double* a = av + (size_t)e*2*nq + q; // bulk AV coefficient
    const double f = mult*forcing[((size_t)e*4 + n_real)*nq + q];
    const double new_av = us_max*f/(us_max + f);
    double wq = 1; // math::pow_outer(node_weights, n_dim)
    #pragma unroll
    for (int d = 0; d < ND; ++d) wq *= p.w[(q/ipow(RS, ND - 1 - d)) % RS];
    const double diff = *a - new_av;
    local += diff*diff*wq*vol;
    *a = new_av;
    const double scale = sqrt(2*s[ND*nq]*s[(ND + 1)*nq]);
    #pragma unroll
    for (int d = 0; d < ND; ++d) s[d*nq] *= scale;
  }
  #pragma unroll
  for (int off = 16; off > 0; off /= 2) local += __shfl_xor_sync(0xffffffffu, local, off);
  if (t % 32 == 0) warp_sum[t/32] = local;
  __syncthreads();
  if (t == 0) {
    double sum = 0;
    for (int i = 0; i < (int)(blockDim.x + 31)/32; ++i) sum += warp_sum[i];
    partial[e] = sum;
  }
}

__global__ void __launch_bounds__(1024)
sum_partials_kernel(const double* partial, int n, double* out)
{
  __shared__ double s[1024];
  double local = 0;
  for (int i = threadIdx.x; i < n; i += 1024) local += partial[i];
  s[threadIdx.x] = local;
  __syncthreads();
  for (int off = 512; off > 0; off /= 2) {
    if ((int)threadIdx.x < off) s[threadIdx.x] += s[threadIdx.x + off];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = s[0];
}

/* dst[e][q] = sum over the 2^ND vertices of prod_d interp[coord_d(q)][bit_d] * vert[e][vertex], contracted innermost dimension first */
template <int ND, int RS>
__global__ void __launch_bounds__(256)
interp_vertices_kernel(const double* vert, double* dst, size_t dst_elem_stride, long long n_point_total, const double* interp /* [RS][2] */)
{
  constexpr int nq = ipow(RS, ND), n_vert = ipow(2, ND);
  const long long gid = (long long)blockIdx.x*blockDim.x + threadIdx.x;
  if (gid >= n_point_total) return;
  const long long e = gid/nq; const int q = (int)(gid % nq);
  double vals[n_vert];
  #pragma unroll
  for (int v = 0; v < n_vert; ++v) vals[v] = vert[(size_t)e*n_vert + v];
  int n = n_vert;
  #pragma unroll
  for (int d = ND - 1; d >= 0; --d) {
    const int node = (q/ipow(RS, ND - 1 - d)) % RS;
    const double c0 = interp[node*2], c1 = interp[node*2 + 1];
    n /= 2;
    #pragma unroll
    for (int m = 0; m < n_vert/2; ++m) if (m < n) vals[m] = c0*vals[2*m] + c1*vals[2*m + 1];
  }
  dst[(size_t)e*dst_elem_stride + q] = vals[0];
}

__global__ void __launch_bounds__(256)
av_swap_kernel(double* av, long long n_point_total, int nq)
{
  const long long gid = (long long)blockIdx.x*blockDim.x + threadIdx.x;
  if (gid >= n_point_total) return;
  const long long e = gid/nq; const int q = (int)(gid % nq);
  double* a = av + (size_t)e*2*nq + q;
  const double t = a[0]; a[0] = a[nq]; a[nq] = t;
}

/* ---------------- Solver::update_art_visc_elwise (reference src/Solver.cpp:584-633) ----------------
 * the scalar ramp that turns Element::uncertainty (a non-smoothness indicator) into an element-wise viscosity (:591-601) ... */
__global__ void __launch_bounds__(256)
av_elwise_ramp_kernel(double* uncert, int n_elem, double ramp_center, double scale)
{
  const int e = blockIdx.x*blockDim.x + threadIdx.x;
  if (e >= n_elem) return;
  double u = uncert[e];
  u = 2*log(u)/log(10.);
  const double half_width = 0.5;
  if (!(u > ramp_center - half_width)) u = 0;
  else if (u >= ramp_center + half_width) u = 1;
  else u = .5*(1 + sin(3.14159265358979323846*(u - ramp_center)/2/half_width));
  uncert[e] = u*scale;
}

/* ... and the point loops of its PDE-based branch: forcing[0] = element value, forcing[1] = laplacian_av_coef (:603-612, dir 0);
 * laplacian_av_coef = forcing[1] after the diffusion (:614-619, dir 1) */
__global__ void __launch_bounds__(256)
av_elwise_forcing_kernel(const double* uncert, double* av, double* forcing, long long n_point_total, int nq, int dir)
{
  const long long gid = (long long)blockIdx.x*blockDim.x + threadIdx.x;
  if (gid >= n_point_total) return;
  const long long e = gid/nq; const int q = (int)(gid % nq);
  double* lap = av + ((size_t)e*2 + 1)*nq + q;
  double* f = forcing + (size_t)e*4*nq + q;
  if (dir == 0) { f[0] = uncert[e]; f[nq] = *lap; }
  else *lap = f[nq];
}

/* the element's value on each of its vertices (the `get` functor of the share_vertex_data call at :621-623) */
__global__ void __launch_bounds__(256)
elem_value_to_vertices_kernel(const double* vals, double* elem_vals, long long n, int n_vert)
{
  const long long i = (long long)blockIdx.x*blockDim.x + threadIdx.x;
  if (i < n) elem_vals[i] = vals[i/n_vert];
}

static int need_array(hexed_b200_ctx* c, double** arr, size_t per_elem)
{
  if (*arr) return 0;
  const size_t n = (size_t)(c->n_elem ? c->n_elem : 1)*per_elem;
  HB_CUDA(c, cudaMalloc(arr, sizeof(double)*n));
  HB_CUDA(c, cudaMemsetAsync(*arr, 0, sizeof(double)*n, c->stream));
  return 0;
}

static int grid_for(long long n) { return (int)((n + 255)/256); }

int launch_av_scale_velocity(hexed_b200_ctx* c, int restore)
{
  if (!c->have_mesh) return fail(c, HEXED_B200_NO_MESH, "no mesh uploaded");
  invalidate_cfl_cache(c);
  const long long n = (long long)c->n_elem*c->nq;
  if (!n) return 0;
  HB_LAUNCH(av_scale_velocity_kernel, grid_for(n), 256, 0, c->stream, c->state, n, c->nd, c->nq, restore);
  ++c->launches;
  HB_CUDA(c, cudaGetLastError());
  return 0;
}

int launch_av_project_forcing(hexed_b200_ctx* c, const double* weights, const double* orth)
{
  if (!c->have_mesh) return fail(c, HEXED_B200_NO_MESH, "no mesh uploaded");
  int rc = need_array(c, &c->forcing, (size_t)4*c->nq); if (rc) return rc;
  rc = need_array(c, &c->adv, (size_t)c->rs*c->nq); if (rc) return rc;
  ProjArgs p;
  for (int i = 0; i < MAX_RS; ++i) { p.w[i] = i < c->rs ? weights[i] : 0.; p.orth[i] = i < c->rs ? orth[i] : 0.; }
  const long long n = (long long)c->n_elem*c->nq;
  if (!n) return 0;
  HB_LAUNCH(av_project_forcing_kernel, grid_for(n), 256, 0, c->stream, c->state, c->adv, c->forcing, n, c->nd, c->nq, c->rs, p);
  ++c->launches;
  HB_CUDA(c, cudaGetLastError());
  return 0;
}

int launch_av_finish(hexed_b200_ctx* c, double mult, double us_max, int n_real, const double* node_weights, double* resid_sq)
{
  if (!c->have_mesh) return fail(c, HEXED_B200_NO_MESH, "no mesh uploaded");
  if (n_real < 0 || n_real > 3) return fail(c, HEXED_B200_BAD_ARGUMENT, "forcing slot out of range");
  int rc = need_array(c, &c->forcing, (size_t)4*c->nq); if (rc) return rc;
  rc = need_array(c, &c->av, (size_t)2*c->nq); if (rc) return rc;
  invalidate_cfl_cache(c);
  *resid_sq = 0.;
  if (!c->n_elem) return 0;
  ProjArgs p;
  for (int i = 0; i < MAX_RS; ++i) { p.w[i] = i < c->rs ? node_weights[i] : 0.; p.orth[i] = 0.; }
  DevScratch<double> scratch;
  HB_CUDA(c, scratch.alloc(c->n_elem));
  double* partial = scratch.p;
  rc = dispatch(c, [&](auto nd, auto rs) {
    constexpr int ND = decltype(nd)::value, RS = decltype(rs)::value;
    constexpr int nq = ipow(RS, ND);
    const int threads = nq >= 256 ? 256 : ((nq + 31)/32)*32;
    auto k = av_finish_kernel<ND, RS>;
    HB_LAUNCH(k, c->n_elem, threads, 0, c->stream, c->state, c->av, c->forcing, c->nom, c->n_elem, mult, us_max, n_real, p, partial);
    HB_LAUNCH(sum_partials_kernel, 1, 1024, 0, c->stream, partial, c->n_elem, c->d_scalar);
    c->launches += 2;
    HB_CUDA(c, cudaGetLastError());
    return 0;
  });
  if (!rc) rc = check(c, cudaMemcpyAsync(c->h_scalar, c->d_scalar, sizeof(double), cudaMemcpyDeviceToHost, c->stream), "read residual");
  if (!rc) rc = check(c, cudaStreamSynchronize(c->stream), "av_finish");
  if (!rc) *resid_sq = *c->h_scalar;
  return rc;
}

int launch_interp_vertices(hexed_b200_ctx* c, int target, const double* d_vert, const double* d_interp)
{
  if (!c->have_mesh) return fail(c, HEXED_B200_NO_MESH, "no mesh uploaded");
  int rc = need_array(c, &c->av, (size_t)2*c->nq); if (rc) return rc;
  if (target != 0 && target != 1) return fail(c, HEXED_B200_BAD_ARGUMENT, "target must be 0 (bulk_av_coef) or 1 (laplacian_av_coef)");
  const long long n = (long long)c->n_elem*c->nq;
  if (!n) return 0;
  double* dst = c->av + (size_t)target*c->nq;
  return dispatch(c, [&](auto nd, auto rs) {
    constexpr int ND = decltype(nd)::value, RS = decltype(rs)::value;
    auto k = interp_vertices_kernel<ND, RS>;
    HB_LAUNCH(k, grid_for(n), 256, 0, c->stream, d_vert, dst, (size_t)2*c->nq, n, d_interp);
    ++c->launches;
    HB_CUDA(c, cudaGetLastError());
    return 0;
  });
}

int launch_av_swap(hexed_b200_ctx* c)
{
  if (!c->have_mesh) return fail(c, HEXED_B200_NO_MESH, "no mesh uploaded");
  int rc = need_array(c, &c->av, (size_t)2*c->nq); if (rc) return rc;
  const long long n = (long long)c->n_elem*c->nq;
  if (!n) return 0;
  HB_LAUNCH(av_swap_kernel, grid_for(n), 256, 0, c->stream, c->av, n, c->nq);
  ++c->launches;
  HB_CUDA(c, cudaGetLastError());
  return 0;
}

/* ---------------- Solver::share_vertex_data (reference src/Solver.cpp:35-54) ----------------
 * Every mesh vertex reduces (min or max) the values its elements hold for it (Element::push_shareable_value / fetch_shareable_value,
 * src/Element.cpp:149-161), then every Hanging_vertex_matcher overwrites the vertices of its fine elements that lie on the coarse
 * face with the multilinear interpolant of the coarse corners (src/Hanging_vertex_matcher.cpp:13-41). The vertex connectivity is not
 * part of Kernel_mesh: the caller provides it once per mesh epoch (hexed_b200_vertex_topology). */
__device__ __forceinline__ void atomic_minmax(double* addr, double v, int is_max)
{
  unsigned long long* a = reinterpret_cast<unsigned long long*>(addr);
  unsigned long long old = *a;
  while (true) {
    const double cur = __longlong_as_double((long long)old);
    if (is_max ? !(v > cur) : !(v < cur)) return;
    const unsigned long long prev = atomicCAS(a, old, (unsigned long long)__double_as_longlong(v));
    if (prev == old) return;
    old = prev;
  }
}

__global__ void __launch_bounds__(256)
fill_double_kernel(double* dst, int n, double value)
{
  const int i = blockIdx.x*blockDim.x + threadIdx.x;
  if (i < n) dst[i] = value;
}

__global__ void __launch_bounds__(256)
vertex_reduce_kernel(const double* elem_vals, const int* elem_vertex, long long n, double* vertex_vals, int is_max)
{
  const long long i = (long long)blockIdx.x*blockDim.x + threadIdx.x;
  if (i < n) atomic_minmax(vertex_vals + elem_vertex[i], elem_vals[i], is_max);
}

__global__ void __launch_bounds__(256)
vertex_fetch_kernel(double* elem_vals, const int* elem_vertex, long long n, const double* vertex_vals)
{
  const long long i = (long long)blockIdx.x*blockDim.x + threadIdx.x;
  if (i < n) elem_vals[i] = vertex_vals[elem_vertex[i]];
}

/* one thread per Hanging_vertex_matcher; row = {i_dim, is_positive, stretch0, stretch1, fine element 0..3 (-1 = unused)} */
__global__ void __launch_bounds__(128)
hanging_vertex_kernel(double* elem_vals, const int* matchers, int n_match, int nd)
{
  const int im = blockIdx.x*blockDim.x + threadIdx.x;
  if (im >= n_match) return;
  const int* row = matchers + (size_t)im*8;
  const int id = row[0], isp = row[1];
  const int str[2] = {row[2], row[3]};
  const int n_vert = 1 << (nd - 1), n_vert_elem = 1 << nd;
  int n_fine = 0;
  for (int i = 0; i < 4; ++i) if (row[4 + i] >= 0) ++n_fine;
  int inds[4]; double values[4];
  const int stride = 1 << (nd - 1 - id);
  for (int iv = 0; iv < n_vert; ++iv) {
    inds[iv] = iv/stride*stride*2 + iv % stride + isp*stride;
    // math::stretched_ind (include/math.hpp:188-199): collapse the face-vertex index along stretched dimensions
    int si = 0, st = 1;
    for (int d = nd - 2; d >= 0; --d) {
      const int bit = (iv >> (nd - 2 - d)) & 1;
      if (!str[d]) { si += bit*st; st *= 2; }
    }
    values[iv] = elem_vals[(size_t)row[4 + si]*n_vert_elem + inds[iv]];
  }
  // hypercube_matvec of the 3 x 2 matrix {{1, 0}, {.5, .5}, {0, 1}}: corners and midpoints, 3 [x 3] values
  double interp[9];
  if (nd == 1) interp[0] = values[0];
  else if (nd == 2) { interp[0] = values[0]; interp[1] = .5*values[0] + .5*values[1]; interp[2] = values[1]; }
  else {
    double rowv[2][3];
    for (int a = 0; a < 2; ++a) { rowv[a][0] = values[2*a]; rowv[a][1] = .5*values[2*a] + .5*values[2*a + 1]; rowv[a][2] = values[2*a + 1]; }
    for (int b = 0; b < 3; ++b) { interp[b] = rowv[0][b]; interp[3 + b] = .5*rowv[0][b] + .5*rowv[1][b]; interp[6 + b] = rowv[1][b]; }
  }
  for (int ie = 0; ie < n_fine; ++ie) {
    for (int iv = 0; iv < n_vert; ++iv) {
      int k = nd >= 2 ? (ie*!str[nd - 2]) % 2 + (iv % 2)*(1 + str[nd - 2]) : 0;
      if (nd == 3) k += ((ie*!str[0])/(1 + !str[1]) + iv/2*(1 + str[0]))*3;
      elem_vals[(size_t)row[4 + ie]*n_vert_elem + inds[iv]] = interp[k];
    }
  }
}

/* element's own record -> every vertex of the element (Solver.cpp:1000-1006) */
__global__ void __launch_bounds__(256)
record_to_vertices_kernel(const int* record, double* elem_vals, long long n, int n_vert)
{
  const long long i = (long long)blockIdx.x*blockDim.x + threadIdx.x;
  if (i < n) elem_vals[i] = record[i/n_vert];
}

/* every vertex of an element takes the element's maximum (Solver.cpp:1008-1018) */
__global__ void __launch_bounds__(256)
element_max_kernel(double* elem_vals, int n_elem, int n_vert)
{
  const int e = blockIdx.x*blockDim.x + threadIdx.x;
  if (e >= n_elem) return;
  double m = 0;
  for (int v = 0; v < n_vert; ++v) m = fmax(m, elem_vals[(size_t)e*n_vert + v]);
  for (int v = 0; v < n_vert; ++v) elem_vals[(size_t)e*n_vert + v] = m;
}

int launch_share_vertex_data(hexed_b200_ctx* c, double* elem_vals, int is_max)
{
  if (!c->have_mesh) return fail(c, HEXED_B200_NO_MESH, "no mesh uploaded");
  if (!c->elem_vertex) return fail(c, HEXED_B200_BAD_ARGUMENT, "no vertex topology: call hexed_b200_vertex_topology first");
  const long long n = (long long)c->n_elem*c->n_vert;
  if (!n) return 0;
  HB_LAUNCH(fill_double_kernel, (c->n_vertex + 255)/256, 256, 0, c->stream, c->vertex_vals, c->n_vertex, is_max ? -DBL_MAX : DBL_MAX);
  HB_LAUNCH(vertex_reduce_kernel, grid_for(n), 256, 0, c->stream, elem_vals, c->elem_vertex, n, c->vertex_vals, is_max);
  HB_LAUNCH(vertex_fetch_kernel, grid_for(n), 256, 0, c->stream, elem_vals, c->elem_vertex, n, c->vertex_vals);
  c->launches += 3;
  if (c->n_match) {
    // matchers of one epoch touch disjoint fine elements except through shared vertices, which they set to the same interpolant
    HB_LAUNCH(hanging_vertex_kernel, (c->n_match + 127)/128, 128, 0, c->stream, elem_vals, c->matchers, c->n_match, c->nd);
    ++c->launches;
  }
  HB_CUDA(c, cudaGetLastError());
  return 0;
}

int launch_fix_admis_spread(hexed_b200_ctx* c, const double* d_interp)
{
  if (!c->record) return fail(c, HEXED_B200_BAD_ARGUMENT, "no record yet: call hexed_b200_is_admissible first");
  int rc = need_array(c, &c->vertex_scratch, (size_t)c->n_vert); if (rc) return rc;
  const long long n = (long long)c->n_elem*c->n_vert;
  if (!n) return 0;
  HB_LAUNCH(record_to_vertices_kernel, grid_for(n), 256, 0, c->stream, c->record, c->vertex_scratch, n, c->n_vert);
  ++c->launches;
  rc = launch_share_vertex_data(c, c->vertex_scratch, 1); if (rc) return rc;
  HB_LAUNCH(element_max_kernel, (c->n_elem + 255)/256, 256, 0, c->stream, c->vertex_scratch, c->n_elem, c->n_vert);
  ++c->launches;
  rc = launch_share_vertex_data(c, c->vertex_scratch, 1); if (rc) return rc;
  rc = launch_interp_vertices(c, 1, c->vertex_scratch, d_interp); if (rc) return rc;
  return launch_av_swap(c);
}

int launch_av_elwise_ramp(hexed_b200_ctx* c, double scale)
{
  if (!c->have_mesh) return fail(c, HEXED_B200_NO_MESH, "no mesh uploaded");
  if (!c->n_elem) return 0;
  const double ramp_center = -4 - 4.25*std::log(double(c->rs - 1))/std::log(10.); // Solver.cpp:595
  HB_LAUNCH(av_elwise_ramp_kernel, (c->n_elem + 255)/256, 256, 0, c->stream, c->uncert, c->n_elem, ramp_center, scale);
  ++c->launches;
  HB_CUDA(c, cudaGetLastError());
  return 0;
}

int launch_av_elwise_forcing(hexed_b200_ctx* c, int dir)
{
  if (!c->have_mesh) return fail(c, HEXED_B200_NO_MESH, "no mesh uploaded");
  int rc = need_array(c, &c->av, (size_t)2*c->nq); if (rc) return rc;
  rc = need_array(c, &c->forcing, (size_t)4*c->nq); if (rc) return rc;
  const long long n = (long long)c->n_elem*c->nq;
  if (!n) return 0;
  HB_LAUNCH(av_elwise_forcing_kernel, grid_for(n), 256, 0, c->stream, c->uncert, c->av, c->forcing, n, c->nq, dir);
  ++c->launches;
  HB_CUDA(c, cudaGetLastError());
  return 0;
}

/* the vertex-based branch (:620-632): element value -> its vertices, share_vertex_data(max), multilinear interpolation to laplacian_av_coef */
int launch_av_elwise_vertices(hexed_b200_ctx* c, const double* d_interp)
{
  if (!c->have_mesh) return fail(c, HEXED_B200_NO_MESH, "no mesh uploaded");
  int rc = need_array(c, &c->vertex_scratch, (size_t)c->n_vert); if (rc) return rc;
  const long long n = (long long)c->n_elem*c->n_vert;
  if (!n) return 0;
  HB_LAUNCH(elem_value_to_vertices_kernel, grid_for(n), 256, 0, c->stream, c->uncert, c->vertex_scratch, n, c->n_vert);
  ++c->launches;
  rc = launch_share_vertex_data(c, c->vertex_scratch, 1); if (rc) return rc;
  return launch_interp_vertices(c, 1, c->vertex_scratch, d_interp);
}

} // namespace hb
