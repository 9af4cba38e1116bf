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
  double* partial = nullptr;
  HB_CUDA(c, cudaMalloc(&partial, sizeof(double)*c->n_elem));
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
  cudaFree(partial);
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

} // namespace hb
