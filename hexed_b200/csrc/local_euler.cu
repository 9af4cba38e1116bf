/* local_euler.cu -- volume residual + two-stage update + face extrapolation for the Euler equations.
 *
 * Replaces Spatial<Navier_stokes<false>::Pde, is_deformed>::Local (reference include/Spatial.hpp:326-509) including the
 * trailing write_face (include/Spatial.hpp:41-57,507).
 *
 * Generic kernel (any n_dim 1..3, row_size 2..8): one thread per quadrature point, `epb` elements per CTA so that small
 * elements still fill a CTA. Per element in shared memory: the pointwise flux [n_dim][nv][nq] (later reused for the
 * updated state) and the element's 2*n_dim numerical-flux faces [2 n_dim][nv][nfq]. The sum-factorised derivative is a
 * row_size-term dot product per (dimension, variable) read from shared memory; the 1-D operator rows of this thread's
 * node are kept in registers.
 *
 * HBM traffic per element (Euler): R state nv*nq + tss nq + faces 2 n_dim nv nfq + cache nv*nq (stage 1)
 *                                  W state nv*nq + cache nv*nq (stage 0) + faces 2 n_dim nv nfq
 *                       deformed: + R normals n_dim^2 nq + det nq   (face normals are not needed by the inviscid Local)
 */
#include "euler.cuh"

namespace hb {

template <int ND, int RS>
struct LocalCfg
{
  static constexpr int nq = ipow(RS, ND), nfq = nq/RS, nv = ND + 2;
  static constexpr int cs = nv > RS ? nv : RS; // slots per element in the residual cache array (src/Storage_params.cpp:32-35)
  static constexpr int epb = nq >= 128 ? 1 : (128 + nq - 1)/nq;
  static constexpr int threads = epb*nq;
  static constexpr int flux_doubles = ND*nv*nq;
  static constexpr int face_doubles = 2*ND*nv*nfq;
  static constexpr int ops_doubles = RS*RS + 2*RS; // dfull + lift
  static constexpr int smem_doubles = epb*(flux_doubles + face_doubles) + ops_doubles;
};

struct LocalArgs
{
  double* state; const double* tss; double* cache; const double* nom; const double* refn; const double* det; double* faces;
  int elem_begin, elem_end, n_car;
  double update; int stage; int compute_residual; int use_filter;
  const double* dt_dev; // non-null: the time step lives on the device (hexed_b200_update_euler) and multiplies `update`
};

template <int ND, int RS, bool DEF>
__global__ void __launch_bounds__(LocalCfg<ND, RS>::threads)
local_euler_kernel(LocalArgs a, Ops ops, FilterOp filt)
{
  using C = LocalCfg<ND, RS>;
  constexpr int nq = C::nq, nfq = C::nfq, nv = C::nv;
  HB_DYN_SMEM(double, smem);
  const int t = threadIdx.x;
  const int le = t/nq, q = t % nq;
  const int e = a.elem_begin + blockIdx.x*C::epb + le;
  const bool active = e < a.elem_end;
  double* s_ops = smem;                                    // dfull[RS][RS], lift[RS][2]
  double* sflux = smem + C::ops_doubles + le*(C::flux_doubles + C::face_doubles); // [ND][nv][nq]
  double* sface = sflux + C::flux_doubles;                 // [2 ND][nv][nfq]

  for (int i = t; i < RS*RS; i += C::threads) s_ops[i] = ops.dfull[i/RS][i % RS];
  for (int i = t; i < 2*RS; i += C::threads) s_ops[RS*RS + i] = ops.lift[i/2][i % 2];

  // numerical flux of this element's faces -> shared
  if (active) {
    const double* f = a.faces + (size_t)e*C::face_doubles;
    for (int i = q; i < C::face_doubles; i += nq) sface[i] = f[i];
  }
  // pointwise flux -> shared
  EulerPoint<ND> p;
  double det = 1.;
  if (active) {
    #pragma unroll
    for (int v = 0; v < nv; ++v) p.s[v] = a.state[((size_t)e*nv + v)*nq + q];
    p.scalars();
    if constexpr (DEF) {
      const double* rn = a.refn + (size_t)(e - a.n_car)*ND*ND*nq;
      det = a.det[(size_t)(e - a.n_car)*nq + q];
      #pragma unroll
      for (int d = 0; d < ND; ++d) {
        double n[ND], f[nv];
        #pragma unroll
        for (int j = 0; j < ND; ++j) n[j] = rn[(d*ND + j)*nq + q];
        p.flux(n, f);
        #pragma unroll
        for (int v = 0; v < nv; ++v) sflux[(d*nv + v)*nq + q] = f[v];
      }
    } else {
      #pragma unroll
      for (int d = 0; d < ND; ++d) {
        double f[nv];
        p.flux_axis(d, f);
        #pragma unroll
        for (int v = 0; v < nv; ++v) sflux[(d*nv + v)*nq + q] = f[v];
      }
    }
  }
  __syncthreads();

  // residual: r = - sum_d D_d(flux_d, face flux_d)
  double r[nv];
  #pragma unroll
  for (int v = 0; v < nv; ++v) r[v] = 0.;
  if (active) {
    #pragma unroll
    for (int d = 0; d < ND; ++d) {
      constexpr int dummy = 0; (void)dummy;
      const int stride = ipow(RS, ND - 1 - d);
      const int node = (q/stride) % RS;
      const int base = q - node*stride;
      const int fq = (q/(stride*RS))*stride + q % stride;
      double m[RS], l0, l1;
      #pragma unroll
      for (int k = 0; k < RS; ++k) m[k] = s_ops[node*RS + k];
      l0 = s_ops[RS*RS + node*2]; l1 = s_ops[RS*RS + node*2 + 1];
      #pragma unroll
      for (int v = 0; v < nv; ++v) {
        const double* row = sflux + (d*nv + v)*nq + base;
        double acc = 0;
        #pragma unroll
        for (int k = 0; k < RS; ++k) acc += m[k]*row[k*stride];
        acc += l0*sface[((2*d)*nv + v)*nfq + fq];
        acc += l1*sface[((2*d + 1)*nv + v)*nfq + fq];
        r[v] -= acc;
      }
    }
  }
  __syncthreads(); // everyone is done with sflux; it is reused below

  // optional modal filter of the time rate along every dimension (reference include/Spatial.hpp:473-481)
  if (a.use_filter) {
    #pragma unroll
    for (int d = 0; d < ND; ++d) {
      if (active) {
        #pragma unroll
        for (int v = 0; v < nv; ++v) sflux[v*nq + q] = r[v];
      }
      __syncthreads();
      if (active) {
        const int stride = ipow(RS, ND - 1 - d);
        const int node = (q/stride) % RS;
        const int base = q - node*stride;
        #pragma unroll
        for (int v = 0; v < nv; ++v) {
          double acc = 0;
          for (int k = 0; k < RS; ++k) acc += filt.filter[node][k]*sflux[v*nq + base + k*stride];
          r[v] = acc;
        }
      }
      __syncthreads();
    }
  }

  // two-stage update (reference include/Spatial.hpp:311-324,484-503)
  if (active) {
    const double update = a.dt_dev ? *a.dt_dev*a.update : a.update;
    double mult = update*a.tss[(size_t)e*nq + q]/a.nom[e];
    if constexpr (DEF) mult /= det;
    double* cache = a.cache + (size_t)e*C::cs*nq + q;
    #pragma unroll
    for (int v = 0; v < nv; ++v) {
      double u = r[v];
      if (a.stage) u -= cache[v*nq];
      else if (!a.compute_residual) cache[v*nq] = u;
      u *= mult;
      if (a.compute_residual) cache[v*nq] = u;
      else p.s[v] += u;
    }
    if (!a.compute_residual) {
      #pragma unroll
      for (int v = 0; v < nv; ++v) a.state[((size_t)e*nv + v)*nq + q] = p.s[v];
    }
    // updated state -> shared for the face extrapolation
    #pragma unroll
    for (int v = 0; v < nv; ++v) sflux[v*nq + q] = p.s[v];
  }
  __syncthreads();

  // write_face: both faces of every row from one pass over the row
  if (active) {
    double* fout = a.faces + (size_t)e*C::face_doubles;
    for (int item = q; item < ND*nv*nfq; item += nq) {
      const int d = item/(nv*nfq), v = (item/nfq) % nv, fq = item % nfq;
      const int stride = ipow(RS, ND - 1 - d);
      const int base = (fq/stride)*stride*RS + fq % stride;
      double e0 = 0, e1 = 0;
      #pragma unroll
      for (int k = 0; k < RS; ++k) {
        const double x = sflux[v*nq + base + k*stride];
        e0 += ops.bnd[0][k]*x;
        e1 += ops.bnd[1][k]*x;
      }
      fout[((2*d)*nv + v)*nfq + fq] = e0;
      fout[((2*d + 1)*nv + v)*nfq + fq] = e1;
    }
  }
}

/* standalone face extrapolation: Spatial<..>::Write_face (reference include/Spatial.hpp:41-70) over all elements */
template <int ND, int RS>
__global__ void __launch_bounds__(LocalCfg<ND, RS>::threads)
write_face_kernel(const double* state, double* faces, int n_elem, Ops ops)
{
  using C = LocalCfg<ND, RS>;
  constexpr int nq = C::nq, nfq = C::nfq, nv = C::nv;
  HB_DYN_SMEM(double, smem);
  const int t = threadIdx.x;
  const int le = t/nq, q = t % nq;
  const int e = blockIdx.x*C::epb + le;
  const bool active = e < n_elem;
  double* sst = smem + le*nv*nq;
  if (active) {
    #pragma unroll
    for (int v = 0; v < nv; ++v) sst[v*nq + q] = state[((size_t)e*nv + v)*nq + q];
  }
  __syncthreads();
  if (active) {
    double* fout = faces + (size_t)e*C::face_doubles;
    for (int item = q; item < ND*nv*nfq; item += nq) {
      const int d = item/(nv*nfq), v = (item/nfq) % nv, fq = item % nfq;
      const int stride = ipow(RS, ND - 1 - d);
      const int base = (fq/stride)*stride*RS + fq % stride;
      double e0 = 0, e1 = 0;
      #pragma unroll
      for (int k = 0; k < RS; ++k) {
        const double x = sst[v*nq + base + k*stride];
        e0 += ops.bnd[0][k]*x;
        e1 += ops.bnd[1][k]*x;
      }
      fout[((2*d)*nv + v)*nfq + fq] = e0;
      fout[((2*d + 1)*nv + v)*nfq + fq] = e1;
    }
  }
}

int launch_local_euler(hexed_b200_ctx* c, int deformed, hexed_b200_options o)
{
  if (!c->have_mesh) return fail(c, HEXED_B200_NO_MESH, "no mesh uploaded");
  // same argument checks as the reference constructor (include/Spatial.hpp:322-323)
  if (o.i_stage && o.compute_residual) return fail(c, HEXED_B200_BAD_ARGUMENT, "residual calculation is a single-stage operation");
  const int begin = deformed ? c->n_car : 0, end = deformed ? c->n_elem : c->n_car;
  StatScope scope(c, deformed ? ST_LOCAL_DEF : ST_LOCAL_CAR, end - begin);
  if (end == begin) return 0;
  {
    int rc = launch_local_euler_pipe(c, deformed, o, begin, end);
    if (rc == -1) rc = launch_local_euler_pipe2d(c, deformed, o, begin, end);
    if (rc == 0) count_launch(c, deformed ? ST_LOCAL_DEF : ST_LOCAL_CAR);
    if (rc >= 0) return rc;
  }
  c->cfl_valid[deformed ? 1 : 0] = false; // the general kernel rewrites the state without leaving CFL ratios behind
  invalidate_admis(c);
  LocalArgs a;
  a.state = c->state; a.tss = c->tss; a.cache = c->cache; a.nom = c->nom; a.refn = c->refn; a.det = c->det; a.faces = c->face_state;
  a.elem_begin = begin; a.elem_end = end; a.n_car = c->n_car;
  a.update = o.i_stage ? o.dt*(.5/c->quad_safety) : o.dt; // Spatial.hpp:317 with Basis::step_ratio (src/Basis.cpp:11-14)
  a.dt_dev = c->dt_dev_active;
  a.stage = o.i_stage != 0; a.compute_residual = o.compute_residual; a.use_filter = o.use_filter;
  return dispatch(c, [&](auto nd, auto rs) {
    constexpr int ND = decltype(nd)::value, RS = decltype(rs)::value;
    using C = LocalCfg<ND, RS>;
    const int grid = (end - begin + C::epb - 1)/C::epb;
    const size_t smem = sizeof(double)*C::smem_doubles;
    if (deformed) {
      auto k = local_euler_kernel<ND, RS, true>;
      if (smem > 48*1024) HB_CUDA(c, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      HB_LAUNCH(k, grid, C::threads, smem, c->stream, a, c->ops, c->filt);
    } else {
      auto k = local_euler_kernel<ND, RS, false>;
      if (smem > 48*1024) HB_CUDA(c, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      HB_LAUNCH(k, grid, C::threads, smem, c->stream, a, c->ops, c->filt);
    }
    count_launch(c, deformed ? ST_LOCAL_DEF : ST_LOCAL_CAR);
    HB_CUDA(c, cudaGetLastError());
    return 0;
  });
}

int launch_write_face(hexed_b200_ctx* c)
{
  if (!c->have_mesh) return fail(c, HEXED_B200_NO_MESH, "no mesh uploaded");
  StatScope scope(c, ST_WRITE_FACE, c->n_elem);
  if (!c->n_elem) return 0;
  invalidate_admis(c);
  return dispatch(c, [&](auto nd, auto rs) {
    constexpr int ND = decltype(nd)::value, RS = decltype(rs)::value;
    using C = LocalCfg<ND, RS>;
    const int grid = (c->n_elem + C::epb - 1)/C::epb;
    const size_t smem = sizeof(double)*C::epb*C::nv*C::nq;
    auto k = write_face_kernel<ND, RS>;
    if (smem > 48*1024) HB_CUDA(c, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    HB_LAUNCH(k, grid, C::threads, smem, c->stream, c->state, c->face_state, c->n_elem, c->ops);
    count_launch(c, ST_WRITE_FACE);
    HB_CUDA(c, cudaGetLastError());
    return 0;
  });
}

} // namespace hb
