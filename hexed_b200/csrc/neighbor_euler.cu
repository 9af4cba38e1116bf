/* neighbor_euler.cu -- shared numerical flux (local Lax-Friedrichs) on conforming, hanging-node (fine/mortar) and
 * boundary face pairs. Replaces Spatial<Navier_stokes<false>::Pde, is_deformed>::Neighbor
 * (reference include/Spatial.hpp:613-704).
 *
 * One thread per (connection, face quadrature point); both sides' states are gathered through the connection table.
 * For deformed connections the reordering of side 1 (`Face_permutation::match_faces`/`restore`,
 * include/Spatial.hpp:85-129) is a precomputed index table per connection direction instead of an in-place
 * transpose/flip, and the sign flips of `Connection_direction::flip_normal` are applied on the fly.
 * HBM traffic per connection: R 2 nv nfq (+ n_dim nfq normal), W 2 nv nfq.
 */
#include "euler.cuh"

namespace hb {

struct NeighborArgs
{
  double* faces; const double* normals; const int* con; const int* perm; int n_con;
};

template <int ND, int RS, bool DEF>
__global__ void __launch_bounds__(256, 4)
neighbor_euler_kernel(NeighborArgs a)
{
  constexpr int nfq = ipow(RS, ND - 1), nv = ND + 2, w = nv*nfq;
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
    // flip_normal(side) = (face_sign[side] == side)   (reference include/Kernel_connection.hpp:19)
    sign0 = ((code/9) % 2) ? 1. : -1.;
    sign1 = ((code/18) % 2) ? -1. : 1.;
    const double* nr = a.normals + (size_t)tab[3]*ND*nfq;
    #pragma unroll
    for (int d = 0; d < ND; ++d) n[d] = sign0*nr[d*nfq + q0];
  } else {
    #pragma unroll
    for (int d = 0; d < ND; ++d) n[d] = (d == tab[2]) ? 1. : 0.;
  }
  double* f0 = a.faces + (size_t)slot0*w + q0;
  double* f1 = a.faces + (size_t)slot1*w + q1;
  EulerPoint<ND> p0, p1;
  #pragma unroll
  for (int v = 0; v < nv; ++v) { p0.s[v] = f0[v*nfq]; p1.s[v] = f1[v*nfq]; }
  p0.scalars(); p1.scalars();
  double fl0[nv], fl1[nv];
  p0.flux(n, fl0); p1.flux(n, fl1);
  double nsq = 0;
  #pragma unroll
  for (int d = 0; d < ND; ++d) nsq += n[d]*n[d];
  const double speed = fmax(p0.char_speed(), p1.char_speed())*sqrt(nsq);
  #pragma unroll
  for (int v = 0; v < nv; ++v) {
    const double flux = .5*(fl0[v] + fl1[v] + speed*(p0.s[v] - p1.s[v]));
    f0[v*nfq] = sign0*flux;
    f1[v*nfq] = sign1*flux;
  }
}

int launch_neighbor_euler(hexed_b200_ctx* c, int deformed, int first, int count)
{
  if (!c->have_mesh) return fail(c, HEXED_B200_NO_MESH, "no mesh uploaded");
  const int n_all = deformed ? c->n_def_con : c->n_car_con;
  if (count < 0) count = n_all - first;
  if (first < 0 || first + count > n_all) return fail(c, HEXED_B200_BAD_ARGUMENT, "connection range out of bounds");
  const int n_con = count;
  StatScope scope(c, deformed ? ST_NEIGHBOR_DEF : ST_NEIGHBOR_CAR, n_con);
  if (!n_con) return 0;
  invalidate_admis(c); // the element faces now hold numerical fluxes
  NeighborArgs a;
  a.faces = c->face_state; a.normals = c->normals; a.con = (deformed ? c->def_con : c->car_con) + (size_t)first*4; a.perm = c->perm; a.n_con = n_con;
  return dispatch(c, [&](auto nd, auto rs) {
    constexpr int ND = decltype(nd)::value, RS = decltype(rs)::value;
    const long long total = (long long)n_con*ipow(RS, ND - 1);
    const int grid = (int)((total + 255)/256);
    if (deformed) { auto k = neighbor_euler_kernel<ND, RS, true>; HB_LAUNCH(k, grid, 256, 0, c->stream, a); }
    else { auto k = neighbor_euler_kernel<ND, RS, false>; HB_LAUNCH(k, grid, 256, 0, c->stream, a); }
    count_launch(c, deformed ? ST_NEIGHBOR_DEF : ST_NEIGHBOR_CAR);
    HB_CUDA(c, cudaGetLastError());
    return 0;
  });
}

} // namespace hb
