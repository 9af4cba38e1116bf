/* metrics.cu -- metric terms of deformed elements on the device.
 *
 * Replaces the per-element loop of Solver::calc_jacobian (reference src/Solver.cpp:281-286):
 *   Deformed_element::position   src/Deformed_element.cpp:15-58   multilinear vertex interpolation + face-warping node adjustments
 *   Deformed_element::set_jacobian   :60-136   Jacobian by the basis derivative, reference-level normals (replaced-column
 *                                    determinants), determinant, element-face normals written into the face storage, vertex time-step scale
 *   Element::set_jacobian        src/Element.cpp:99-112   unit face Jacobian of Cartesian elements
 * After mesh adaptation only the vertex positions and node adjustments (a few hundred bytes per element) cross PCIe instead of the
 * 17 KB of metric terms per element the host would otherwise compute and upload.
 *
 * One CTA per deformed element; the element's positions, Jacobian and normals stay in shared memory between the steps.
 * Determinants are cofactor expansions (the reference calls Eigen's dynamic-size determinant(), an LU with partial pivoting: the
 * two differ in the last bits only).
 */
#include "common.cuh"

namespace hb {

template <int ND> __device__ __forceinline__ double det_small(const double (&m)[ND][ND])
{
  if constexpr (ND == 1) return m[0][0];
  else if constexpr (ND == 2) return m[0][0]*m[1][1] - m[0][1]*m[1][0];
  else return m[0][0]*(m[1][1]*m[2][2] - m[1][2]*m[2][1]) - m[0][1]*(m[1][0]*m[2][2] - m[1][2]*m[2][0])
            + m[0][2]*(m[1][0]*m[2][1] - m[1][1]*m[2][0]);
}

/* determinant of m with column `col` replaced by the unit vector e_row (reference: copy(all, i).setUnit(j); copy.determinant()) */
template <int ND> __device__ __forceinline__ double det_unit_column(const double (&m)[ND][ND], int col, int row)
{
  double c[ND][ND];
  #pragma unroll
  for (int a = 0; a < ND; ++a)
    #pragma unroll
    for (int b = 0; b < ND; ++b) c[a][b] = b == col ? (a == row ? 1. : 0.) : m[a][b];
  return det_small<ND>(c);
}

struct MetricArgs
{
  const double* vert;     // [n_def][2^ND][ND]
  const double* node_adj; // [n_def][2 ND][nfq] or nullptr
  const double* nom; double* refn; double* det; double* faces; double* vtss;
  int n_car, n_def, face_width;
};

template <int ND, int RS>
__global__ void __launch_bounds__(256)
set_jacobian_kernel(MetricArgs a, Ops ops)
{
  constexpr int nq = ipow(RS, ND), nfq = nq/RS, n_vert = ipow(2, ND);
  HB_DYN_SMEM(double, smem);
  double* s_pos = smem;                 // [ND][nq]
  double* s_jac = s_pos + ND*nq;        // [ND*ND][nq]  (i*ND + j): d pos_i / d ref_j / nom
  double* s_nrm = s_jac + ND*ND*nq;     // [ND*ND + 1][nq] reference-level normals, determinant last
  double* s_vert = s_nrm + (ND*ND + 1)*nq; // [n_vert][ND]
  double* s_vx = s_vert + n_vert*ND;    // [ND*ND + 1][n_vert] fields extrapolated to the vertices
  const int t = threadIdx.x, T = blockDim.x;
  const int ed = blockIdx.x;
  if (ed >= a.n_def) return;
  const int e = a.n_car + ed;
  const double nom = a.nom[e];
  for (int i = t; i < n_vert*ND; i += T) s_vert[i] = a.vert[(size_t)ed*n_vert*ND + i];
  __syncthreads();

  /* positions */
  for (int q = t; q < nq; q += T) {
    double w[ND], s_adj[ND];
    #pragma unroll
    for (int d = 0; d < ND; ++d) {
      const int stride = ipow(RS, ND - 1 - d);
      w[d] = ops.node[(q/stride) % RS];
      // face quadrature point in the same row of dimension d
      const int fq = (q/(stride*RS))*stride + q % stride;
      const double a0 = a.node_adj ? a.node_adj[((size_t)ed*2*ND + 2*d)*nfq + fq] : 0.;
      const double a1 = a.node_adj ? a.node_adj[((size_t)ed*2*ND + 2*d + 1)*nfq + fq] : 0.;
      s_adj[d] = a0*(1. - w[d]) + a1*w[d];
    }
    #pragma unroll
    for (int i = 0; i < ND; ++i) {
      // successive contraction of the 2^ND vertex values, last dimension first. k = ND: plain interpolation; k < ND: the adjustment
      // of reference direction k, where the contraction along k is the scaled difference [-s, s] instead of [1 - w, w]
      auto contract = [&](int k) {
        double vals[n_vert];
        #pragma unroll
        for (int v = 0; v < n_vert; ++v) vals[v] = s_vert[v*ND + i];
        int n = n_vert;
        #pragma unroll
        for (int j = ND - 1; j >= 0; --j) {
          n /= 2;
          const double c0 = j == k ? -s_adj[j] : 1. - w[j], c1 = j == k ? s_adj[j] : w[j];
          #pragma unroll
          for (int m = 0; m < n_vert/2; ++m) if (m < n) vals[m] = c0*vals[2*m] + c1*vals[2*m + 1];
        }
        return vals[0];
      };
      double total = contract(ND);
      #pragma unroll
      for (int k = 0; k < ND; ++k) total += contract(k); // added in the order of the dimensions, like the reference
      s_pos[i*nq + q] = total;
    }
  }
  __syncthreads();

  /* Jacobian: derivative of the position along every reference direction */
  for (int item = t; item < ND*ND*nq; item += T) {
    const int q = item % nq, ij = item/nq, i = ij/ND, j = ij % ND;
    const int stride = ipow(RS, ND - 1 - j);
    const int node = (q/stride) % RS, base = q - node*stride;
    double acc = 0.;
    #pragma unroll
    for (int m = 0; m < RS; ++m) acc += ops.diff[node][m]*s_pos[i*nq + base + m*stride];
    s_jac[ij*nq + q] = acc/nom;
  }
  __syncthreads();

  /* reference-level normals and determinant */
  for (int q = t; q < nq; q += T) {
    double J[ND][ND];
    #pragma unroll
    for (int i = 0; i < ND; ++i)
      #pragma unroll
      for (int j = 0; j < ND; ++j) J[i][j] = s_jac[(i*ND + j)*nq + q];
    const double dt = det_small<ND>(J);
    s_nrm[ND*ND*nq + q] = dt;
    a.det[(size_t)ed*nq + q] = dt;
    #pragma unroll
    for (int i = 0; i < ND; ++i)
      #pragma unroll
      for (int j = 0; j < ND; ++j) {
        const double n = det_unit_column<ND>(J, i, j);
        s_nrm[(i*ND + j)*nq + q] = n;
        a.refn[((size_t)ed*ND*ND + i*ND + j)*nq + q] = n;
      }
  }

  /* element-face normals into the face storage (first ND*nfq doubles of every face) */
  for (int item = t; item < 2*ND*nfq; item += T) {
    const int f = item/nfq, fq = item % nfq, d = f/2, sign = f % 2;
    const int stride = ipow(RS, ND - 1 - d);
    const int base = (fq/stride)*stride*RS + fq % stride;
    double J[ND][ND];
    #pragma unroll
    for (int i = 0; i < ND; ++i)
      #pragma unroll
      for (int j = 0; j < ND; ++j) {
        double acc = 0.;
        #pragma unroll
        for (int m = 0; m < RS; ++m) acc += ops.bnd[sign][m]*s_jac[(i*ND + j)*nq + base + m*stride];
        J[i][j] = acc;
      }
    double* dst = a.faces + ((size_t)e*2*ND + f)*a.face_width;
    #pragma unroll
    for (int j = 0; j < ND; ++j) dst[j*nfq + fq] = det_unit_column<ND>(J, d, j);
  }
  __syncthreads();

  /* vertex time-step scale: extrapolate normals and determinant to the vertices, innermost dimension first */
  for (int item = t; item < (ND*ND + 1)*n_vert; item += T) {
    const int field = item/n_vert, iv = item % n_vert;
    const double* fld = s_nrm + field*nq;
    double total = 0.;
    if constexpr (ND == 1) {
      for (int i = 0; i < RS; ++i) total += ops.bnd[iv][i]*fld[i];
    } else if constexpr (ND == 2) {
      for (int i = 0; i < RS; ++i) {
        double r1 = 0.;
        for (int j = 0; j < RS; ++j) r1 += ops.bnd[iv & 1][j]*fld[i*RS + j];
        total += ops.bnd[(iv >> 1) & 1][i]*r1;
      }
    } else {
      for (int i = 0; i < RS; ++i) {
        double r2 = 0.;
        for (int j = 0; j < RS; ++j) {
          double r1 = 0.;
          for (int k = 0; k < RS; ++k) r1 += ops.bnd[iv & 1][k]*fld[(i*RS + j)*RS + k];
          r2 += ops.bnd[(iv >> 1) & 1][j]*r1;
        }
        total += ops.bnd[(iv >> 2) & 1][i]*r2;
      }
    }
    s_vx[field*n_vert + iv] = total;
  }
  __syncthreads();
  for (int iv = t; iv < n_vert; iv += T) {
    double norm_sum = 0.;
    #pragma unroll
    for (int i = 0; i < ND; ++i) {
      double norm_sq = 0.;
      #pragma unroll
      for (int j = 0; j < ND; ++j) { const double c = s_vx[(i*ND + j)*n_vert + iv]; norm_sq += c*c; }
      norm_sum += sqrt(norm_sq);
    }
    a.vtss[(size_t)e*n_vert + iv] = nom*s_vx[ND*ND*n_vert + iv]/norm_sum;
  }
}

/* Element::set_jacobian: unit face Jacobian of the Cartesian elements */
__global__ void __launch_bounds__(256)
cartesian_face_jacobian_kernel(double* faces, int n_car, int nd, int nfq, int face_width)
{
  const long long gid = (long long)blockIdx.x*blockDim.x + threadIdx.x;
  const long long per_elem = (long long)2*nd*nd*nfq;
  if (gid >= n_car*per_elem) return;
  const int e = (int)(gid/per_elem), r = (int)(gid % per_elem);
  const int f = r/(nd*nfq), j = (r/nfq) % nd, fq = r % nfq;
  faces[((size_t)e*2*nd + f)*face_width + j*nfq + fq] = (f/2 == j) ? 1. : 0.;
}

int launch_set_jacobian(hexed_b200_ctx* c, const double* d_vert, const double* d_node_adj)
{
  if (!c->have_mesh) return fail(c, HEXED_B200_NO_MESH, "no mesh uploaded");
  invalidate_cfl_cache(c); // the vertex time-step scale changes
  if (c->n_car) {
    const long long total = (long long)c->n_car*2*c->nd*c->nd*c->nfq;
    HB_LAUNCH(cartesian_face_jacobian_kernel, (int)((total + 255)/256), 256, 0, c->stream, c->face_state, c->n_car, c->nd, c->nfq, c->nv*c->nfq);
    ++c->launches;
    HB_CUDA(c, cudaGetLastError());
  }
  if (!c->n_def) return 0;
  MetricArgs a;
  a.vert = d_vert; a.node_adj = d_node_adj; a.nom = c->nom; a.refn = c->refn; a.det = c->det; a.faces = c->face_state; a.vtss = c->vtss;
  a.n_car = c->n_car; a.n_def = c->n_def; a.face_width = c->nv*c->nfq;
  return dispatch(c, [&](auto nd, auto rs) {
    constexpr int ND = decltype(nd)::value, RS = decltype(rs)::value;
    constexpr int nq = ipow(RS, ND), n_vert = ipow(2, ND);
    constexpr size_t smem = sizeof(double)*((ND + ND*ND + ND*ND + 1)*nq + n_vert*ND + (ND*ND + 1)*n_vert);
    auto k = set_jacobian_kernel<ND, RS>;
    if (smem > 48*1024) HB_CUDA(c, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int threads = nq >= 256 ? 256 : ((nq + 31)/32)*32;
    HB_LAUNCH(k, c->n_def, threads, smem, c->stream, a, c->ops);
    ++c->launches;
    HB_CUDA(c, cudaGetLastError());
    return 0;
  });
}

/* ---------------- shared face normals: the connection passes of Solver::calc_jacobian (reference src/Solver.cpp:287-357) ----------------
 * After set_jacobian every element face holds ITS element's normal in the first n_dim*nfq doubles of its face storage. The passes:
 *   1. coarse normals prolonged onto the mortar faces (compute_prolong(mesh, true), :299) -- the existing prolong kernel;
 *   2. fine side of every fine connection := +-(coarse normal), through the face permutation (:301-317);
 *   3. ghost face of every boundary connection := inside face (:319-326);
 *   4. every deformed connection: permute side 1, average the two element normals with the flip signs, give each side its signed copy,
 *      un-permute, store as Kernel_connection::normal() and as the kernel_face_normal() of the deformed elements on either side (:328-355);
 *   5. coarse element face normal := its own normal (:357-369).
 * All factors are 0.5 and +-1, so the result is bit-identical to the reference's whatever the evaluation order.
 * slot_kind[s] = 1 marks mortar faces (fine faces of a refined face). Direction code = i_dim0 + 3*i_dim1 + 9*sign0 + 18*sign1. */
__global__ void __launch_bounds__(128)
mark_mortar_kernel(const int* ref_face, int n_ref, int nd, int* slot_kind)
{
  const int i = blockIdx.x*blockDim.x + threadIdx.x;
  if (i >= n_ref) return;
  const int* rf = ref_face + (size_t)i*8;
  int n_fine = 1 << (nd - 1);
  for (int k = 0; k < nd - 1; ++k) n_fine /= 1 + rf[5 + k];
  for (int k = 0; k < n_fine; ++k) slot_kind[rf[1 + k]] = 1;
}

struct NormalArgs
{
  double* faces; double* normals; const int* def_con; const int* perm; const int* slot_kind; const int* ref_face;
  int n_con, n_ref, nd, nfq, face_width, n_car, n_elem;
};

__device__ __forceinline__ void decode_flips(int code, int& s0, int& s1)
{
  const int sign0 = (code/9) % 2, sign1 = (code/18) % 2;
  s0 = 1 - 2*(sign0 == 0); // 1 - 2*flip_normal(0), flip_normal(i) = (face_sign[i] == i)  (include/Kernel_connection.hpp:22-24)
  s1 = 1 - 2*(sign1 == 1);
}

/* pass: 0 = fine connections, 1 = boundary ghosts, 2 = shared normal */
__global__ void __launch_bounds__(256)
connection_normal_kernel(NormalArgs a, int pass)
{
  const long long gid = (long long)blockIdx.x*blockDim.x + threadIdx.x;
  const int i = (int)(gid/a.nfq), p = (int)(gid % a.nfq);
  if (i >= a.n_con) return;
  const int* row = a.def_con + (size_t)i*4;
  const int slot0 = row[0], slot1 = row[1], code = row[2];
  const int p1 = a.perm[code*a.nfq + p]; // matched[p] = original[p1]
  const int n_elem_face = 2*a.nd*a.n_elem;
  const bool mortar0 = a.slot_kind[slot0] == 1, mortar1 = a.slot_kind[slot1] == 1;
  int s0, s1;
  decode_flips(code, s0, s1);
  double* f0 = a.faces + (size_t)slot0*a.face_width;
  double* f1 = a.faces + (size_t)slot1*a.face_width;
  if (pass == 0) {
    if (mortar0 == mortar1) return;
    // face[0] = the mortar (coarse) side, face[1] = the fine element's side; the reference permutes face[1] with the connection's
    // side-1 permutation whichever side it is (:309-315)
    const double* mort = mortar1 ? f1 : f0;
    double* other = mortar1 ? f0 : f1;
    const int sign = 1 - 2*((s0 < 0) != (s1 < 0));
    for (int j = 0; j < a.nd; ++j) other[j*a.nfq + p1] = sign*mort[j*a.nfq + p];
    return;
  }
  if (pass == 1) {
    if (slot1 < n_elem_face || mortar1 || slot0 >= n_elem_face) return;
    for (int j = 0; j < a.nd; ++j) f1[j*a.nfq + p] = f0[j*a.nfq + p];
    return;
  }
  const int def_first = 2*a.nd*a.n_car;
  for (int j = 0; j < a.nd; ++j) {
    double n = 0;
    n += 0.5*s0*f0[j*a.nfq + p];
    n += 0.5*s1*f1[j*a.nfq + p1];
    const double o0 = s0*n, o1 = s1*n;
    f0[j*a.nfq + p] = o0;
    f1[j*a.nfq + p1] = o1;
    // Kernel_connection::normal() = normal(0) is the connection's own storage in the reference (include/connection.hpp:85-86); a table
    // that points it at the side-1 element's face normal (which holds normal(1)) keeps the element's copy
    const bool side1_owns = slot1 >= def_first && slot1 < n_elem_face && row[3] == slot1 - def_first;
    if (!side1_owns) a.normals[((size_t)row[3]*a.nd + j)*a.nfq + p] = o0;
    if (slot0 >= def_first && slot0 < n_elem_face) a.normals[((size_t)(slot0 - def_first)*a.nd + j)*a.nfq + p] = o0;
    if (slot1 >= def_first && slot1 < n_elem_face) a.normals[((size_t)(slot1 - def_first)*a.nd + j)*a.nfq + p1] = o1;
  }
}

__global__ void __launch_bounds__(256)
coarse_normal_kernel(NormalArgs a)
{
  const long long gid = (long long)blockIdx.x*blockDim.x + threadIdx.x;
  const int i = (int)(gid/a.nfq), p = (int)(gid % a.nfq);
  if (i >= a.n_ref) return;
  const int coarse = a.ref_face[(size_t)i*8];
  const int def_first = 2*a.nd*a.n_car;
  if (coarse < def_first || coarse >= 2*a.nd*a.n_elem) return;
  for (int j = 0; j < a.nd; ++j) a.normals[((size_t)(coarse - def_first)*a.nd + j)*a.nfq + p] = a.faces[(size_t)coarse*a.face_width + j*a.nfq + p];
}

int launch_shared_normals(hexed_b200_ctx* c)
{
  if (!c->have_mesh) return fail(c, HEXED_B200_NO_MESH, "no mesh uploaded");
  invalidate_admis(c); // the face storage is used as scratch for the normals
  DevScratch<int> scratch;
  HB_CUDA(c, scratch.alloc(c->n_face_slot));
  int* slot_kind = scratch.p;
  HB_CUDA(c, cudaMemsetAsync(slot_kind, 0, sizeof(int)*(c->n_face_slot ? c->n_face_slot : 1), c->stream));
  int rc = 0;
  if (c->n_ref) {
    HB_LAUNCH(mark_mortar_kernel, (c->n_ref + 127)/128, 128, 0, c->stream, c->ref_face, c->n_ref, c->nd, slot_kind);
    ++c->launches;
    rc = launch_prolong(c, 0, c->nv, 1);
  }
  NormalArgs a;
  a.faces = c->face_state; a.normals = c->normals; a.def_con = c->def_con; a.perm = c->perm; a.slot_kind = slot_kind; a.ref_face = c->ref_face;
  a.n_con = c->n_def_con; a.n_ref = c->n_ref; a.nd = c->nd; a.nfq = c->nfq; a.face_width = c->nv*c->nfq; a.n_car = c->n_car; a.n_elem = c->n_elem;
  const long long n = (long long)c->n_def_con*c->nfq;
  if (!rc && n) {
    for (int pass = 0; pass < 3; ++pass) {
      HB_LAUNCH(connection_normal_kernel, (int)((n + 255)/256), 256, 0, c->stream, a, pass);
      ++c->launches;
    }
  }
  const long long nr = (long long)c->n_ref*c->nfq;
  if (!rc && nr) { HB_LAUNCH(coarse_normal_kernel, (int)((nr + 255)/256), 256, 0, c->stream, a); ++c->launches; }
  if (!rc) rc = check(c, cudaGetLastError(), "shared normals");
  if (!rc) rc = check(c, cudaStreamSynchronize(c->stream), "shared normals");
  return rc; // the scratch is released after the synchronisation
}

} // namespace hb
