/* api.cu -- C ABI (include/hexed_b200.h): context life cycle, device mirror of the flattened mesh, transfers and the
 * stage drivers that sequence the kernels exactly like the reference's macros
 * (src/kernels_convective.cpp:8-16, src/kernels_max_dt.cpp:8-12). */
#include "common.cuh"
#include <cmath>
#include <cstring>

namespace hb {

static std::string g_create_error;

int fail(hexed_b200_ctx* c, int code, const std::string& msg)
{
  if (c) c->err = msg; else g_create_error = msg;
  return code;
}

int check(hexed_b200_ctx* c, cudaError_t e, const char* what)
{
  if (e == cudaSuccess) return 0;
  return fail(c, HEXED_B200_CUDA_ERROR, std::string(what) + ": " + cudaGetErrorString(e));
}

template <class T>
static int dev_alloc(hexed_b200_ctx* c, T** p, size_t n, bool zero = true)
{
  *p = nullptr;
  if (!n) n = 1;
  HB_CUDA(c, cudaMalloc(p, sizeof(T)*n));
  if (zero) HB_CUDA(c, cudaMemsetAsync(*p, 0, sizeof(T)*n, c->stream));
  return 0;
}

template <class T> static void dev_free(T*& p) { if (p) cudaFree(p); p = nullptr; }

static void free_mesh(hexed_b200_ctx* c)
{
  dev_free(c->state); dev_free(c->tss); dev_free(c->cache); dev_free(c->av); dev_free(c->forcing); dev_free(c->adv);
  dev_free(c->nom); dev_free(c->vtss); dev_free(c->uncert); dev_free(c->refn); dev_free(c->det);
  dev_free(c->face_state); dev_free(c->face_ldg); dev_free(c->face_wide); dev_free(c->normals);
  dev_free(c->car_con); dev_free(c->def_con); dev_free(c->ref_face); dev_free(c->pre_prolong);
  dev_free(c->cfl_approx); invalidate_cfl_cache(c); c->tss_is_one = false; // (also clears the fused admissibility flags)
  dev_free(c->record); dev_free(c->elem_vertex); dev_free(c->matchers); dev_free(c->vertex_vals); dev_free(c->vertex_scratch);
  c->n_vertex = c->n_match = 0;
  c->n_cut_car = c->n_cut_def = c->n_pre_prolong = 0;
  for (auto& l : c->lists) {
    dev_free(l.d_slots); dev_free(l.d_buf); dev_free(l.d_up);
    if (l.h_down) cudaFreeHost(l.h_down);
    if (l.h_up) cudaFreeHost(l.h_up);
    if (l.ev_down) cudaEventDestroy(l.ev_down);
    if (l.ev_up) cudaEventDestroy(l.ev_up);
  }
  c->n_pending_uploads = 0;
  c->lists.clear();
  for (auto& b : c->bcs) { dev_free(b.inside); dev_free(b.ghost); dev_free(b.normal); dev_free(b.params); dev_free(b.cache); }
  c->bcs.clear();
  c->have_mesh = false;
}

/* match_faces index table for one direction: matched[p] = original[table[p]]; transpose, then flip
 * (reference include/Spatial.hpp:85-129, include/Kernel_connection.hpp:22-36) */
static void build_perm(int nd, int rs, int d0, int d1, int s0, int s1, int* table)
{
  const int nfq = ipow(rs, nd - 1);
  const bool flip_n0 = (s0 == 0), flip_n1 = (s1 == 1);
  const bool flip_tang = (d0 != d1) && (flip_n0 == flip_n1);
  const bool transpose = (d0 == 0 && d1 == 2) || (d0 == 2 && d1 == 0);
  for (int p = 0; p < nfq; ++p) table[p] = p;
  if (nd == 3) {
    std::vector<int> cur(table, table + nfq), next(nfq);
    if (transpose) {
      for (int a = 0; a < rs; ++a) for (int b = 0; b < rs; ++b) next[a*rs + b] = cur[b*rs + a];
      cur = next;
    }
    if (flip_tang) {
      // Eigen colwise().reverse() on the column-major map == reversal of the fastest face index
      const bool fast = (d0 > 3 - d0 - d1) != transpose;
      for (int a = 0; a < rs; ++a) for (int b = 0; b < rs; ++b) next[a*rs + b] = fast ? cur[a*rs + (rs - 1 - b)] : cur[(rs - 1 - a)*rs + b];
      cur = next;
    }
    for (int p = 0; p < nfq; ++p) table[p] = cur[p];
  } else if (nd == 2) {
    if (flip_tang) for (int p = 0; p < nfq; ++p) table[p] = nfq - 1 - p;
  }
}

static int slot_array(hexed_b200_ctx* c, int slot, double** base, size_t* elem_stride, bool alloc)
{
  // maps a reference element slot to (device array base of that slot for element 0, stride in doubles)
  const int nv = c->nv, nq = c->nq, rs = c->rs;
  const size_t ne = c->n_elem;
  auto lazy = [&](double** arr, size_t per_elem) -> int {
    if (!*arr) { if (!alloc) { *base = nullptr; return 0; } int rc = dev_alloc(c, arr, ne*per_elem); if (rc) return rc; }
    return 0;
  };
  int rc = 0;
  if (slot < nv) { *base = c->state + (size_t)slot*nq; *elem_stride = (size_t)nv*nq; }
  else if (slot == nv) { *base = c->tss; *elem_stride = nq; }
  else if (slot < nv + 3) { rc = lazy(&c->av, 2*(size_t)nq); *base = c->av ? c->av + (size_t)(slot - nv - 1)*nq : nullptr; *elem_stride = 2*(size_t)nq; }
  else if (slot < nv + 7) { rc = lazy(&c->forcing, 4*(size_t)nq); *base = c->forcing ? c->forcing + (size_t)(slot - nv - 3)*nq : nullptr; *elem_stride = 4*(size_t)nq; }
  else if (slot < nv + 7 + rs) { rc = lazy(&c->adv, (size_t)rs*nq); *base = c->adv ? c->adv + (size_t)(slot - nv - 7)*nq : nullptr; *elem_stride = (size_t)rs*nq; }
  else {
    const int k = slot - (nv + 7 + rs);
    const int cs = nv > rs ? nv : rs; // residual cache: max(n_var, row_size) slots (reference src/Storage_params.cpp:32-35)
    *base = c->cache + (size_t)k*nq; *elem_stride = (size_t)cs*nq;
  }
  return rc;
}

} // namespace hb

using namespace hb;

extern "C" {

int hexed_b200_device_count(int* count)
{
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  *count = (e == cudaSuccess) ? n : 0;
  return 0;
}

const char* hexed_b200_last_error(const hexed_b200_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int hexed_b200_create(hexed_b200_ctx** out, int device, int n_dim, int row_size, const double* basis, int n_basis)
{
  *out = nullptr;
  if (n_dim < 1 || n_dim > 3 || row_size < 2 || row_size > MAX_RS) return fail(nullptr, HEXED_B200_INVALID_KERNEL, "demand for invalid kernel");
  const int rs = row_size;
  const int expect = 2*rs + 3*rs*rs + 2*rs + 4*rs*rs + 3 + rs;
  if (n_basis != expect) return fail(nullptr, HEXED_B200_BAD_ARGUMENT, "basis table has the wrong length");
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev <= device || device < 0)
    return fail(nullptr, HEXED_B200_NO_DEVICE, "no CUDA device available: hexed_b200 has no CPU fallback");
  hexed_b200_ctx* c = new hexed_b200_ctx();
  c->device = device;
  if (cudaSetDevice(device) != cudaSuccess) { delete c; return fail(nullptr, HEXED_B200_NO_DEVICE, "cudaSetDevice failed"); }
  c->nd = n_dim; c->rs = rs; c->nq = ipow(rs, n_dim); c->nfq = c->nq/rs; c->nv = n_dim + 2; c->n_vert = ipow(2, n_dim);
  const double* p = basis;
  double diff[MAX_RS][MAX_RS] = {}, bnd[2][MAX_RS] = {};
  std::memset(&c->ops, 0, sizeof(c->ops)); std::memset(&c->filt, 0, sizeof(c->filt)); std::memset(&c->transfer, 0, sizeof(c->transfer));
  for (int i = 0; i < rs; ++i) c->ops.node[i] = *p++;
  for (int i = 0; i < rs; ++i) c->weight[i] = *p++;
  for (int i = 0; i < rs; ++i) for (int j = 0; j < rs; ++j) diff[i][j] = *p++;
  for (int s = 0; s < 2; ++s) for (int j = 0; j < rs; ++j) bnd[s][j] = *p++;
  for (int i = 0; i < rs; ++i) for (int j = 0; j < rs; ++j) c->orthogonal[i][j] = *p++;
  for (int i = 0; i < rs; ++i) for (int j = 0; j < rs; ++j) c->filt.filter[i][j] = *p++;
  for (int h = 0; h < 2; ++h) for (int i = 0; i < rs; ++i) for (int j = 0; j < rs; ++j) c->transfer.prolong[h][i][j] = *p++;
  for (int h = 0; h < 2; ++h) for (int i = 0; i < rs; ++i) for (int j = 0; j < rs; ++j) c->transfer.restrict_[h][i][j] = *p++;
  c->min_eig_conv = *p++; c->min_eig_diff = *p++; c->quad_safety = *p++;
  for (int i = 0; i < rs; ++i) c->gl_node[i] = *p++;
  for (int i = 0; i < rs; ++i) {
    // lift = diag(1/w) boundary^T diag(-1, +1)   (reference include/Derivative.hpp:20-29)
    const double inv_w = 1./c->weight[i];
    c->ops.lift[i][0] = inv_w*bnd[0][i]*-1.;
    c->ops.lift[i][1] = inv_w*bnd[1][i]*1.;
    for (int j = 0; j < rs; ++j) {
      c->ops.diff[i][j] = diff[i][j];
      c->ops.bnd[0][j] = bnd[0][j]; c->ops.bnd[1][j] = bnd[1][j];
    }
  }
  for (int i = 0; i < rs; ++i) for (int j = 0; j < rs; ++j)
    c->ops.dfull[i][j] = diff[i][j] - (c->ops.lift[i][0]*bnd[0][j] + c->ops.lift[i][1]*bnd[1][j]);
  { // even-odd halves (common.cuh); only meaningful for node sets symmetric about 1/2, which is checked here entry by entry
    const int h = rs/2;
    double scale = 0, asym = 0;
    for (int i = 0; i < rs; ++i) {
      for (int j = 0; j < rs; ++j) {
        scale = std::max(scale, std::abs(c->ops.dfull[i][j]));
        asym = std::max(asym, std::abs(c->ops.dfull[i][j] + c->ops.dfull[rs - 1 - i][rs - 1 - j]));
        asym = std::max(asym, std::abs(c->ops.diff[i][j] + c->ops.diff[rs - 1 - i][rs - 1 - j]));
      }
      asym = std::max(asym, std::abs(c->ops.lift[i][0] + c->ops.lift[rs - 1 - i][1]));
      asym = std::max(asym, std::abs(bnd[0][i] - bnd[1][rs - 1 - i]));
    }
    c->ops_symmetric = rs % 2 == 0 && asym <= 1e-13*scale;
    for (int i = 0; i < h; ++i) {
      for (int k = 0; k < h; ++k) {
        c->ops.eo_a[i][k] = .5*(c->ops.dfull[i][k] - c->ops.dfull[i][rs - 1 - k]);
        c->ops.eo_s[i][k] = .5*(c->ops.dfull[i][k] + c->ops.dfull[i][rs - 1 - k]);
        c->ops.eo_da[i][k] = .5*(diff[i][k] - diff[i][rs - 1 - k]);
        c->ops.eo_ds[i][k] = .5*(diff[i][k] + diff[i][rs - 1 - k]);
      }
      c->ops.eo_la[i] = .5*(c->ops.lift[i][0] - c->ops.lift[i][1]);
      c->ops.eo_ls[i] = .5*(c->ops.lift[i][0] + c->ops.lift[i][1]);
      c->ops.eo_ba[i] = .5*(bnd[0][i] - bnd[0][rs - 1 - i]);
      c->ops.eo_bs[i] = .5*(bnd[0][i] + bnd[0][rs - 1 - i]);
    }
  }
  static const char* names[ST_COUNT] = {"neighbor", "neighbor", "local", "local", "compute time step", "compute time step",
                                        "prolong/restrict", "boundary conditions", "write face", "reconcile LDG flux", "reconcile LDG flux", "check admis."};
  static const int trees[ST_COUNT] = {0, 1, 0, 1, 0, 1, 2, 2, 2, 0, 1, 2};
  for (int i = 0; i < ST_COUNT; ++i) { c->stats[i].name = names[i]; c->stats[i].deformed = trees[i]; }
  int rc = check(c, cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking), "cudaStreamCreate");
  if (!rc) rc = check(c, cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking), "cudaStreamCreate");
  if (!rc) rc = check(c, cudaEventCreateWithFlags(&c->ev_copy, cudaEventDisableTiming), "cudaEventCreate");
  if (!rc) rc = check(c, cudaEventCreate(&c->ev0), "cudaEventCreate");
  if (!rc) rc = check(c, cudaEventCreate(&c->ev1), "cudaEventCreate");
  if (!rc) rc = dev_alloc(c, &c->d_scalar, 8);
  if (!rc) rc = dev_alloc(c, &c->d_step, 3);
  if (!rc) rc = check(c, cudaMallocHost(&c->h_scalar, sizeof(double)*8), "cudaMallocHost");
  // permutation tables for all 36 direction codes
  c->h_perm.assign((size_t)36*c->nfq, 0);
  for (int d0 = 0; d0 < 3; ++d0) for (int d1 = 0; d1 < 3; ++d1) for (int s0 = 0; s0 < 2; ++s0) for (int s1 = 0; s1 < 2; ++s1) {
    int* t = c->h_perm.data() + (size_t)dir_code(d0, d1, s0, s1)*c->nfq;
    if (d0 < n_dim && d1 < n_dim) build_perm(n_dim, rs, d0, d1, s0, s1, t);
    else for (int q = 0; q < c->nfq; ++q) t[q] = q;
  }
  if (!rc) rc = dev_alloc(c, &c->perm, c->h_perm.size(), false);
  if (!rc) rc = check(c, cudaMemcpyAsync(c->perm, c->h_perm.data(), sizeof(int)*c->h_perm.size(), cudaMemcpyHostToDevice, c->stream), "upload perm");
  if (!rc) rc = dev_alloc(c, &c->d_face_scratch, 2*(size_t)(c->nd + c->rs)*c->nfq);
  if (!rc) rc = check(c, cudaStreamSynchronize(c->stream), "create sync");
  if (rc) { g_create_error = c->err; hexed_b200_destroy(c); return rc; }
  *out = c;
  return 0;
}

int hexed_b200_destroy(hexed_b200_ctx* c)
{
  if (!c) return 0;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  if (c->copy_stream) cudaStreamSynchronize(c->copy_stream);
  free_mesh(c);
  dev_free(c->perm); dev_free(c->d_scalar); dev_free(c->d_step); dev_free(c->d_face_scratch);
  if (c->h_scalar) cudaFreeHost(c->h_scalar);
  if (c->h_flags) cudaFreeHost(c->h_flags);
  if (c->d_flags) cudaFree(c->d_flags);
  if (c->ev0) cudaEventDestroy(c->ev0);
  if (c->ev1) cudaEventDestroy(c->ev1);
  if (c->ev_copy) cudaEventDestroy(c->ev_copy);
  if (c->copy_stream) { cudaStreamSynchronize(c->copy_stream); cudaStreamDestroy(c->copy_stream); }
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
  return 0;
}

int hexed_b200_synchronize(hexed_b200_ctx* c) { HB_ENTER(c); HB_CUDA(c, cudaStreamSynchronize(c->stream)); return 0; }
int hexed_b200_cuda_stream(hexed_b200_ctx* c, void** stream) { *stream = (void*)c->stream; return 0; }

int hexed_b200_mesh_create(hexed_b200_ctx* c, const hexed_b200_mesh_desc* d)
{
  HB_ENTER(c);
  HB_CUDA(c, cudaSetDevice(c->device));
  HB_CUDA(c, cudaStreamSynchronize(c->stream));
  free_mesh(c);
  if (d->n_car < 0 || d->n_def < 0 || d->n_car_con < 0 || d->n_def_con < 0 || d->n_ref < 0) return fail(c, HEXED_B200_BAD_ARGUMENT, "negative count");
  const size_t ne = (size_t)d->n_car + d->n_def;
  const int nf = 2*c->nd;
  if ((size_t)d->n_face_slot < nf*ne) return fail(c, HEXED_B200_BAD_ARGUMENT, "n_face_slot smaller than 2*n_dim*n_elem");
  if ((size_t)d->n_normal_slot < (size_t)nf*d->n_def) return fail(c, HEXED_B200_BAD_ARGUMENT, "n_normal_slot smaller than 2*n_dim*n_def");
  c->n_car = d->n_car; c->n_def = d->n_def; c->n_elem = (int)ne;
  c->n_face_slot = d->n_face_slot; c->n_normal_slot = d->n_normal_slot;
  c->n_car_con = d->n_car_con; c->n_def_con = d->n_def_con; c->n_ref = d->n_ref;
  const size_t nq = c->nq, nfq = c->nfq, nv = c->nv;
  int rc = 0;
  if (!rc) rc = dev_alloc(c, &c->state, ne*nv*nq);
  if (!rc) rc = dev_alloc(c, &c->tss, ne*nq);
  if (!rc) rc = dev_alloc(c, &c->cache, ne*(nv > (size_t)c->rs ? nv : (size_t)c->rs)*nq);
  if (!rc) rc = dev_alloc(c, &c->nom, ne);
  if (!rc) rc = dev_alloc(c, &c->vtss, ne*c->n_vert);
  if (!rc) rc = dev_alloc(c, &c->uncert, ne);
  if (!rc) rc = dev_alloc(c, &c->refn, (size_t)d->n_def*c->nd*c->nd*nq);
  if (!rc) rc = dev_alloc(c, &c->det, (size_t)d->n_def*nq);
  if (!rc) rc = dev_alloc(c, &c->face_state, (size_t)d->n_face_slot*nv*nfq);
  if (!rc) rc = dev_alloc(c, &c->normals, (size_t)d->n_normal_slot*c->nd*nfq);
  if (rc) { free_mesh(c); return rc; }
  // validate + repack the integer tables
  std::vector<int> car((size_t)d->n_car_con*4), def((size_t)d->n_def_con*4), ref((size_t)d->n_ref*8);
  auto slot_ok = [&](int s) { return s >= 0 && s < d->n_face_slot; };
  for (int i = 0; i < d->n_car_con; ++i) {
    const int* t = d->car_con + (size_t)i*3;
    if (!slot_ok(t[0]) || !slot_ok(t[1]) || t[2] < 0 || t[2] >= c->nd) { free_mesh(c); return fail(c, HEXED_B200_BAD_ARGUMENT, "bad cartesian connection table entry"); }
    car[(size_t)i*4] = t[0]; car[(size_t)i*4 + 1] = t[1]; car[(size_t)i*4 + 2] = t[2]; car[(size_t)i*4 + 3] = 0;
  }
  for (int i = 0; i < d->n_def_con; ++i) {
    const int* t = d->def_con + (size_t)i*7;
    const bool ok = slot_ok(t[0]) && slot_ok(t[1]) && t[2] >= 0 && t[2] < c->nd && t[3] >= 0 && t[3] < c->nd
                    && (t[4] == 0 || t[4] == 1) && (t[5] == 0 || t[5] == 1) && t[6] >= 0 && t[6] < d->n_normal_slot;
    if (!ok) { free_mesh(c); return fail(c, HEXED_B200_BAD_ARGUMENT, "bad deformed connection table entry"); }
    def[(size_t)i*4] = t[0]; def[(size_t)i*4 + 1] = t[1]; def[(size_t)i*4 + 2] = dir_code(t[2], t[3], t[4], t[5]); def[(size_t)i*4 + 3] = t[6];
  }
  for (int i = 0; i < d->n_ref; ++i) {
    const int* t = d->ref_face + (size_t)i*7;
    int nfine = ipow(2, c->nd - 1);
    for (int k = 0; k < c->nd - 1; ++k) nfine /= 1 + (t[5 + k] != 0);
    bool ok = slot_ok(t[0]);
    for (int k = 0; k < nfine; ++k) ok = ok && slot_ok(t[1 + k]);
    if (!ok) { free_mesh(c); return fail(c, HEXED_B200_BAD_ARGUMENT, "bad refined face table entry"); }
    for (int k = 0; k < 5; ++k) ref[(size_t)i*8 + k] = t[k];
    ref[(size_t)i*8 + 5] = t[5] != 0; ref[(size_t)i*8 + 6] = t[6] != 0; ref[(size_t)i*8 + 7] = 0;
  }
  if (!rc) rc = dev_alloc(c, &c->car_con, car.size(), false);
  if (!rc) rc = dev_alloc(c, &c->def_con, def.size(), false);
  if (!rc) rc = dev_alloc(c, &c->ref_face, ref.size(), false);
  if (!rc && !car.empty()) rc = check(c, cudaMemcpyAsync(c->car_con, car.data(), sizeof(int)*car.size(), cudaMemcpyHostToDevice, c->stream), "upload car_con");
  if (!rc && !def.empty()) rc = check(c, cudaMemcpyAsync(c->def_con, def.data(), sizeof(int)*def.size(), cudaMemcpyHostToDevice, c->stream), "upload def_con");
  if (!rc && !ref.empty()) rc = check(c, cudaMemcpyAsync(c->ref_face, ref.data(), sizeof(int)*ref.size(), cudaMemcpyHostToDevice, c->stream), "upload ref_face");
  if (!rc) rc = check(c, cudaStreamSynchronize(c->stream), "mesh_create sync");
  if (rc) { free_mesh(c); return rc; }
  c->have_mesh = true;
  return 0;
}

static int array_info(hexed_b200_ctx* c, int which, double*** arr, size_t* item, size_t* count, bool alloc)
{
  const size_t nq = c->nq, nfq = c->nfq;
  switch (which) {
    case HEXED_B200_NOMINAL_SIZE: *arr = &c->nom; *item = 1; *count = c->n_elem; break;
    case HEXED_B200_VERTEX_TSS: *arr = &c->vtss; *item = c->n_vert; *count = c->n_elem; break;
    case HEXED_B200_REF_NORMALS: *arr = &c->refn; *item = (size_t)c->nd*c->nd*nq; *count = c->n_def; break;
    case HEXED_B200_JAC_DET: *arr = &c->det; *item = nq; *count = c->n_def; break;
    case HEXED_B200_FACE_STATE: *arr = &c->face_state; *item = c->nv*nfq; *count = c->n_face_slot; break;
    case HEXED_B200_FACE_LDG: *arr = &c->face_ldg; *item = c->nv*nfq; *count = c->n_face_slot; break;
    case HEXED_B200_FACE_WIDE: *arr = &c->face_wide; *item = (size_t)(c->nd + c->rs)*nfq; *count = c->n_face_slot; break;
    case HEXED_B200_NORMALS: *arr = &c->normals; *item = c->nd*nfq; *count = c->n_normal_slot; break;
    case HEXED_B200_UNCERT: *arr = &c->uncert; *item = 1; *count = c->n_elem; break;
    case HEXED_B200_VERTEX_SCRATCH: *arr = &c->vertex_scratch; *item = c->n_vert; *count = c->n_elem; break;
    default: return fail(c, HEXED_B200_BAD_ARGUMENT, "unknown array id");
  }
  if (!**arr && alloc) { int rc = dev_alloc(c, *arr, (*item)*(*count)); if (rc) return rc; }
  return 0;
}

int hexed_b200_upload(hexed_b200_ctx* c, int which, const double* src, size_t first, size_t n)
{
  HB_ENTER(c);
  if (!c->have_mesh) return fail(c, HEXED_B200_NO_MESH, "no mesh uploaded");
  double** arr; size_t item, count;
  int rc = array_info(c, which, &arr, &item, &count, true); if (rc) return rc;
  if (first + n > count) return fail(c, HEXED_B200_BAD_ARGUMENT, "upload range out of bounds");
  if (!n) return 0;
  if (which == HEXED_B200_VERTEX_TSS) invalidate_cfl_cache(c);
  if (which == HEXED_B200_FACE_STATE) invalidate_admis(c);
  HB_CUDA(c, cudaMemcpyAsync(*arr + first*item, src, sizeof(double)*n*item, cudaMemcpyDefault, c->stream));
  HB_CUDA(c, cudaStreamSynchronize(c->stream));
  return 0;
}

int hexed_b200_download(hexed_b200_ctx* c, int which, double* dst, size_t first, size_t n)
{
  HB_ENTER(c);
  if (!c->have_mesh) return fail(c, HEXED_B200_NO_MESH, "no mesh uploaded");
  double** arr; size_t item, count;
  int rc = array_info(c, which, &arr, &item, &count, true); if (rc) return rc;
  if (first + n > count) return fail(c, HEXED_B200_BAD_ARGUMENT, "download range out of bounds");
  if (!n) return 0;
  HB_CUDA(c, cudaMemcpyAsync(dst, *arr + first*item, sizeof(double)*n*item, cudaMemcpyDefault, c->stream));
  HB_CUDA(c, cudaStreamSynchronize(c->stream));
  return 0;
}

static bool is_device_ptr(const void* p)
{
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

static int move_slots(hexed_b200_ctx* c, double* host, size_t elem_stride, int first_slot, int n_slots, int first_elem, int n_elem, bool up)
{
  if (!c->have_mesh) return fail(c, HEXED_B200_NO_MESH, "no mesh uploaded");
  const int total_slots = c->nv + 7 + c->rs + (c->nv > c->rs ? c->nv : c->rs);
  if (first_slot < 0 || n_slots < 0 || first_slot + n_slots > total_slots) return fail(c, HEXED_B200_BAD_ARGUMENT, "slot range out of bounds");
  if (first_elem < 0 || n_elem < 0 || first_elem + n_elem > c->n_elem) return fail(c, HEXED_B200_BAD_ARGUMENT, "element range out of bounds");
  if (!n_elem) return 0;
  if (up) { // the caller may be rewriting the flow state or the time-step scale
    if (first_slot < c->nv) invalidate_cfl_cache(c);
    if (first_slot <= c->nv && first_slot + n_slots > c->nv) c->tss_is_one = false;
  }
  for (int s = 0; s < n_slots; ++s) {
    double* base; size_t dstride;
    int rc = slot_array(c, first_slot + s, &base, &dstride, up); if (rc) return rc;
    double* h = host + (size_t)s*c->nq;
    if (!base) { // array never allocated (or slot not mirrored): reads as zero, writes are dropped
      if (!up) {
        if (is_device_ptr(h)) HB_CUDA(c, cudaMemset2DAsync(h, sizeof(double)*elem_stride, 0, sizeof(double)*c->nq, n_elem, c->stream));
        else for (int e = 0; e < n_elem; ++e) std::memset(h + (size_t)e*elem_stride, 0, sizeof(double)*c->nq);
      }
      continue;
    }
    double* dptr = base + (size_t)first_elem*dstride;
    if (up) HB_CUDA(c, cudaMemcpy2DAsync(dptr, sizeof(double)*dstride, h, sizeof(double)*elem_stride, sizeof(double)*c->nq, n_elem, cudaMemcpyDefault, c->stream));
    else HB_CUDA(c, cudaMemcpy2DAsync(h, sizeof(double)*elem_stride, dptr, sizeof(double)*dstride, sizeof(double)*c->nq, n_elem, cudaMemcpyDefault, c->stream));
  }
  HB_CUDA(c, cudaStreamSynchronize(c->stream));
  return 0;
}

int hexed_b200_upload_elem_slots(hexed_b200_ctx* c, const double* src, size_t elem_stride, int first_slot, int n_slots, int first_elem, int n_elem)
{ HB_ENTER(c); return move_slots(c, const_cast<double*>(src), elem_stride, first_slot, n_slots, first_elem, n_elem, true); }

int hexed_b200_download_elem_slots(hexed_b200_ctx* c, double* dst, size_t elem_stride, int first_slot, int n_slots, int first_elem, int n_elem)
{ HB_ENTER(c); return move_slots(c, dst, elem_stride, first_slot, n_slots, first_elem, n_elem, false); }

int hexed_b200_face_list_create(hexed_b200_ctx* c, const int* slots, int n, int* list_id)
{
  HB_ENTER(c);
  if (!c->have_mesh) return fail(c, HEXED_B200_NO_MESH, "no mesh uploaded");
  for (int i = 0; i < n; ++i) if (slots[i] < 0 || slots[i] >= c->n_face_slot) return fail(c, HEXED_B200_BAD_ARGUMENT, "face slot out of range");
  FaceList l; l.n = n;
  int rc = dev_alloc(c, &l.d_slots, n, false); if (rc) return rc;
  l.buf_doubles = (size_t)n*(c->nd + c->rs)*c->nfq;
  rc = dev_alloc(c, &l.d_buf, l.buf_doubles, false); if (rc) return rc;
  if (n) HB_CUDA(c, cudaMemcpyAsync(l.d_slots, slots, sizeof(int)*n, cudaMemcpyHostToDevice, c->stream));
  HB_CUDA(c, cudaStreamSynchronize(c->stream));
  c->lists.push_back(l);
  *list_id = (int)c->lists.size() - 1;
  return 0;
}

static int list_face_array(hexed_b200_ctx* c, int kind, double** arr, int* width)
{
  double** a; size_t item, count;
  const int which = kind == 0 ? HEXED_B200_FACE_STATE : kind == 1 ? HEXED_B200_FACE_LDG : kind == 2 ? HEXED_B200_FACE_WIDE : -1;
  int rc = array_info(c, which, &a, &item, &count, true); if (rc) return rc;
  *arr = *a; *width = (int)item;
  return 0;
}

int hexed_b200_face_list_download(hexed_b200_ctx* c, int list_id, int kind, double* dst)
{
  HB_ENTER(c);
  if (list_id < 0 || list_id >= (int)c->lists.size()) return fail(c, HEXED_B200_BAD_ARGUMENT, "unknown face list");
  FaceList& l = c->lists[list_id];
  double* arr; int width;
  int rc = list_face_array(c, kind, &arr, &width); if (rc) return rc;
  if (l.down_kind >= 0) HB_CUDA(c, cudaStreamWaitEvent(c->stream, l.ev_down, 0)); // a prefetch is still reading d_buf
  rc = launch_gather_faces(c, arr, width, l.d_slots, l.n, l.d_buf); if (rc) return rc;
  if (l.n) HB_CUDA(c, cudaMemcpyAsync(dst, l.d_buf, sizeof(double)*l.n*width, cudaMemcpyDefault, c->stream));
  HB_CUDA(c, cudaStreamSynchronize(c->stream));
  return 0;
}

int hexed_b200_face_list_upload(hexed_b200_ctx* c, int list_id, int kind, const double* src)
{
  HB_ENTER(c);
  if (list_id < 0 || list_id >= (int)c->lists.size()) return fail(c, HEXED_B200_BAD_ARGUMENT, "unknown face list");
  FaceList& l = c->lists[list_id];
  double* arr; int width;
  int rc = list_face_array(c, kind, &arr, &width); if (rc) return rc;
  if (l.n) HB_CUDA(c, cudaMemcpyAsync(l.d_buf, src, sizeof(double)*l.n*width, cudaMemcpyDefault, c->stream));
  return launch_scatter_faces(c, arr, width, l.d_slots, l.n, l.d_buf);
}

/* ---- asynchronous host traffic for HOST-applied boundary conditions (Solver::apply_state_bcs, src/Solver.cpp:56-67, when the
 * conditions are not registered on the device): downloads are started as soon as the faces exist and collected later; uploads are
 * started at once and land in the face storage inside the next stage driver, after its interior Neighbor kernels. ---- */
static int list_async_buffers(hexed_b200_ctx* c, FaceList& l)
{
  if (l.h_down) return 0;
  const size_t bytes = sizeof(double)*(l.buf_doubles ? l.buf_doubles : 1);
  HB_CUDA(c, cudaMallocHost(&l.h_down, bytes));
  HB_CUDA(c, cudaMallocHost(&l.h_up, bytes));
  int rc = dev_alloc(c, &l.d_up, l.buf_doubles, false); if (rc) return rc;
  HB_CUDA(c, cudaEventCreateWithFlags(&l.ev_down, cudaEventDisableTiming));
  HB_CUDA(c, cudaEventCreateWithFlags(&l.ev_up, cudaEventDisableTiming));
  return 0;
}

int hexed_b200_face_list_prefetch(hexed_b200_ctx* c, int list_id, int kind)
{
  HB_ENTER(c);
  if (list_id < 0 || list_id >= (int)c->lists.size()) return fail(c, HEXED_B200_BAD_ARGUMENT, "unknown face list");
  FaceList& l = c->lists[list_id];
  int rc = list_async_buffers(c, l); if (rc) return rc;
  double* arr; int width;
  rc = list_face_array(c, kind, &arr, &width); if (rc) return rc;
  if (l.down_kind >= 0) HB_CUDA(c, cudaStreamWaitEvent(c->stream, l.ev_down, 0)); // an uncollected earlier prefetch still reads d_buf
  rc = launch_gather_faces(c, arr, width, l.d_slots, l.n, l.d_buf); if (rc) return rc;
  HB_CUDA(c, cudaEventRecord(c->ev_copy, c->stream));
  HB_CUDA(c, cudaStreamWaitEvent(c->copy_stream, c->ev_copy, 0));
  if (l.n) HB_CUDA(c, cudaMemcpyAsync(l.h_down, l.d_buf, sizeof(double)*l.n*width, cudaMemcpyDeviceToHost, c->copy_stream));
  HB_CUDA(c, cudaEventRecord(l.ev_down, c->copy_stream));
  l.down_kind = kind;
  return 0;
}

int hexed_b200_face_list_prefetched(hexed_b200_ctx* c, int list_id, int kind, const double** host)
{
  HB_ENTER(c);
  if (list_id < 0 || list_id >= (int)c->lists.size()) return fail(c, HEXED_B200_BAD_ARGUMENT, "unknown face list");
  FaceList& l = c->lists[list_id];
  if (l.down_kind != kind) { // nothing in flight for this kind: fetch now
    int rc = hexed_b200_face_list_prefetch(c, list_id, kind); if (rc) return rc;
  }
  HB_CUDA(c, cudaEventSynchronize(l.ev_down));
  l.down_kind = -1;
  *host = l.h_down;
  return 0;
}

int hexed_b200_face_list_staging(hexed_b200_ctx* c, int list_id, double** host)
{
  HB_ENTER_KEEP(c);
  if (list_id < 0 || list_id >= (int)c->lists.size()) return fail(c, HEXED_B200_BAD_ARGUMENT, "unknown face list");
  FaceList& l = c->lists[list_id];
  int rc = list_async_buffers(c, l); if (rc) return rc;
  if (l.up_kind >= 0) HB_CUDA(c, cudaEventSynchronize(l.ev_up)); // the previous contents are still being read by the copy engine
  *host = l.h_up;
  return 0;
}

int hexed_b200_face_list_upload_deferred(hexed_b200_ctx* c, int list_id, int kind)
{
  HB_ENTER_KEEP(c);
  if (list_id < 0 || list_id >= (int)c->lists.size()) return fail(c, HEXED_B200_BAD_ARGUMENT, "unknown face list");
  FaceList& l = c->lists[list_id];
  int rc = list_async_buffers(c, l); if (rc) return rc;
  double* arr; int width;
  rc = list_face_array(c, kind, &arr, &width); if (rc) return rc;
  if (l.up_kind >= 0) { rc = flush_pending_uploads(c); if (rc) return rc; }
  // the scatter of the previous upload from d_up has been enqueued on `stream`: the copy engine must not overwrite d_up before it ran
  HB_CUDA(c, cudaEventRecord(c->ev_copy, c->stream));
  HB_CUDA(c, cudaStreamWaitEvent(c->copy_stream, c->ev_copy, 0));
  if (l.n) HB_CUDA(c, cudaMemcpyAsync(l.d_up, l.h_up, sizeof(double)*l.n*width, cudaMemcpyHostToDevice, c->copy_stream));
  HB_CUDA(c, cudaEventRecord(l.ev_up, c->copy_stream));
  l.up_kind = kind;
  ++c->n_pending_uploads;
  return 0;
}

extern "C++" {
namespace hb {
int flush_pending_uploads(hexed_b200_ctx* c)
{
  for (FaceList& l : c->lists) if (l.up_kind >= 0) {
    double* arr; int width;
    int rc = list_face_array(c, l.up_kind, &arr, &width); if (rc) return rc;
    HB_CUDA(c, cudaStreamWaitEvent(c->stream, l.ev_up, 0));
    rc = launch_scatter_faces(c, arr, width, l.d_slots, l.n, l.d_up); if (rc) return rc;
    l.up_kind = -1;
  }
  c->n_pending_uploads = 0;
  return 0;
}
}
}

int hexed_b200_face_permutation_table(hexed_b200_ctx* c, const int dir[4], int* out)
{
  HB_ENTER(c);
  if (dir[0] < 0 || dir[0] >= c->nd || dir[1] < 0 || dir[1] >= c->nd) return fail(c, HEXED_B200_BAD_ARGUMENT, "bad direction");
  const int* t = c->h_perm.data() + (size_t)dir_code(dir[0], dir[1], dir[2] != 0, dir[3] != 0)*c->nfq;
  for (int q = 0; q < c->nfq; ++q) out[q] = t[q];
  return 0;
}

int hexed_b200_face_permutation_indices(int n_dim, int row_size, const int dir[4], int* out)
{
  if (n_dim < 1 || n_dim > 3 || row_size < 2 || row_size > MAX_RS) return HEXED_B200_INVALID_KERNEL;
  if (dir[0] < 0 || dir[0] >= n_dim || dir[1] < 0 || dir[1] >= n_dim) return HEXED_B200_BAD_ARGUMENT;
  build_perm(n_dim, row_size, dir[0], dir[1], dir[2] != 0, dir[3] != 0, out);
  return 0;
}

int hexed_b200_face_permutation(hexed_b200_ctx* c, const int dir[4], int restore, double* data)
{
  HB_ENTER(c);
  if (dir[0] < 0 || dir[0] >= c->nd || dir[1] < 0 || dir[1] >= c->nd) return fail(c, HEXED_B200_BAD_ARGUMENT, "bad direction");
  const size_t n = (size_t)c->nv*c->nfq;
  HB_CUDA(c, cudaMemcpyAsync(c->d_face_scratch, data, sizeof(double)*n, cudaMemcpyDefault, c->stream));
  int rc = launch_permute_face(c, c->d_face_scratch, c->nv, dir_code(dir[0], dir[1], dir[2] != 0, dir[3] != 0), restore); if (rc) return rc;
  HB_CUDA(c, cudaMemcpyAsync(data, c->d_face_scratch, sizeof(double)*n, cudaMemcpyDefault, c->stream));
  HB_CUDA(c, cudaStreamSynchronize(c->stream));
  return 0;
}

/* ---- stage drivers ---- */

int hexed_b200_compute_euler(hexed_b200_ctx* c, hexed_b200_options o)
{
  HB_ENTER_KEEP(c);
  if (c->n_pending_uploads) { // ghost faces still crossing PCIe: the connections declared late (hexed_b200_set_partition) wait for them, the rest runs now
    int rc = hexed_b200_compute_euler_begin(c);
    if (!rc) rc = hexed_b200_compute_euler_finish(c, o);
    return rc;
  }
  // reference src/kernels_convective.cpp:8-16: Neighbor(car) Neighbor(def) Restrict Local(car) Local(def) Prolong
  int rc;
  if ((rc = launch_neighbor_euler(c, 0))) return rc;
  if ((rc = launch_neighbor_euler(c, 1))) return rc;
  if ((rc = launch_restrict(c, 0, c->nv, 1))) return rc;
  if ((rc = launch_local_euler(c, 0, o))) return rc;
  if ((rc = launch_local_euler(c, 1, o))) return rc;
  if ((rc = launch_prolong(c, 0, c->nv, 0))) return rc;
  return 0;
}

/* ---- the time loop of Solver::update for the inviscid case with device boundary conditions, without a host round trip per step ----
 * One step = max_dt + 2 x (ghost fill + compute_euler) (reference src/Solver.cpp:834-886 with n_cheby_flow = 1). The reference API
 * passes dt through the host (max_dt returns it, Kernel_options carries it), i.e. one device synchronisation per step; here the
 * reduction leaves dt on the device, the Local kernels read it from there, and the step is captured ONCE in a CUDA graph and replayed.
 * On the large meshes of the headline the step is 30 ms and none of this matters; on the meshes of the 2-D sample cases (vortex:
 * 256 elements, ~10 launches of a few microseconds each) launch latency and the synchronisation ARE the step. */
static double chebyshev_step(int n_steps, int i_step) { return 1/(1 - std::cos((n_steps - i_step - 0.5)*M_PI/n_steps)); } // src/math.cpp:104-107

static int enqueue_euler_step(hexed_b200_ctx* c, double safety_conv, double cheby_factor)
{
  int rc = launch_max_dt_euler_device(c, safety_conv, c->d_step);
  if (!rc) rc = launch_scale_dt(c, c->d_step, cheby_factor); // dt = nominal_dt*chebyshev_step(n_cheby, i_cheby), :850
  for (int stage = 0; stage < 2 && !rc; ++stage) {
    hexed_b200_options o; o.dt = 1.; o.i_stage = stage; o.compute_residual = 0; o.use_filter = 0;
    rc = launch_bcs(c);
    if (!rc) rc = hexed_b200_compute_euler(c, o);
  }
  if (!rc) rc = launch_accumulate_time(c, c->d_step);
  return rc;
}

/* viscous step (Solver::update with use_ldg(), src/Solver.cpp:857-865): max_dt_navier_stokes + ghost fill + compute_navier_stokes (stage 0,
 * flux boundary conditions on the device) + ghost fill + compute_euler (stage 1) */
struct ViscousStep { double safety_diff; hexed_b200_transport visc, cond; };
static const GenericOps* generic_ops(int pde);
static PdeParams make_params(hexed_b200_ctx* c, int pde, hexed_b200_transport visc, hexed_b200_transport cond, double p0, double p1);
static int diffusion_stage(hexed_b200_ctx* c, int pde, hexed_b200_options o, const PdeParams& pp, hexed_b200_callback flux_bc, void* user);

static void device_flux_bcs(void* user) { launch_flux_bcs(static_cast<hexed_b200_ctx*>(user)); }

static int enqueue_viscous_step(hexed_b200_ctx* c, double safety_conv, const ViscousStep& v, double cheby_factor)
{
  const PdeParams pp = make_params(c, 1, v.visc, v.cond, 0., 0.);
  double unused = 0;
  c->max_dt_device_out = c->d_step;
  int rc = generic_ops(1)->max_dt(c, pp, safety_conv, v.safety_diff, 0, &unused);
  c->max_dt_device_out = nullptr;
  if (!rc) rc = launch_scale_dt(c, c->d_step, cheby_factor);
  hexed_b200_options o; o.dt = 1.; o.i_stage = 0; o.compute_residual = 0; o.use_filter = 0;
  if (!rc) rc = launch_bcs(c);
  if (!rc) rc = diffusion_stage(c, 1, o, pp, device_flux_bcs, c);
  o.i_stage = 1;
  if (!rc) rc = launch_bcs(c);
  if (!rc) rc = hexed_b200_compute_euler(c, o);
  if (!rc) rc = launch_accumulate_time(c, c->d_step);
  return rc;
}

static int update_loop(hexed_b200_ctx* c, double safety, const ViscousStep* viscous, int n_cheby, int n_steps, int use_graph, double* last_dt, double* time_advanced)
{
  if (!c->have_mesh) return fail(c, HEXED_B200_NO_MESH, "no mesh uploaded");
  if (n_steps < 0 || n_cheby < 1) return fail(c, HEXED_B200_BAD_ARGUMENT, "bad step or Chebyshev count");
  // Solver::update (:842-851): nominal_dt = max_dt(safety/max_cheby, safety), dt = nominal_dt*chebyshev_step(n_cheby, i_cheby)
  const double safety_conv = safety/chebyshev_step(n_cheby, n_cheby - 1);
  auto enqueue = [&](int i_step) {
    const double f = chebyshev_step(n_cheby, i_step % n_cheby);
    return viscous ? enqueue_viscous_step(c, safety_conv, *viscous, f) : enqueue_euler_step(c, safety_conv, f);
  };
  const bool timing = c->timing;
  c->timing = false; // per-launch events cannot be queried inside a capture; this entry point is timed as a whole by its caller
  HB_CUDA(c, cudaMemsetAsync(c->d_step, 0, 3*sizeof(double), c->stream));
  c->dt_dev_active = c->d_step + 2;
  int rc = 0, done = 0;
  for (; !rc && done < n_steps && done < n_cheby; ++done) rc = enqueue(done); // one eager cycle: also performs every lazy allocation
#ifndef HB_EMULATE
  if (!rc && use_graph && n_steps - done >= n_cheby) {
    cudaGraph_t graph = nullptr; cudaGraphExec_t exec = nullptr;
    rc = check(c, cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal), "begin capture");
    if (!rc) {
      int rc_step = 0;
      for (int i = 0; i < n_cheby && !rc_step; ++i) rc_step = enqueue(i); // one whole Chebyshev cycle per graph
      const int rc_end = check(c, cudaStreamEndCapture(c->stream, &graph), "end capture");
      rc = rc_step ? rc_step : rc_end;
    }
    if (!rc) rc = check(c, cudaGraphInstantiate(&exec, graph, 0), "instantiate graph");
    for (; !rc && n_steps - done >= n_cheby; done += n_cheby) rc = check(c, cudaGraphLaunch(exec, c->stream), "launch graph");
    if (exec) cudaGraphExecDestroy(exec);
    if (graph) cudaGraphDestroy(graph);
  }
#else
  (void)use_graph;
#endif
  for (; !rc && done < n_steps; ++done) rc = enqueue(done);
  c->dt_dev_active = nullptr; c->max_dt_device_out = nullptr;
  c->timing = timing;
  if (rc) return rc;
  HB_CUDA(c, cudaMemcpyAsync(c->h_scalar, c->d_step + 2, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  HB_CUDA(c, cudaStreamSynchronize(c->stream));
  if (last_dt) *last_dt = *c->h_scalar;
  HB_CUDA(c, cudaMemcpyAsync(c->h_scalar, c->d_step + 1, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  HB_CUDA(c, cudaStreamSynchronize(c->stream));
  if (time_advanced) *time_advanced = *c->h_scalar;
  return 0;
}

int hexed_b200_update_euler(hexed_b200_ctx* c, double safety, int n_cheby, int n_steps, int use_graph, double* last_dt, double* time_advanced)
{ HB_ENTER(c); return update_loop(c, safety, nullptr, n_cheby, n_steps, use_graph, last_dt, time_advanced); }

int hexed_b200_update_navier_stokes(hexed_b200_ctx* c, double safety, hexed_b200_transport visc, hexed_b200_transport therm_cond,
                                    int n_cheby, int n_steps, int use_graph, double* last_dt, double* time_advanced)
{
  HB_ENTER(c);
  const ViscousStep v{safety, visc, therm_cond}; // max_dt(safety/max_cheby, safety): the diffusive safety is not divided (:849)
  return update_loop(c, safety, &v, n_cheby, n_steps, use_graph, last_dt, time_advanced);
}

/* ---- domain decomposition ---- */
int hexed_b200_set_partition(hexed_b200_ctx* c, int n_cut_car, int n_cut_def, int n_pre_prolong, const int* pre_prolong_ref)
{
  HB_ENTER(c);
  if (!c->have_mesh) return fail(c, HEXED_B200_NO_MESH, "no mesh uploaded");
  if (n_cut_car < 0 || n_cut_car > c->n_car_con || n_cut_def < 0 || n_cut_def > c->n_def_con || n_pre_prolong < 0)
    return fail(c, HEXED_B200_BAD_ARGUMENT, "cut connection counts out of range");
  for (int i = 0; i < n_pre_prolong; ++i) if (pre_prolong_ref[i] < 0 || pre_prolong_ref[i] >= c->n_ref) return fail(c, HEXED_B200_BAD_ARGUMENT, "refined face index out of range");
  dev_free(c->pre_prolong);
  int rc = dev_alloc(c, &c->pre_prolong, n_pre_prolong, false); if (rc) return rc;
  if (n_pre_prolong) HB_CUDA(c, cudaMemcpyAsync(c->pre_prolong, pre_prolong_ref, sizeof(int)*n_pre_prolong, cudaMemcpyHostToDevice, c->stream));
  HB_CUDA(c, cudaStreamSynchronize(c->stream));
  c->n_cut_car = n_cut_car; c->n_cut_def = n_cut_def; c->n_pre_prolong = n_pre_prolong;
  return 0;
}

int hexed_b200_face_list_gather(hexed_b200_ctx* c, int list_id, int kind, double* device_dst)
{
  HB_ENTER(c);
  if (list_id < 0 || list_id >= (int)c->lists.size()) return fail(c, HEXED_B200_BAD_ARGUMENT, "unknown face list");
  FaceList& l = c->lists[list_id];
  double* arr; int width;
  int rc = list_face_array(c, kind, &arr, &width); if (rc) return rc;
  return launch_gather_faces(c, arr, width, l.d_slots, l.n, device_dst);
}

int hexed_b200_face_list_scatter(hexed_b200_ctx* c, int list_id, int kind, const double* device_src)
{
  HB_ENTER(c);
  if (list_id < 0 || list_id >= (int)c->lists.size()) return fail(c, HEXED_B200_BAD_ARGUMENT, "unknown face list");
  FaceList& l = c->lists[list_id];
  double* arr; int width;
  int rc = list_face_array(c, kind, &arr, &width); if (rc) return rc;
  return launch_scatter_faces(c, arr, width, l.d_slots, l.n, device_src);
}

/* compute_euler in two halves so that the halo exchange overlaps the interior flux work:
 *   begin : Neighbor on the connections that touch no halo face
 *   finish: Prolong of hanging faces whose coarse face arrived through a halo, Neighbor on the cut connections, then the rest
 *           of the reference sequence (src/kernels_convective.cpp:8-16) */
int hexed_b200_compute_euler_begin(hexed_b200_ctx* c)
{
  HB_ENTER_KEEP(c);
  int rc;
  if ((rc = launch_neighbor_euler(c, 0, 0, c->n_car_con - c->n_cut_car))) return rc;
  if ((rc = launch_neighbor_euler(c, 1, 0, c->n_def_con - c->n_cut_def))) return rc;
  return 0;
}

int hexed_b200_compute_euler_finish(hexed_b200_ctx* c, hexed_b200_options o)
{
  HB_ENTER(c); // (scatters deferred ghost-face uploads: they have had the interior Neighbor kernels to arrive)
  int rc;
  if (c->n_pre_prolong && (rc = launch_prolong(c, 0, c->nv, 0, c->pre_prolong, c->n_pre_prolong))) return rc;
  if ((rc = launch_neighbor_euler(c, 0, c->n_car_con - c->n_cut_car, c->n_cut_car))) return rc;
  if ((rc = launch_neighbor_euler(c, 1, c->n_def_con - c->n_cut_def, c->n_cut_def))) return rc;
  if ((rc = launch_restrict(c, 0, c->nv, 1))) return rc;
  if ((rc = launch_local_euler(c, 0, o))) return rc;
  if ((rc = launch_local_euler(c, 1, o))) return rc;
  if ((rc = launch_prolong(c, 0, c->nv, 0))) return rc;
  return 0;
}

int hexed_b200_max_dt_euler(hexed_b200_ctx* c, hexed_b200_options, double safety_conv, double, int local_time, double* dt)
{ HB_ENTER(c); return launch_max_dt_euler(c, safety_conv, local_time, dt); }

/* ---- generic PDEs ---- */
static const GenericOps* generic_ops(int pde)
{
  switch (pde) {
    case 1: return &generic_ops_pde1;
    case 2: return &generic_ops_pde2;
    case 3: return &generic_ops_pde3;
    case 4: return &generic_ops_pde4;
  }
  return nullptr;
}

static PdeParams make_params(hexed_b200_ctx* c, int pde, hexed_b200_transport visc, hexed_b200_transport cond, double p0, double p1)
{
  PdeParams pp;
  std::memset(&pp, 0, sizeof(pp));
  pp.visc = visc; pp.cond = cond; pp.p0 = p0; pp.p1 = p1;
  pp.visc_inv_sqrt_ref = 1./visc.sqrt_ref_temp; pp.cond_inv_sqrt_ref = 1./cond.sqrt_ref_temp;
  // pde::Advection uses the nodes of Gauss_legendre(row_size) mapped to [-1, 1] whatever the solution basis is (include/pde.hpp:281)
  for (int i = 0; i < c->rs; ++i) pp.adv_nodes[i] = 2*c->gl_node[i] - 1;
  (void)pde;
  return pp;
}

static const hexed_b200_transport no_transport = {0., 0., 1., 1., 1., 0}; // Transport_model::inviscid (include/Transport_model.hpp:40)

static int n_extrap_of(hexed_b200_ctx* c, int pde) { return pde == 2 ? c->nd + c->rs : pde == 3 ? 3 : c->nv; }

/* reference src/kernels_convective.cpp:8-16 */
static int convection_stage(hexed_b200_ctx* c, int pde, hexed_b200_options o, const PdeParams& pp)
{
  if (!c->have_mesh) return fail(c, HEXED_B200_NO_MESH, "no mesh uploaded");
  const GenericOps* g = generic_ops(pde);
  const int kind = pde == 2 ? 2 : 0, ne = n_extrap_of(c, pde);
  int rc;
  if ((rc = g->neighbor(c, 0, pp, false, 0, -1))) return rc;
  if ((rc = g->neighbor(c, 1, pp, false, 0, -1))) return rc;
  if ((rc = launch_restrict(c, kind, ne, 1))) return rc;
  if ((rc = g->local(c, 0, o, pp, false))) return rc;
  if ((rc = g->local(c, 1, o, pp, false))) return rc;
  if ((rc = launch_prolong(c, kind, ne, 0))) return rc;
  return 0;
}

/* reference src/kernels_diffusive.cpp:8-26 */
static int diffusion_stage(hexed_b200_ctx* c, int pde, hexed_b200_options o, const PdeParams& pp, hexed_b200_callback flux_bc, void* user)
{
  if (!c->have_mesh) return fail(c, HEXED_B200_NO_MESH, "no mesh uploaded");
  const GenericOps* g = generic_ops(pde);
  const int ne = n_extrap_of(c, pde);
  int rc;
  if ((rc = g->neighbor(c, 0, pp, false, 0, -1))) return rc;
  if ((rc = g->neighbor(c, 1, pp, false, 0, -1))) return rc;
  if ((rc = launch_restrict(c, 0, ne, 1))) return rc;
  if ((rc = launch_restrict(c, 1, ne, 0))) return rc;
  if ((rc = g->local(c, 0, o, pp, false))) return rc;
  if ((rc = g->local(c, 1, o, pp, false))) return rc;
  if (!o.i_stage) {
    if ((rc = launch_prolong(c, 1, ne, 1))) return rc;
    if (flux_bc) flux_bc(user); // host callback = Solver::apply_flux_bcs; it may enqueue device work on this context's stream
    if ((rc = g->neighbor(c, 0, pp, true, 0, -1))) return rc;
    if ((rc = g->neighbor(c, 1, pp, true, 0, -1))) return rc;
    if ((rc = launch_restrict(c, 1, ne, 1))) return rc;
    if ((rc = g->local(c, 0, o, pp, true))) return rc;
    if ((rc = g->local(c, 1, o, pp, true))) return rc;
  }
  if ((rc = launch_prolong(c, 0, ne, 0))) return rc;
  return 0;
}

/* compute_navier_stokes in three parts around the two halo exchanges of a partitioned mesh (see include/hexed_b200.h) */
int hexed_b200_compute_navier_stokes_begin(hexed_b200_ctx* c, hexed_b200_options o, hexed_b200_transport visc, hexed_b200_transport therm_cond)
{
  HB_ENTER_KEEP(c);
  if (!c->have_mesh) return fail(c, HEXED_B200_NO_MESH, "no mesh uploaded");
  const PdeParams pp = make_params(c, 1, visc, therm_cond, 0., 0.);
  const GenericOps* g = generic_ops(1);
  int rc;
  if ((rc = g->neighbor(c, 0, pp, false, 0, c->n_car_con - c->n_cut_car))) return rc;
  if ((rc = g->neighbor(c, 1, pp, false, 0, c->n_def_con - c->n_cut_def))) return rc;
  return 0;
}

/* the middle part in its two halves: everything up to the Prolong of the viscous flux (what the flux boundary conditions read), and
 * the reconciliation of the interior connections that overlaps the second exchange. hexed_b200_compute_navier_stokes_middle runs both
 * around the callback; the device group (group.cu) runs them for every rank around ONE call of the host callback. */
int hexed_b200_compute_navier_stokes_middle_local(hexed_b200_ctx* c, hexed_b200_options o, hexed_b200_transport visc, hexed_b200_transport therm_cond)
{
  HB_ENTER(c);
  if (!c->have_mesh) return fail(c, HEXED_B200_NO_MESH, "no mesh uploaded");
  const PdeParams pp = make_params(c, 1, visc, therm_cond, 0., 0.);
  const GenericOps* g = generic_ops(1);
  const int ne = c->nv;
  int rc;
  if (c->n_pre_prolong && (rc = launch_prolong(c, 0, ne, 0, c->pre_prolong, c->n_pre_prolong))) return rc;
  if ((rc = g->neighbor(c, 0, pp, false, c->n_car_con - c->n_cut_car, c->n_cut_car))) return rc;
  if ((rc = g->neighbor(c, 1, pp, false, c->n_def_con - c->n_cut_def, c->n_cut_def))) return rc;
  if ((rc = launch_restrict(c, 0, ne, 1))) return rc;
  if ((rc = launch_restrict(c, 1, ne, 0))) return rc;
  if ((rc = g->local(c, 0, o, pp, false))) return rc;
  if ((rc = g->local(c, 1, o, pp, false))) return rc;
  if (!o.i_stage && (rc = launch_prolong(c, 1, ne, 1))) return rc;
  return 0;
}

int hexed_b200_compute_navier_stokes_middle_reconcile(hexed_b200_ctx* c, hexed_b200_options o, hexed_b200_transport visc, hexed_b200_transport therm_cond)
{
  HB_ENTER(c);
  if (!c->have_mesh) return fail(c, HEXED_B200_NO_MESH, "no mesh uploaded");
  if (o.i_stage) return 0;
  const PdeParams pp = make_params(c, 1, visc, therm_cond, 0., 0.);
  const GenericOps* g = generic_ops(1);
  int rc;
  // the viscous-flux faces of the cut connections travel now; the interior connections are reconciled meanwhile
  if ((rc = g->neighbor(c, 0, pp, true, 0, c->n_car_con - c->n_cut_car))) return rc;
  if ((rc = g->neighbor(c, 1, pp, true, 0, c->n_def_con - c->n_cut_def))) return rc;
  return 0;
}

int hexed_b200_compute_navier_stokes_middle(hexed_b200_ctx* c, hexed_b200_options o, hexed_b200_callback flux_bc, void* user,
                                            hexed_b200_transport visc, hexed_b200_transport therm_cond)
{
  int rc = hexed_b200_compute_navier_stokes_middle_local(c, o, visc, therm_cond);
  if (rc) return rc;
  if (!o.i_stage && flux_bc) flux_bc(user);
  return hexed_b200_compute_navier_stokes_middle_reconcile(c, o, visc, therm_cond);
}

/* max_dt of any of the five PDEs with the result LEFT ON THE DEVICE at `d_out` (no read-back, no synchronisation): what the device
 * group reduces across ranks with ncclAllReduce(min). pde: 0 Euler, 1 Navier-Stokes, 2 advection, 3 smooth AV, 4 fix therm admis. */
int hexed_b200_max_dt_device(hexed_b200_ctx* c, int pde, double sc, double sd, int local_time, hexed_b200_transport visc,
                             hexed_b200_transport therm_cond, double advect_length, double* d_out)
{
  HB_ENTER(c);
  if (!c->have_mesh) return fail(c, HEXED_B200_NO_MESH, "no mesh uploaded");
  if (pde < 0 || pde > 4) return fail(c, HEXED_B200_BAD_ARGUMENT, "unknown PDE");
  double dummy = 0.;
  if (local_time) { // no reduction: the kernels write time_step_scale and the answer is 1 (include/Spatial.hpp:827)
    if (pde == 0) return launch_max_dt_euler(c, sc, 1, &dummy);
    return generic_ops(pde)->max_dt(c, make_params(c, pde, visc, therm_cond, pde == 3 ? 1. : advect_length, pde == 3 ? 1. : 0.), sc, sd, 1, &dummy);
  }
  if (!c->n_elem) { HB_CUDA(c, cudaMemsetAsync(d_out, 0x7f, sizeof(double), c->stream)); return 0; }
  if (pde == 0) return launch_max_dt_euler_device(c, sc, d_out);
  c->max_dt_device_out = d_out;
  int rc = generic_ops(pde)->max_dt(c, make_params(c, pde, visc, therm_cond, pde == 3 ? 1. : advect_length, pde == 3 ? 1. : 0.), sc, sd, 0, &dummy);
  c->max_dt_device_out = nullptr;
  return rc;
}

int hexed_b200_compute_navier_stokes_finish(hexed_b200_ctx* c, hexed_b200_options o, hexed_b200_transport visc, hexed_b200_transport therm_cond)
{
  HB_ENTER(c);
  if (!c->have_mesh) return fail(c, HEXED_B200_NO_MESH, "no mesh uploaded");
  const PdeParams pp = make_params(c, 1, visc, therm_cond, 0., 0.);
  const GenericOps* g = generic_ops(1);
  const int ne = c->nv;
  int rc;
  if (!o.i_stage) {
    if (c->n_pre_prolong && (rc = launch_prolong(c, 1, ne, 1, c->pre_prolong, c->n_pre_prolong))) return rc;
    if ((rc = g->neighbor(c, 0, pp, true, c->n_car_con - c->n_cut_car, c->n_cut_car))) return rc;
    if ((rc = g->neighbor(c, 1, pp, true, c->n_def_con - c->n_cut_def, c->n_cut_def))) return rc;
    if ((rc = launch_restrict(c, 1, ne, 1))) return rc;
    if ((rc = g->local(c, 0, o, pp, true))) return rc;
    if ((rc = g->local(c, 1, o, pp, true))) return rc;
  }
  if ((rc = launch_prolong(c, 0, ne, 0))) return rc;
  return 0;
}

int hexed_b200_compute_advection(hexed_b200_ctx* c, hexed_b200_options o, double advect_length)
{ HB_ENTER(c); return convection_stage(c, 2, o, make_params(c, 2, no_transport, no_transport, advect_length, 0.)); }

int hexed_b200_compute_navier_stokes(hexed_b200_ctx* c, hexed_b200_options o, hexed_b200_callback flux_bc, void* user,
                                     hexed_b200_transport visc, hexed_b200_transport therm_cond)
{
  HB_ENTER_KEEP(c);
  if (c->n_pending_uploads) { // as in hexed_b200_compute_euler
    int rc = hexed_b200_compute_navier_stokes_begin(c, o, visc, therm_cond);
    if (!rc) rc = hexed_b200_compute_navier_stokes_middle(c, o, flux_bc, user, visc, therm_cond);
    if (!rc) rc = hexed_b200_compute_navier_stokes_finish(c, o, visc, therm_cond);
    return rc;
  }
  return diffusion_stage(c, 1, o, make_params(c, 1, visc, therm_cond, 0., 0.), flux_bc, user);
}

int hexed_b200_compute_smooth_av(hexed_b200_ctx* c, hexed_b200_options o, hexed_b200_callback flux_bc, void* user, double diff_time, double chebyshev_step)
{ HB_ENTER(c); return diffusion_stage(c, 3, o, make_params(c, 3, no_transport, no_transport, diff_time, chebyshev_step), flux_bc, user); }

int hexed_b200_compute_fix_therm_admis(hexed_b200_ctx* c, hexed_b200_options o, hexed_b200_callback flux_bc, void* user)
{ HB_ENTER(c); return diffusion_stage(c, 4, o, make_params(c, 4, no_transport, no_transport, 0., 0.), flux_bc, user); }

int hexed_b200_max_dt_navier_stokes(hexed_b200_ctx* c, hexed_b200_options, double sc, double sd, int local_time,
                                    hexed_b200_transport visc, hexed_b200_transport therm_cond, double* dt)
{
  HB_ENTER(c);
  if (!c->have_mesh) return fail(c, HEXED_B200_NO_MESH, "no mesh uploaded");
  return generic_ops(1)->max_dt(c, make_params(c, 1, visc, therm_cond, 0., 0.), sc, sd, local_time, dt);
}

int hexed_b200_max_dt_advection(hexed_b200_ctx* c, hexed_b200_options, double sc, double sd, int local_time, double advect_length, double* dt)
{
  HB_ENTER(c);
  if (!c->have_mesh) return fail(c, HEXED_B200_NO_MESH, "no mesh uploaded");
  return generic_ops(2)->max_dt(c, make_params(c, 2, no_transport, no_transport, advect_length, 0.), sc, sd, local_time, dt);
}

int hexed_b200_max_dt_smooth_av(hexed_b200_ctx* c, hexed_b200_options, double sc, double sd, int local_time, double* dt)
{
  HB_ENTER(c);
  if (!c->have_mesh) return fail(c, HEXED_B200_NO_MESH, "no mesh uploaded");
  return generic_ops(3)->max_dt(c, make_params(c, 3, no_transport, no_transport, 1., 1.), sc, sd, local_time, dt); // src/kernels_max_dt.cpp:20
}

int hexed_b200_max_dt_fix_therm_admis(hexed_b200_ctx* c, hexed_b200_options, double sc, double sd, int local_time, double* dt)
{
  HB_ENTER(c);
  if (!c->have_mesh) return fail(c, HEXED_B200_NO_MESH, "no mesh uploaded");
  return generic_ops(4)->max_dt(c, make_params(c, 4, no_transport, no_transport, 0., 0.), sc, sd, local_time, dt);
}

int hexed_b200_compute_write_face_advection(hexed_b200_ctx* c)
{
  HB_ENTER(c);
  if (!c->have_mesh) return fail(c, HEXED_B200_NO_MESH, "no mesh uploaded");
  return generic_ops(2)->write_face(c, make_params(c, 2, no_transport, no_transport, 1., 0.)); // src/kernels_convective.cpp:48-51
}

int hexed_b200_compute_write_face_smooth_av(hexed_b200_ctx* c)
{
  HB_ENTER(c);
  if (!c->have_mesh) return fail(c, HEXED_B200_NO_MESH, "no mesh uploaded");
  return generic_ops(3)->write_face(c, make_params(c, 3, no_transport, no_transport, 1., 1.)); // src/kernels_convective.cpp:53-56
}

int hexed_b200_compute_prolong_advection(hexed_b200_ctx* c)
{
  HB_ENTER(c);
  if (!c->have_mesh) return fail(c, HEXED_B200_NO_MESH, "no mesh uploaded");
  if (!c->face_wide) { int rc = dev_alloc(c, &c->face_wide, (size_t)c->n_face_slot*(c->nd + c->rs)*c->nfq); if (rc) return rc; }
  return launch_prolong(c, 2, c->nd + c->rs, 0); // src/kernels_convective.cpp:33-36
}

int hexed_b200_stabilizing_art_visc(hexed_b200_ctx* c, double char_speed) { HB_ENTER(c); return launch_stab_art_visc(c, char_speed); }

int hexed_b200_apply_flux_bcs(hexed_b200_ctx* c) { HB_ENTER(c); return launch_flux_bcs(c); }

int hexed_b200_set_jacobian(hexed_b200_ctx* c, const double* vertex_pos, const double* node_adj)
{
  HB_ENTER(c);
  if (!c->have_mesh) return fail(c, HEXED_B200_NO_MESH, "no mesh uploaded");
  if (c->n_def && !vertex_pos) return fail(c, HEXED_B200_BAD_ARGUMENT, "vertex positions of the deformed elements are required");
  double *d_vert = nullptr, *d_adj = nullptr;
  const size_t n_vert_d = (size_t)c->n_def*c->n_vert*c->nd, n_adj = (size_t)c->n_def*2*c->nd*c->nfq;
  int rc = 0;
  if (c->n_def) {
    rc = dev_alloc(c, &d_vert, n_vert_d, false);
    if (!rc) rc = check(c, cudaMemcpyAsync(d_vert, vertex_pos, sizeof(double)*n_vert_d, cudaMemcpyHostToDevice, c->stream), "upload vertex positions");
    if (!rc && node_adj) {
      rc = dev_alloc(c, &d_adj, n_adj, false);
      if (!rc) rc = check(c, cudaMemcpyAsync(d_adj, node_adj, sizeof(double)*n_adj, cudaMemcpyHostToDevice, c->stream), "upload node adjustments");
    }
  }
  if (!rc) rc = launch_set_jacobian(c, d_vert, d_adj);
  if (!rc) rc = check(c, cudaStreamSynchronize(c->stream), "set_jacobian");
  dev_free(d_vert); dev_free(d_adj);
  return rc;
}

int hexed_b200_calc_shared_normals(hexed_b200_ctx* c) { HB_ENTER(c); return launch_shared_normals(c); }

int hexed_b200_av_scale_velocity(hexed_b200_ctx* c, int restore) { HB_ENTER(c); return launch_av_scale_velocity(c, restore); }

int hexed_b200_av_project_forcing(hexed_b200_ctx* c, const double* node_weights, const double* orthogonal)
{
  HB_ENTER(c);
  if (!node_weights || !orthogonal) return fail(c, HEXED_B200_BAD_ARGUMENT, "null operator");
  return launch_av_project_forcing(c, node_weights, orthogonal);
}

int hexed_b200_av_finish(hexed_b200_ctx* c, double mult, double us_max, int n_real, const double* node_weights, double* residual)
{
  HB_ENTER(c);
  if (!node_weights || !residual) return fail(c, HEXED_B200_BAD_ARGUMENT, "null argument");
  double sq = 0;
  int rc = launch_av_finish(c, mult, us_max, n_real, node_weights, &sq);
  if (!rc) *residual = sqrt(sq);
  return rc;
}

int hexed_b200_interp_vertices(hexed_b200_ctx* c, int target, const double* vertex_values, const double* interp)
{
  HB_ENTER(c);
  if (!c->have_mesh) return fail(c, HEXED_B200_NO_MESH, "no mesh uploaded");
  if (!vertex_values || !interp) return fail(c, HEXED_B200_BAD_ARGUMENT, "null argument");
  double *d_vert = nullptr, *d_interp = nullptr;
  const size_t nv = (size_t)c->n_elem*c->n_vert;
  int rc = dev_alloc(c, &d_vert, nv, false);
  if (!rc) rc = dev_alloc(c, &d_interp, (size_t)2*c->rs, false);
  if (!rc && nv) rc = check(c, cudaMemcpyAsync(d_vert, vertex_values, sizeof(double)*nv, cudaMemcpyHostToDevice, c->stream), "upload vertex values");
  if (!rc) rc = check(c, cudaMemcpyAsync(d_interp, interp, sizeof(double)*2*c->rs, cudaMemcpyHostToDevice, c->stream), "upload interpolation matrix");
  if (!rc) rc = launch_interp_vertices(c, target, d_vert, d_interp);
  if (!rc) rc = check(c, cudaStreamSynchronize(c->stream), "interp_vertices");
  dev_free(d_vert); dev_free(d_interp);
  return rc;
}

int hexed_b200_av_swap(hexed_b200_ctx* c) { HB_ENTER(c); return launch_av_swap(c); }

int hexed_b200_av_elwise_ramp(hexed_b200_ctx* c, double scale) { HB_ENTER(c); return launch_av_elwise_ramp(c, scale); }
int hexed_b200_av_elwise_forcing(hexed_b200_ctx* c, int restore)
{
  HB_ENTER(c);
  if (restore != 0 && restore != 1) return fail(c, HEXED_B200_BAD_ARGUMENT, "restore must be 0 or 1");
  return launch_av_elwise_forcing(c, restore);
}
int hexed_b200_av_elwise_vertices(hexed_b200_ctx* c, const double* interp)
{
  HB_ENTER(c);
  if (!c->have_mesh) return fail(c, HEXED_B200_NO_MESH, "no mesh uploaded");
  if (!interp) return fail(c, HEXED_B200_BAD_ARGUMENT, "null interpolation matrix");
  double* d_interp = nullptr;
  int rc = dev_alloc(c, &d_interp, (size_t)2*c->rs, false);
  if (!rc) rc = check(c, cudaMemcpyAsync(d_interp, interp, sizeof(double)*2*c->rs, cudaMemcpyHostToDevice, c->stream), "upload interpolation matrix");
  if (!rc) rc = launch_av_elwise_vertices(c, d_interp);
  if (!rc) rc = check(c, cudaStreamSynchronize(c->stream), "av_elwise_vertices");
  dev_free(d_interp);
  return rc;
}
int hexed_b200_apply_aux_bcs(hexed_b200_ctx* c, int mode) { HB_ENTER(c); return launch_aux_bcs(c, mode); }

int hexed_b200_vertex_topology(hexed_b200_ctx* c, const int* elem_vertex, int n_vertex, const int* matchers, int n_match)
{
  HB_ENTER(c);
  if (!c->have_mesh) return fail(c, HEXED_B200_NO_MESH, "no mesh uploaded");
  if (n_vertex < 0 || n_match < 0 || (c->n_elem && !elem_vertex) || (n_match && !matchers)) return fail(c, HEXED_B200_BAD_ARGUMENT, "bad vertex topology");
  const size_t n = (size_t)c->n_elem*c->n_vert;
  for (size_t i = 0; i < n; ++i) if (elem_vertex[i] < 0 || elem_vertex[i] >= n_vertex) return fail(c, HEXED_B200_BAD_ARGUMENT, "vertex id out of range");
  for (int i = 0; i < n_match; ++i) {
    const int* row = matchers + (size_t)i*8;
    if (row[0] < 0 || row[0] >= c->nd) return fail(c, HEXED_B200_BAD_ARGUMENT, "hanging vertex matcher: bad dimension");
    for (int k = 0; k < 4; ++k) if (row[4 + k] >= c->n_elem) return fail(c, HEXED_B200_BAD_ARGUMENT, "hanging vertex matcher: element out of range");
  }
  dev_free(c->elem_vertex); dev_free(c->matchers); dev_free(c->vertex_vals);
  int rc = dev_alloc(c, &c->elem_vertex, n, false);
  if (!rc) rc = dev_alloc(c, &c->matchers, (size_t)n_match*8, false);
  if (!rc) rc = dev_alloc(c, &c->vertex_vals, (size_t)n_vertex, true);
  if (!rc && n) rc = check(c, cudaMemcpyAsync(c->elem_vertex, elem_vertex, sizeof(int)*n, cudaMemcpyHostToDevice, c->stream), "upload vertex ids");
  if (!rc && n_match) rc = check(c, cudaMemcpyAsync(c->matchers, matchers, sizeof(int)*8*n_match, cudaMemcpyHostToDevice, c->stream), "upload matchers");
  if (!rc) rc = check(c, cudaStreamSynchronize(c->stream), "vertex_topology");
  if (rc) return rc;
  c->n_vertex = n_vertex; c->n_match = n_match;
  return 0;
}

int hexed_b200_share_vertex_data(hexed_b200_ctx* c, int which, int op)
{
  HB_ENTER(c);
  if (!c->have_mesh) return fail(c, HEXED_B200_NO_MESH, "no mesh uploaded");
  if (op != 0 && op != 1) return fail(c, HEXED_B200_BAD_ARGUMENT, "op must be 0 (min) or 1 (max)");
  double** arr; size_t item, count;
  if (which != HEXED_B200_VERTEX_TSS && which != HEXED_B200_VERTEX_SCRATCH) return fail(c, HEXED_B200_BAD_ARGUMENT, "not a per-vertex array");
  int rc = array_info(c, which, &arr, &item, &count, true); if (rc) return rc;
  if (which == HEXED_B200_VERTEX_TSS) invalidate_cfl_cache(c);
  return launch_share_vertex_data(c, *arr, op);
}

int hexed_b200_fix_admis_spread(hexed_b200_ctx* c, const double* interp)
{
  HB_ENTER(c);
  if (!c->have_mesh) return fail(c, HEXED_B200_NO_MESH, "no mesh uploaded");
  if (!interp) return fail(c, HEXED_B200_BAD_ARGUMENT, "null interpolation matrix");
  double* d_interp = nullptr;
  int rc = dev_alloc(c, &d_interp, (size_t)2*c->rs, false);
  if (!rc) rc = check(c, cudaMemcpyAsync(d_interp, interp, sizeof(double)*2*c->rs, cudaMemcpyHostToDevice, c->stream), "upload interpolation matrix");
  if (!rc) rc = launch_fix_admis_spread(c, d_interp);
  if (!rc) rc = check(c, cudaStreamSynchronize(c->stream), "fix_admis_spread");
  dev_free(d_interp);
  return rc;
}

int hexed_b200_is_admissible(hexed_b200_ctx* c, int* admissible)
{
  HB_ENTER(c);
  if (!admissible) return fail(c, HEXED_B200_BAD_ARGUMENT, "null result pointer");
  return launch_is_admissible(c, admissible);
}

int hexed_b200_is_admissible_begin(hexed_b200_ctx* c) { HB_ENTER(c); return launch_is_admissible_begin(c); }
int hexed_b200_is_admissible_finish(hexed_b200_ctx* c, int* admissible)
{
  HB_ENTER_KEEP(c);
  if (!admissible) return fail(c, HEXED_B200_BAD_ARGUMENT, "null result pointer");
  return launch_is_admissible_finish(c, admissible);
}

int hexed_b200_download_record(hexed_b200_ctx* c, int* dst, int first_elem, int n_elem)
{
  HB_ENTER(c);
  if (!c->have_mesh) return fail(c, HEXED_B200_NO_MESH, "no mesh uploaded");
  if (first_elem < 0 || n_elem < 0 || first_elem + n_elem > c->n_elem) return fail(c, HEXED_B200_BAD_ARGUMENT, "element range out of bounds");
  if (!c->record) return fail(c, HEXED_B200_BAD_ARGUMENT, "no record yet: call hexed_b200_is_admissible first");
  HB_CUDA(c, cudaMemcpyAsync(dst, c->record + first_elem, sizeof(int)*n_elem, cudaMemcpyDeviceToHost, c->stream));
  HB_CUDA(c, cudaStreamSynchronize(c->stream));
  return 0;
}

/* individual kernels of a generic PDE, for unit-level parity: which = 0 Neighbor, 1 Local, 2 Neighbor_reconcile, 3 Reconcile_ldg_flux */
int hexed_b200_pde_kernel(hexed_b200_ctx* c, int pde, int which, int deformed, hexed_b200_options o,
                          hexed_b200_transport visc, hexed_b200_transport therm_cond, double p0, double p1)
{
  HB_ENTER(c);
  if (!c->have_mesh) return fail(c, HEXED_B200_NO_MESH, "no mesh uploaded");
  const GenericOps* g = generic_ops(pde);
  if (!g) return fail(c, HEXED_B200_BAD_ARGUMENT, "pde must be 1 (Navier-Stokes), 2 (advection), 3 (smooth AV) or 4 (fix therm admis)");
  const PdeParams pp = make_params(c, pde, visc, therm_cond, p0, p1);
  switch (which) {
    case 0: return g->neighbor(c, deformed, pp, false, 0, -1);
    case 1: return g->local(c, deformed, o, pp, false);
    case 2: return g->neighbor(c, deformed, pp, true, 0, -1);
    case 3: return g->local(c, deformed, o, pp, true);
  }
  return fail(c, HEXED_B200_BAD_ARGUMENT, "unknown kernel id");
}

int hexed_b200_compute_write_face(hexed_b200_ctx* c) { HB_ENTER(c); return launch_write_face(c); }

int hexed_b200_compute_prolong(hexed_b200_ctx* c, int scale, int offset)
{
  HB_ENTER(c);
  if (offset && !c->face_ldg) { int rc = dev_alloc(c, &c->face_ldg, (size_t)c->n_face_slot*c->nv*c->nfq); if (rc) return rc; }
  return launch_prolong(c, offset ? 1 : 0, c->nv, scale);
}

int hexed_b200_compute_restrict(hexed_b200_ctx* c, int scale, int offset)
{
  HB_ENTER(c);
  if (offset && !c->face_ldg) { int rc = dev_alloc(c, &c->face_ldg, (size_t)c->n_face_slot*c->nv*c->nfq); if (rc) return rc; }
  return launch_restrict(c, offset ? 1 : 0, c->nv, scale);
}

int hexed_b200_neighbor_euler(hexed_b200_ctx* c, int deformed) { HB_ENTER(c); return launch_neighbor_euler(c, deformed); }
int hexed_b200_local_euler(hexed_b200_ctx* c, int deformed, hexed_b200_options o) { HB_ENTER(c); return launch_local_euler(c, deformed, o); }

int hexed_b200_bc_create(hexed_b200_ctx* c, int kind, int n, const int* inside, const int* ghost, const int* normal,
                         const double* params, int n_params, int* bc_id)
{
  HB_ENTER(c);
  if (!c->have_mesh) return fail(c, HEXED_B200_NO_MESH, "no mesh uploaded");
  if (kind < 0 || kind > HEXED_B200_BC_RIEMANN_INVARIANTS) return fail(c, HEXED_B200_BAD_ARGUMENT, "unknown boundary condition kind");
  if ((kind == HEXED_B200_BC_FREESTREAM || kind == HEXED_B200_BC_RIEMANN_INVARIANTS) && n_params != c->nv)
    return fail(c, HEXED_B200_BAD_ARGUMENT, "freestream needs n_dim + 2 parameters");
  if (kind == HEXED_B200_BC_PRESSURE_OUTFLOW && n_params != 1) return fail(c, HEXED_B200_BAD_ARGUMENT, "pressure outflow needs 1 parameter");
  if (kind == HEXED_B200_BC_NO_SLIP && n_params != 6) return fail(c, HEXED_B200_BAD_ARGUMENT, "no-slip needs 6 parameters");
  const bool needs_normal = kind == HEXED_B200_BC_NONPENETRATION || kind == HEXED_B200_BC_PRESSURE_OUTFLOW || kind == HEXED_B200_BC_NO_SLIP
                            || kind == HEXED_B200_BC_RIEMANN_INVARIANTS;
  const bool needs_sign = kind == HEXED_B200_BC_PRESSURE_OUTFLOW || kind == HEXED_B200_BC_NO_SLIP || kind == HEXED_B200_BC_RIEMANN_INVARIANTS; // inside_face_sign is read off the slot number
  for (int i = 0; i < n; ++i) {
    if (ghost[i] < 0 || ghost[i] >= c->n_face_slot || inside[i] < 0 || inside[i] >= c->n_face_slot) return fail(c, HEXED_B200_BAD_ARGUMENT, "face slot out of range");
    if (needs_normal && (normal[i] < 0 || normal[i] >= c->n_normal_slot)) return fail(c, HEXED_B200_BAD_ARGUMENT, "normal slot out of range");
    if (needs_sign && inside[i] >= 2*c->nd*c->n_elem) return fail(c, HEXED_B200_BAD_ARGUMENT, "inside face of a boundary condition must be an element face");
  }
  Bc b; b.kind = kind; b.n = n; b.n_params = n_params;
  int rc = 0;
  if (!rc) rc = dev_alloc(c, &b.inside, n, false);
  if (!rc) rc = dev_alloc(c, &b.ghost, n, false);
  if (!rc) rc = dev_alloc(c, &b.normal, n, true);
  if (!rc) rc = dev_alloc(c, &b.params, n_params, false);
  if (!rc && (kind == HEXED_B200_BC_NO_SLIP || kind == HEXED_B200_BC_RIEMANN_INVARIANTS)) rc = dev_alloc(c, &b.cache, (size_t)n*c->nv*c->nfq, true);
  if (rc) return rc;
  if (n) {
    HB_CUDA(c, cudaMemcpyAsync(b.inside, inside, sizeof(int)*n, cudaMemcpyHostToDevice, c->stream));
    HB_CUDA(c, cudaMemcpyAsync(b.ghost, ghost, sizeof(int)*n, cudaMemcpyHostToDevice, c->stream));
    if (normal) HB_CUDA(c, cudaMemcpyAsync(b.normal, normal, sizeof(int)*n, cudaMemcpyHostToDevice, c->stream));
  }
  if (n_params) HB_CUDA(c, cudaMemcpyAsync(b.params, params, sizeof(double)*n_params, cudaMemcpyHostToDevice, c->stream));
  HB_CUDA(c, cudaStreamSynchronize(c->stream));
  c->bcs.push_back(b);
  *bc_id = (int)c->bcs.size() - 1;
  return 0;
}

int hexed_b200_bc_set_params(hexed_b200_ctx* c, int bc_id, const double* params, int n_params)
{
  HB_ENTER(c);
  if (bc_id < 0 || bc_id >= (int)c->bcs.size()) return fail(c, HEXED_B200_BAD_ARGUMENT, "unknown boundary condition");
  Bc& b = c->bcs[bc_id];
  if (n_params != b.n_params || (n_params && !params)) return fail(c, HEXED_B200_BAD_ARGUMENT, "parameter count differs from the registered one");
  // stream-ordered: every apply_*_bcs enqueued so far still reads the old block (small pageable sources are staged by the runtime at call time)
  if (n_params) HB_CUDA(c, cudaMemcpyAsync(b.params, params, sizeof(double)*n_params, cudaMemcpyHostToDevice, c->stream));
  return 0;
}

int hexed_b200_apply_state_bcs(hexed_b200_ctx* c) { HB_ENTER(c); return launch_bcs(c); }

int hexed_b200_set_timing(hexed_b200_ctx* c, int enabled) { c->timing = enabled != 0; return 0; }

int hexed_b200_set_option(hexed_b200_ctx* c, int option, int value)
{
  HB_ENTER(c);
  if (option == HEXED_B200_OPT_PIPELINED_LOCAL) { c->use_pipe = value != 0; c->pipe_lean = value != 2; c->pipe_lean4 = value == 3; c->pipe_end_barrier = value != 4; return 0; }
  if (option == HEXED_B200_OPT_CFL_CACHE) { c->use_cfl_cache = value != 0; invalidate_cfl_cache(c); return 0; }
  if (option == HEXED_B200_OPT_FUSED_ADMIS) { // a change of the setting forgets the bits; setting it again does not
    if (c->use_fused_admis != (value != 0)) { c->use_fused_admis = value != 0; invalidate_admis(c); }
    return 0;
  }
  if (option == HEXED_B200_OPT_NS_LOCAL_LAYOUT) { if (value < 0 || value > 2) return fail(c, HEXED_B200_BAD_ARGUMENT, "layout 0, 1 or 2"); c->ns_layout = value; return 0; }
  return fail(c, HEXED_B200_BAD_ARGUMENT, "unknown option");
}

int hexed_b200_kernel_stats(hexed_b200_ctx* c, hexed_b200_kernel_stat* out, int capacity, int* n_out)
{
  HB_ENTER(c);
  int n = 0;
  for (int i = 0; i < ST_COUNT && n < capacity; ++i, ++n) {
    out[n].name = c->stats[i].name; out[n].deformed = c->stats[i].deformed; out[n].work_units = c->stats[i].work_units;
    out[n].launches = c->stats[i].launches; out[n].device_seconds = c->stats[i].seconds;
  }
  *n_out = n;
  return 0;
}

int hexed_b200_reset_stats(hexed_b200_ctx* c)
{
  HB_ENTER(c);
  for (int i = 0; i < ST_COUNT; ++i) { c->stats[i].work_units = 0; c->stats[i].launches = 0; c->stats[i].seconds = 0; }
  return 0;
}

long long hexed_b200_launch_count(const hexed_b200_ctx* c) { return c->launches; }

} // extern "C"
