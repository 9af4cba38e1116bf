/* harness.cpp -- TEST harness for adapter.cpp: a pointer-graph mesh shaped like the reference's, driven from Python.
 *
 * The genuine Hexed mesh classes cannot be built here (Eigen/HDF5 absent), so this file stands a mesh up the way the reference
 * lays it out in memory: every element owns one heap block of slots (src/Element.cpp:114-142), every face lives in storage of
 * its own that elements and connections merely alias (include/connection.hpp:59-86,111-123), deformed elements carry their
 * Jacobian data and face-normal pointers (src/Deformed_element.cpp:10,148-175), and the kernels see all of it only through
 * Sequence<Kernel_element&> / Sequence<Kernel_connection&> / Sequence<Refined_face&> views. Allocation order is shuffled so
 * nothing is accidentally contiguous. The exported hbh_* functions let tests/test_host_adapter.py fill this mesh from a
 * FlatMesh, call the hexed:: entry points exactly as Solver does, and read the host objects back.
 */
#include "adapter.hpp"

#include <algorithm>
#include <omp.h>
#include <cstring>
#include <random>
#include <unordered_map>

namespace
{

using namespace hexed;

int ipow(int b, int e) {int r = 1; for (int i = 0; i < e; ++i) r *= b; return r;}

class H_basis : public Basis
{
  std::vector<double> p; // packed layout of include/hexed_b200.h
  int rs;
  const double* at(int offset) const {return p.data() + offset;}
  Plain_mat mat(int offset, int rows, int cols) const
  {
    Plain_mat m(rows, cols);
    for (int i = 0; i < rows; ++i) for (int j = 0; j < cols; ++j) m(i, j) = p[offset + i*cols + j];
    return m;
  }
  protected:
  double min_eig_convection() const override {return p[4*rs + 7*rs*rs];}
  double quadratic_safety() const override {return p[4*rs + 7*rs*rs + 2];}
  public:
  H_basis(int row_size_arg, const double* packed, int n) : Basis(row_size_arg), p(packed, packed + n), rs{row_size_arg} {}
  double node(int i) const override {return p[i];}
  Plain_mat node_weights() const override {return mat(rs, rs, 1);}
  Plain_mat diff_mat() const override {return mat(2*rs, rs, rs);}
  Plain_mat boundary() const override {return mat(2*rs + rs*rs, 2, rs);}
  Plain_mat orthogonal(int degree) const override {return mat(4*rs + rs*rs + degree*rs, rs, 1);}
  Plain_mat filter() const override {return mat(4*rs + 2*rs*rs, rs, rs);}
  Plain_mat prolong(int i_half) const override {return mat(4*rs + 3*rs*rs + i_half*rs*rs, rs, rs);}
  Plain_mat restrict(int i_half) const override {return mat(4*rs + 5*rs*rs + i_half*rs*rs, rs, rs);}
  double min_eig_diffusion() const override {return p[4*rs + 7*rs*rs + 1];}
};

struct H_element : public Kernel_element
{
  int nd, nq, nv, n_slots, cache_slot;
  bool is_def;
  std::vector<double> data, jac, vertex_data;
  std::vector<double*> faces, face_normals;
  double nom = 1., uncertainty = 0.;
  double* state() override {return data.data();}
  double* residual_cache() override {return data.data() + size_t(cache_slot)*nq;}
  double* time_step_scale() override {return data.data() + size_t(nv)*nq;}
  double& vertex_time_step_scale(int i_vertex) override {return vertex_data[i_vertex];}
  double nominal_size() override {return nom;}
  double* face(int i_face, bool is_ldg) override {return faces[i_face] ? faces[i_face] + (is_ldg ? ldg_offset : 0) : nullptr;}
  bool deformed() const override {return is_def;}
  double* reference_level_normals() override {return is_def ? jac.data() : nullptr;}
  double* jacobian_determinant() override {return is_def ? jac.data() + size_t(nd)*nd*nq : nullptr;}
  double* kernel_face_normal(int i_face) override {return is_def ? face_normals[i_face] : nullptr;}
  double& uncert() override {return uncertainty;}
  int ldg_offset = 0;
};

struct H_connection : public Kernel_connection
{
  Connection_direction dir;
  double* side [2] {};
  double* nrml = nullptr;
  int ldg_offset = 0;
  Connection_direction get_direction() override {return dir;}
  double* state(int i_side, bool is_ldg) override {return side[i_side] + (is_ldg ? ldg_offset : 0);}
  double* normal() override {return nrml;}
};

template <typename T, typename S>
class Vec_seq : public Sequence<T&>
{
  std::vector<S*> v;
  public:
  void push(S* s) {v.push_back(s);}
  int size() override {return int(v.size());}
  T& operator[](int i) override {return *v[i];}
};

struct Harness
{
  int nd, rs, nq, nfq, nv, n_slots, n_car, n_def, n_face_slot, n_normal_slot, face_sz;
  std::unique_ptr<H_basis> basis;
  std::vector<std::unique_ptr<H_element>> elems;
  std::vector<std::unique_ptr<H_connection>> car_cons, def_cons;
  std::vector<std::unique_ptr<Refined_face>> refs;
  std::vector<std::unique_ptr<std::vector<double>>> face_store, normal_store; // by harness slot; normal may be null
  Vec_seq<Kernel_element, H_element> s_car, s_def, s_all;
  Vec_seq<Kernel_connection, H_connection> s_ccon, s_dcon;
  Vec_seq<Refined_face, Refined_face> s_ref;
  Stopwatch_tree sw_car {"element", {{"neighbor", Stopwatch_tree("connection")}, {"local", Stopwatch_tree("element")},
                                     {"reconcile LDG flux", Stopwatch_tree("element")}, {"compute time step", Stopwatch_tree("element")}}};
  Stopwatch_tree sw_def = sw_car;
  Stopwatch_tree sw_pr {"refined face"};
  std::function<void()> flux_bc;
  std::string error;
  Kernel_mesh mesh() {return {nd, rs, *basis, s_ccon, s_dcon, s_car, s_def, s_all, s_ref};}
  Kernel_options options(const double* a) {return {sw_car, sw_def, sw_pr, a[0], int(a[1]), a[2] != 0, a[3] != 0};}
};

Transport_model make_transport(const double* a)
{
  if (a[0] == 0) return Transport_model::inviscid();
  if (a[0] == 1) return Transport_model::constant(a[1]);
  return Transport_model::sutherland(a[1], a[2], a[3]);
}

} // namespace

extern "C" {

void* hbh_create(int nd, int rs, const double* packed_basis, int n_packed, int n_car, int n_def, int n_face_slot, int n_normal_slot,
                 const int* car_con, int n_car_con, const int* def_con, int n_def_con, const int* ref_face, int n_ref,
                 const int* normal_present, unsigned seed)
{
  auto* h = new Harness;
  h->nd = nd; h->rs = rs; h->nq = ipow(rs, nd); h->nfq = h->nq/rs; h->nv = nd + 2;
  h->n_slots = h->nv + 3 + 4 + rs + std::max(h->nv, rs); // src/Storage_params.cpp:32-35
  h->n_car = n_car; h->n_def = n_def; h->n_face_slot = n_face_slot; h->n_normal_slot = n_normal_slot;
  h->face_sz = std::max({2*h->nv*h->nfq, (nd + rs)*h->nfq, 3*h->nv*h->nfq}); // include/connection.hpp:62,233
  h->basis.reset(new H_basis(rs, packed_basis, n_packed));
  std::mt19937 rng(seed);
  // face and normal storage, allocated in shuffled order with random padding so no two slots are neighbours by construction
  std::vector<int> order(n_face_slot);
  for (int i = 0; i < n_face_slot; ++i) order[i] = i;
  std::shuffle(order.begin(), order.end(), rng);
  h->face_store.resize(n_face_slot);
  std::vector<std::unique_ptr<std::vector<double>>> padding;
  for (int s : order) {
    h->face_store[s].reset(new std::vector<double>(h->face_sz, 0.));
    if (rng() % 3 == 0) padding.emplace_back(new std::vector<double>(1 + rng() % 97));
  }
  h->normal_store.resize(n_normal_slot);
  for (int s = 0; s < n_normal_slot; ++s) if (normal_present[s]) h->normal_store[s].reset(new std::vector<double>(size_t(nd)*h->nfq, 0.));
  const int ne = n_car + n_def, nf = 2*nd;
  for (int e = 0; e < ne; ++e) {
    std::unique_ptr<H_element> el(new H_element);
    el->nd = nd; el->nq = h->nq; el->nv = h->nv; el->n_slots = h->n_slots; el->cache_slot = h->nv + 7 + rs;
    el->is_def = e >= n_car; el->ldg_offset = h->nv*h->nfq;
    el->data.assign(size_t(h->n_slots)*h->nq, 0.);
    if (el->is_def) el->jac.assign(size_t(nd*nd + 1)*h->nq, 0.);
    el->vertex_data.assign(ipow(2, nd), 1.);
    el->faces.assign(nf, nullptr); el->face_normals.assign(nf, nullptr);
    for (int f = 0; f < nf; ++f) {
      int s = e*nf + f;
      el->faces[f] = h->face_store[s]->data(); // a valid Hexed mesh connects every face (the kernels dereference all 2*n_dim of them)
      if (el->is_def) {
        int n = (e - n_car)*nf + f;
        if (h->normal_store[n]) el->face_normals[f] = h->normal_store[n]->data();
      }
    }
    (el->is_def ? h->s_def : h->s_car).push(el.get());
    h->elems.push_back(std::move(el));
  }
  for (auto& el : h->elems) h->s_all.push(el.get());
  for (int i = 0; i < n_car_con; ++i) {
    std::unique_ptr<H_connection> c(new H_connection);
    const int* t = car_con + i*3;
    c->side[0] = h->face_store[t[0]]->data(); c->side[1] = h->face_store[t[1]]->data();
    // a Cartesian connection joins a positive face to a negative face of the same dimension (include/connection.hpp:15-25)
    c->dir.i_dim = {t[2], t[2]}; c->dir.face_sign = {true, false};
    c->ldg_offset = h->nv*h->nfq;
    h->s_ccon.push(c.get()); h->car_cons.push_back(std::move(c));
  }
  for (int i = 0; i < n_def_con; ++i) {
    std::unique_ptr<H_connection> c(new H_connection);
    const int* t = def_con + i*7;
    c->side[0] = h->face_store[t[0]]->data(); c->side[1] = h->face_store[t[1]]->data();
    c->dir.i_dim = {t[2], t[3]}; c->dir.face_sign = {t[4] != 0, t[5] != 0};
    c->nrml = h->normal_store[t[6]]->data();
    c->ldg_offset = h->nv*h->nfq;
    h->s_dcon.push(c.get()); h->def_cons.push_back(std::move(c));
  }
  for (int i = 0; i < n_ref; ++i) {
    std::unique_ptr<Refined_face> r(new Refined_face);
    const int* t = ref_face + i*7;
    r->coarse = h->face_store[t[0]]->data();
    for (int k = 0; k < 4; ++k) r->fine[k] = t[1 + k] >= 0 ? h->face_store[t[1 + k]]->data() : nullptr;
    r->stretch = {t[5] != 0, t[6] != 0};
    h->s_ref.push(r.get()); h->refs.push_back(std::move(r));
  }
  return h;
}

void hbh_destroy(void* handle) {hexed_b200::invalidate(); delete static_cast<Harness*>(handle);}
const char* hbh_error(void* handle) {return static_cast<Harness*>(handle)->error.c_str();}
void hbh_set_flux_bc(void* handle, void (*cb)()) {auto* h = static_cast<Harness*>(handle); if (cb) h->flux_bc = cb; else h->flux_bc = nullptr;}

// host objects <- flat arrays (any pointer may be null)
void hbh_put(void* handle, const double* elem_data, const double* nom, const double* vtss, const double* ref_normals, const double* det,
             const double* normals, const double* face_state, const double* face_ldg, const double* face_wide)
{
  auto* h = static_cast<Harness*>(handle);
  const int ne = int(h->elems.size()), nq = h->nq, nd = h->nd, nfq = h->nfq, nv = h->nv, n_vert = ipow(2, nd);
  for (int e = 0; e < ne; ++e) {
    auto& el = *h->elems[e];
    if (elem_data) std::memcpy(el.data.data(), elem_data + size_t(e)*h->n_slots*nq, sizeof(double)*h->n_slots*nq);
    if (nom) el.nom = nom[e];
    if (vtss) for (int v = 0; v < n_vert; ++v) el.vertex_data[v] = vtss[size_t(e)*n_vert + v];
    if (el.is_def) {
      int d = e - h->n_car;
      if (ref_normals) std::memcpy(el.jac.data(), ref_normals + size_t(d)*nd*nd*nq, sizeof(double)*nd*nd*nq);
      if (det) std::memcpy(el.jac.data() + size_t(nd)*nd*nq, det + size_t(d)*nq, sizeof(double)*nq);
    }
  }
  if (normals) for (int s = 0; s < h->n_normal_slot; ++s) if (h->normal_store[s]) std::memcpy(h->normal_store[s]->data(), normals + size_t(s)*nd*nfq, sizeof(double)*nd*nfq);
  for (int s = 0; s < h->n_face_slot; ++s) {
    double* f = h->face_store[s]->data();
    if (face_wide) std::memcpy(f, face_wide + size_t(s)*(nd + h->rs)*nfq, sizeof(double)*(nd + h->rs)*nfq);
    if (face_state) std::memcpy(f, face_state + size_t(s)*nv*nfq, sizeof(double)*nv*nfq);
    if (face_ldg) std::memcpy(f + nv*nfq, face_ldg + size_t(s)*nv*nfq, sizeof(double)*nv*nfq);
  }
}

// Element::uncertainty <- array (what Solver::set_uncertainty leaves behind, src/Solver.cpp:672-679)
void hbh_put_uncert(void* handle, const double* uncert)
{
  auto* h = static_cast<Harness*>(handle);
  for (size_t e = 0; e < h->elems.size(); ++e) h->elems[e]->uncertainty = uncert[e];
}

// flat arrays <- host objects
void hbh_fetch(void* handle, double* elem_data, double* face_state, double* face_ldg, double* face_wide, double* uncert)
{
  auto* h = static_cast<Harness*>(handle);
  const int ne = int(h->elems.size()), nq = h->nq, nd = h->nd, nfq = h->nfq, nv = h->nv;
  for (int e = 0; e < ne; ++e) {
    if (elem_data) std::memcpy(elem_data + size_t(e)*h->n_slots*nq, h->elems[e]->data.data(), sizeof(double)*h->n_slots*nq);
    if (uncert) uncert[e] = h->elems[e]->uncertainty;
  }
  for (int s = 0; s < h->n_face_slot; ++s) {
    const double* f = h->face_store[s]->data();
    if (face_state) std::memcpy(face_state + size_t(s)*nv*nfq, f, sizeof(double)*nv*nfq);
    if (face_ldg) std::memcpy(face_ldg + size_t(s)*nv*nfq, f + nv*nfq, sizeof(double)*nv*nfq);
    if (face_wide) std::memcpy(face_wide + size_t(s)*(nd + h->rs)*nfq, f, sizeof(double)*(nd + h->rs)*nfq);
  }
}

/* calls one hexed:: entry point. args[0..3] = dt, i_stage, compute_residual, use_filter; the rest per function. */
int hbh_call(void* handle, int fn, const double* a, double* ret)
{
  auto* h = static_cast<Harness*>(handle);
  try {
    auto m = h->mesh();
    auto o = h->options(a);
    double r = 0.;
    switch (fn) {
      case 0: compute_euler(m, o); break;
      case 1: compute_advection(m, o, a[4]); break;
      case 2: compute_navier_stokes(m, o, h->flux_bc, make_transport(a + 4), make_transport(a + 8)); break;
      case 3: compute_smooth_av(m, o, h->flux_bc, a[4], a[5]); break;
      case 4: compute_fix_therm_admis(m, o, h->flux_bc); break;
      case 5: r = max_dt_euler(m, o, a[4], a[5], a[6] != 0); break;
      case 6: r = max_dt_navier_stokes(m, o, a[4], a[5], a[6] != 0, make_transport(a + 7), make_transport(a + 11)); break;
      case 7: r = max_dt_advection(m, o, a[4], a[5], a[6] != 0, a[7]); break;
      case 8: r = max_dt_smooth_av(m, o, a[4], a[5], a[6] != 0); break;
      case 9: r = max_dt_fix_therm_admis(m, o, a[4], a[5], a[6] != 0); break;
      case 10: compute_prolong(m, a[4] != 0, a[5] != 0); break;
      case 11: compute_restrict(m, a[4] != 0, a[5] != 0); break;
      case 12: compute_prolong_advection(m); break;
      case 13: compute_write_face(m); break;
      case 14: compute_write_face_advection(m); break;
      case 15: compute_write_face_smooth_av(m); break;
      case 16: stabilizing_art_visc(m, a[4]); break;
      default: throw std::runtime_error("unknown function code");
    }
    if (ret) *ret = r;
    return 0;
  } catch (const std::exception& ex) {
    h->error = ex.what();
    return 1;
  }
}

/* adapter controls: what = 0 set_sync_mode(arg), 1 to_host(groups = arg), 2 to_device(groups = arg), 3 boundary_faces_to_host,
 * 4 ghost_faces_to_device, 5 invalidate, 6 release, 7 synchronize */
int hbh_control(void* handle, int what, unsigned arg)
{
  auto* h = static_cast<Harness*>(handle);
  try {
    switch (what) {
      case 0: hexed_b200::set_sync_mode(hexed_b200::Sync_mode(arg)); break;
      case 1: hexed_b200::to_host(h->mesh(), arg); break;
      case 2: hexed_b200::to_device(h->mesh(), arg); break;
      case 3: hexed_b200::boundary_faces_to_host(h->mesh()); break;
      case 4: hexed_b200::ghost_faces_to_device(h->mesh()); break;
      case 5: hexed_b200::invalidate(); break;
      case 6: hexed_b200::release(); break;
      case 7: hexed_b200::synchronize(h->mesh()); break;
      case 8: hexed_b200::boundary_faces_to_host(h->mesh(), arg & 3u, arg >> 2); break; // arg = sides | halves << 2
      case 9: hexed_b200::ghost_faces_to_device(h->mesh(), arg & 3u, arg >> 2); break;
      default: throw std::runtime_error("unknown control code");
    }
    return 0;
  } catch (const std::exception& ex) {
    if (h) h->error = ex.what();
    return 1;
  }
}

/* multi-device controls: hexed_b200::set_devices(devices[n]); optional integer element coordinates [n_elem][3] for the Morton split */
int hbh_set_devices(void* handle, const int* devices, int n)
{
  auto* h = static_cast<Harness*>(handle);
  try {hexed_b200::set_devices(std::vector<int>(devices, devices + n)); return 0;}
  catch (const std::exception& ex) {if (h) h->error = ex.what(); return 1;}
}

int hbh_set_element_coordinates(void* handle, const int* coords)
{
  auto* h = static_cast<Harness*>(handle);
  try {
    std::vector<std::array<int, 3>> c(h->elems.size());
    for (size_t e = 0; e < c.size(); ++e) for (int d = 0; d < 3; ++d) c[e][d] = coords[e*3 + d];
    hexed_b200::set_element_coordinates(h->mesh(), c);
    return 0;
  } catch (const std::exception& ex) {h->error = ex.what(); return 1;}
}

int hbh_element_owners(void* handle, int* out)
{
  auto* h = static_cast<Harness*>(handle);
  try {
    auto o = hexed_b200::element_owners(h->mesh());
    std::copy(o.begin(), o.end(), out);
    return 0;
  } catch (const std::exception& ex) {h->error = ex.what(); return 1;}
}

int hbh_transport_description(void* handle, char* out, int cap)
{
  auto* h = static_cast<Harness*>(handle);
  try {
    std::string d = hexed_b200::transport_description(h->mesh());
    std::snprintf(out, cap, "%s", d.c_str());
    return 0;
  } catch (const std::exception& ex) {h->error = ex.what(); return 1;}
}

/* registers a device boundary condition for the boundary connections whose def_con rows are listed, then the two apply calls */
int hbh_add_device_bc(void* handle, int kind, const int* def_con_index, int n, const double* params, int n_params)
{
  auto* h = static_cast<Harness*>(handle);
  try {
    std::vector<double*> inside;
    for (int i = 0; i < n; ++i) inside.push_back(h->def_cons[def_con_index[i]]->state(0, false));
    hexed_b200::add_device_bc(h->mesh(), kind, inside, std::vector<double>(params, params + n_params));
    return 0;
  } catch (const std::exception& ex) {
    h->error = ex.what();
    return 1;
  }
}

int hbh_set_device_bc_params(void* handle, int id, const double* params, int n_params)
{
  auto* h = static_cast<Harness*>(handle);
  try {
    hexed_b200::set_device_bc_params(h->mesh(), id, std::vector<double>(params, params + n_params));
    return 0;
  } catch (const std::exception& ex) {
    h->error = ex.what();
    return 1;
  }
}

int hbh_apply_bcs(void* handle, int flux)
{
  auto* h = static_cast<Harness*>(handle);
  try {
    if (flux) hexed_b200::apply_flux_bcs(h->mesh()); else hexed_b200::apply_state_bcs(h->mesh());
    return 0;
  } catch (const std::exception& ex) {
    h->error = ex.what();
    return 1;
  }
}

/* hexed_b200::is_admissible: *ok = result, record[n_elem] in Kernel_mesh::elems order */
int hbh_is_admissible(void* handle, int* ok, int* record)
{
  auto* h = static_cast<Harness*>(handle);
  try {
    if (!record) { // the answer alone, as Solver::update needs it after a stage: Element::record is only read when the answer is "no"
      *ok = hexed_b200::is_admissible(h->mesh(), nullptr) ? 1 : 0;
      return 0;
    }
    std::vector<int> rec;
    *ok = hexed_b200::is_admissible(h->mesh(), &rec) ? 1 : 0;
    for (size_t i = 0; i < rec.size(); ++i) record[i] = rec[i];
    return 0;
  } catch (const std::exception& ex) {
    h->error = ex.what();
    return 1;
  }
}

/* the adapter's wrappers of the AV glue (hexed_b200::av_*). what: 0 scale velocity (a = restore), 1 project forcing, 2 finish
 * (a = mult, b = us_max, n = n_real; *out = residual), 3 interp_vertices (n = target, values = per element vertex), 4 swap, 5 apply_aux_bcs (n = mode) */
int hbh_av_glue(void* handle, int what, double a, double b, int n, const double* values, int n_values, double* out)
{
  auto* h = static_cast<Harness*>(handle);
  try {
    switch (what) {
      case 0: hexed_b200::av_scale_velocity(h->mesh(), a != 0.); break;
      case 1: hexed_b200::av_project_forcing(h->mesh()); break;
      case 2: *out = hexed_b200::av_finish(h->mesh(), a, b, n); break;
      case 3: hexed_b200::interp_vertices(h->mesh(), n, std::vector<double>(values, values + n_values)); break;
      case 4: hexed_b200::av_swap(h->mesh()); break;
      case 5: hexed_b200::apply_aux_bcs(h->mesh(), n); break;
      case 6: hexed_b200::av_elwise_ramp(h->mesh(), a); break;
      case 7: hexed_b200::av_elwise_forcing(h->mesh(), n != 0); break;
      case 8: hexed_b200::av_elwise_vertices(h->mesh()); break;
      case 9: { // values = [n_vertex, n_elem*2^nd vertex ids..., matcher rows...] as doubles
        const size_t n_ev = h->elems.size()*size_t(ipow(2, h->nd));
        std::vector<int> ev(n_ev), mt;
        for (size_t i = 0; i < n_ev; ++i) ev[i] = int(values[1 + i]);
        for (size_t i = 1 + n_ev; i < size_t(n_values); ++i) mt.push_back(int(values[i]));
        hexed_b200::vertex_topology(h->mesh(), ev, int(values[0]), mt);
        break;
      }
      default: throw std::runtime_error("unknown glue call");
    }
    return 0;
  } catch (const std::exception& ex) {
    h->error = ex.what();
    return 1;
  }
}

/* the adapter's flattening (device-free), translated back to the harness's slot numbering through the host pointers so the test
 * can compare it entry by entry with the tables the mesh was built from. counts[6] = n_car, n_def, n_face_slot, n_normal_slot,
 * n_boundary, n_null_normal */
int hbh_flatten(void* handle, int* counts, int* car_con, int* def_con, int* ref_face, int* boundary_con)
{
  auto* h = static_cast<Harness*>(handle);
  try {
    auto t = hexed_b200::flatten(h->mesh());
    std::unordered_map<const double*, int> fslot, nslot;
    for (int s = 0; s < h->n_face_slot; ++s) fslot[h->face_store[s]->data()] = s;
    for (int s = 0; s < h->n_normal_slot; ++s) if (h->normal_store[s]) nslot[h->normal_store[s]->data()] = s;
    auto f = [&](int adapter_slot) {return adapter_slot < 0 ? -1 : fslot.at(t.face_ptr[adapter_slot]);};
    for (size_t i = 0; i < t.car_con.size()/3; ++i) {car_con[i*3] = f(t.car_con[i*3]); car_con[i*3 + 1] = f(t.car_con[i*3 + 1]); car_con[i*3 + 2] = t.car_con[i*3 + 2];}
    for (size_t i = 0; i < t.def_con.size()/7; ++i) {
      def_con[i*7] = f(t.def_con[i*7]); def_con[i*7 + 1] = f(t.def_con[i*7 + 1]);
      for (int k = 2; k < 6; ++k) def_con[i*7 + k] = t.def_con[i*7 + k];
      def_con[i*7 + 6] = nslot.at(t.normal_ptr[t.def_con[i*7 + 6]]);
    }
    for (size_t i = 0; i < t.ref_face.size()/7; ++i) {
      for (int k = 0; k < 5; ++k) ref_face[i*7 + k] = f(t.ref_face[i*7 + k]);
      ref_face[i*7 + 5] = t.ref_face[i*7 + 5]; ref_face[i*7 + 6] = t.ref_face[i*7 + 6];
    }
    for (size_t i = 0; i < t.boundary_con.size(); ++i) boundary_con[i] = t.boundary_con[i];
    int n_null = 0;
    for (auto p : t.normal_ptr) n_null += !p;
    // element-owned slots must keep the canonical numbering e*2*n_dim + f
    const int nf = 2*h->nd, ne = t.n_car + t.n_def;
    for (int s = 0; s < nf*ne; ++s) if (t.face_ptr[s] && fslot.at(t.face_ptr[s]) != s) throw std::runtime_error("element face slot renumbered");
    counts[0] = t.n_car; counts[1] = t.n_def; counts[2] = t.n_face_slot; counts[3] = t.n_normal_slot; counts[4] = int(t.boundary_con.size()); counts[5] = n_null;
    return 0;
  } catch (const std::exception& ex) {
    h->error = ex.what();
    return 1;
  }
}

/* work units recorded through the Stopwatch_tree side-contract: out[0..3] = sw_car {neighbor, local, reconcile LDG flux, compute time step},
 * out[4..7] = sw_def likewise, out[8] = sw_pr, out[9] = calls seen by sw_car's own stopwatch */
void hbh_work_units(void* handle, long long* out)
{
  auto* h = static_cast<Harness*>(handle);
  const char* names [4] {"neighbor", "local", "reconcile LDG flux", "compute time step"};
  for (int i = 0; i < 4; ++i) {
    out[i] = h->sw_car.children.at(names[i]).work_units_completed;
    out[4 + i] = h->sw_def.children.at(names[i]).work_units_completed;
  }
  out[8] = h->sw_pr.work_units_completed;
  out[9] = h->sw_car.stopwatch.n_calls();
}

int hbh_face_permutation(int nd, int rs, const int* dir, int restore, double* data, char* err, int err_len)
{
  try {
    Connection_direction d;
    d.i_dim = {dir[0], dir[1]}; d.face_sign = {dir[2] != 0, dir[3] != 0};
    auto p = face_permutation(nd, rs, d, data);
    if (restore) p->restore(); else p->match_faces();
    return 0;
  } catch (const std::exception& ex) {
    std::strncpy(err, ex.what(), err_len - 1); err[err_len - 1] = 0;
    return 1;
  }
}

}

// ---- the C++ partitioner (partition.hpp), exposed table by table so that tests can compare it with hexed_b200/partition.py ----
#include "partition.hpp"

namespace
{
struct Partition_result {std::vector<hexed_b200::Rank_mesh> parts;};
int give(const std::vector<int>& v, int* out) {if (out) std::copy(v.begin(), v.end(), out); return int(v.size());}
}

extern "C" {

void* hbp_partition(int n_dim, int n_car, int n_def, int n_face_slot, int n_normal_slot, const int* car_con, int n_cc,
                    const int* def_con, int n_dc, const int* ref_face, int n_rf, const int* boundary_con, int n_bc,
                    const int* owner, int n_parts)
{
  try {
    hexed_b200::Mesh_graph g;
    g.n_dim = n_dim; g.n_car = n_car; g.n_def = n_def; g.n_face_slot = n_face_slot; g.n_normal_slot = n_normal_slot;
    g.car_con.assign(car_con, car_con + size_t(n_cc)*3); g.def_con.assign(def_con, def_con + size_t(n_dc)*7);
    g.ref_face.assign(ref_face, ref_face + size_t(n_rf)*7); g.boundary_con.assign(boundary_con, boundary_con + n_bc);
    auto* r = new Partition_result;
    r->parts = hexed_b200::partition(g, std::vector<int>(owner, owner + n_car + n_def), n_parts);
    return r;
  } catch (...) {return nullptr;}
}

void hbp_free(void* h) {delete static_cast<Partition_result*>(h);}

/* what: 0 car_con, 1 def_con, 2 ref_face, 3 global_elem, 4 global_face, 5 global_normal, 6 pre_prolong, 7 peers,
 * 8 {n_car, n_def, n_cut_car, n_cut_def, n_face_slot, n_normal_slot}, 9 boundary_con (local def_con rows), 10 face_owned, 11 global_def_con,
 * 100 + k send slots of peer k, 200 + k receive slots of peer k. Returns the number of ints; `out` may be null. */
int hbp_query(void* h, int rank, int what, int* out)
{
  auto& m = static_cast<Partition_result*>(h)->parts.at(rank);
  switch (what) {
    case 0: return give(m.graph.car_con, out);
    case 1: return give(m.graph.def_con, out);
    case 2: return give(m.graph.ref_face, out);
    case 3: return give(m.global_elem, out);
    case 4: return give(m.global_face, out);
    case 5: return give(m.global_normal, out);
    case 6: return give(m.pre_prolong, out);
    case 7: return give(m.peers, out);
    case 8: return give({m.graph.n_car, m.graph.n_def, m.n_cut_car, m.n_cut_def, m.graph.n_face_slot, m.graph.n_normal_slot}, out);
    case 9: return give(m.graph.boundary_con, out);
    case 10: return give(std::vector<int>(m.face_owned.begin(), m.face_owned.end()), out);
    case 11: return give(m.global_def_con, out);
  }
  if (what >= 100 && what < 100 + int(m.peers.size())) return give(m.send_slots[what - 100], out);
  if (what >= 200 && what < 200 + int(m.peers.size())) return give(m.recv_slots[what - 200], out);
  return -1;
}

int hbp_owners_by_graph(int n_dim, int n_car, int n_def, const int* car_con, int n_cc, const int* def_con, int n_dc, const int* ref_face, int n_rf,
                        int n_parts, int* owner_out)
{
  hexed_b200::Mesh_graph g;
  g.n_dim = n_dim; g.n_car = n_car; g.n_def = n_def;
  g.car_con.assign(car_con, car_con + size_t(n_cc)*3); g.def_con.assign(def_con, def_con + size_t(n_dc)*7); g.ref_face.assign(ref_face, ref_face + size_t(n_rf)*7);
  auto o = hexed_b200::owners_by_graph(g, n_parts);
  std::copy(o.begin(), o.end(), owner_out);
  return 0;
}

int hbp_owners_by_morton(int n_dim, int n_car, int n_elem, const int* coords, int n_parts, int* owner_out)
{
  std::vector<std::array<int, 3>> c(n_elem, std::array<int, 3>{0, 0, 0});
  for (int e = 0; e < n_elem; ++e) for (int d = 0; d < n_dim; ++d) c[e][d] = coords[e*n_dim + d];
  auto o = hexed_b200::owners_by_morton(c, n_dim, n_car, n_parts);
  std::copy(o.begin(), o.end(), owner_out);
  return 0;
}

}

// ---- the host boundary-condition loop of Solver::apply_state_bcs (src/Solver.cpp:56-67): an OpenMP loop over the boundary
// connections with one virtual Flow_bc::apply_state call per face, written the way the reference's conditions are (this is what a
// Solver keeps doing on the host between kernel calls when its conditions are not registered on the device). bench.py times it
// inside its end-to-end step; kind: 0 Freestream (src/Boundary_condition.cpp:66-76), 2 Nonpenetration (:301-327). ----
namespace
{
struct Host_bc {virtual ~Host_bc() = default; virtual void apply_state(double* inside, double* ghost, const double* normal, int nd, int nfq) = 0;};
struct Host_freestream : Host_bc
{
  std::vector<double> fs;
  void apply_state(double*, double* ghost, const double*, int nd, int nfq) override
  {
    for (int v = 0; v < nd + 2; ++v) for (int q = 0; q < nfq; ++q) ghost[v*nfq + q] = fs[v];
  }
};
struct Host_nonpenetration : Host_bc
{
  void apply_state(double* inside, double* ghost, const double* n, int nd, int nfq) override
  {
    for (int k = 0; k < (nd + 2)*nfq; ++k) ghost[k] = inside[k];
    for (int q = 0; q < nfq; ++q) {
      double dot = 0., nsq = 0.;
      for (int d = 0; d < nd; ++d) {dot += ghost[d*nfq + q]*n[d*nfq + q]; nsq += n[d*nfq + q]*n[d*nfq + q];}
      for (int d = 0; d < nd; ++d) ghost[d*nfq + q] -= 2*dot*n[d*nfq + q]/nsq;
    }
  }
};
}

extern "C" int hbh_host_state_bcs(void* handle, int kind, const double* params, int n_params, const int* def_con_index, int n, int n_threads)
{
  auto* h = static_cast<Harness*>(handle);
  try {
    std::unique_ptr<Host_bc> bc;
    if (kind == 0) {auto* f = new Host_freestream; f->fs.assign(params, params + n_params); bc.reset(f);}
    else if (kind == 2) bc.reset(new Host_nonpenetration);
    else throw std::runtime_error("host boundary condition kind not provided by the harness");
    const int nd = h->nd, nfq = h->nfq;
    #pragma omp parallel for num_threads(n_threads > 0 ? n_threads : omp_get_max_threads())
    for (int i = 0; i < n; ++i) {
      H_connection& c = *h->def_cons[def_con_index[i]];
      bc->apply_state(c.state(0, false), c.state(1, false), c.normal(), nd, nfq);
    }
    return 0;
  } catch (const std::exception& ex) {h->error = ex.what(); return 1;}
}

// `Freestream::apply_flux` / `Copy::apply_flux` = copy_state of the LDG halves (src/Boundary_condition.cpp:12-23,304-305,455-458): the host
// loop of Solver::apply_flux_bcs (src/Solver.cpp:69-81) for such conditions
extern "C" int hbh_host_flux_bcs(void* handle, const int* def_con_index, int n, int n_threads)
{
  auto* h = static_cast<Harness*>(handle);
  const size_t w = size_t(h->nv)*h->nfq;
  #pragma omp parallel for num_threads(n_threads > 0 ? n_threads : omp_get_max_threads())
  for (int i = 0; i < n; ++i) {
    H_connection& c = *h->def_cons[def_con_index[i]];
    std::memcpy(c.state(1, true), c.state(0, true), w*sizeof(double));
  }
  return 0;
}
