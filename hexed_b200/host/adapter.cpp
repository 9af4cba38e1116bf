/* adapter.cpp -- hexed::compute_euler(Kernel_mesh, Kernel_options) & co. implemented on the B200 C ABI (include/hexed_b200.h).
 *
 * Replaces, in a Hexed build, src/kernels_convective.cpp, src/kernels_diffusive.cpp, src/kernels_max_dt.cpp and
 * src/stabilizing_art_visc.cpp (and nothing else). Per mesh epoch it walks the Sequence<> views once, turns the pointer graph
 * (faces owned by connections and aliased by elements, include/connection.hpp:111-123) into the slot tables the device uses, and
 * from then on only moves data (policy: adapter.hpp) and drives the C ABI. No arithmetic of the hot path happens here: if the
 * library finds no CUDA device every entry point throws.
 */
#include "adapter.hpp"
#include "partition.hpp"
#include "../../include/hexed_b200.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <string>
#include <memory>
#include <thread>
#include <unordered_map>

#ifdef HEXED_B200_WITH_HEXED_HEADERS
#include <Gauss_legendre.hpp>
#endif

namespace hexed_b200
{

namespace
{

using hexed::Kernel_mesh;
using hexed::Kernel_options;

int ipow(int b, int e) {int r = 1; for (int i = 0; i < e; ++i) r *= b; return r;}

// Basis keeps two of the numbers the kernels need protected (include/Basis.hpp:19-22); a derived class may re-export them
struct Basis_access : public hexed::Basis
{
  using hexed::Basis::min_eig_convection;
  using hexed::Basis::quadratic_safety;
};

void legendre_nodes(int row_size, double* out)
{
  #ifdef HEXED_B200_WITH_HEXED_HEADERS
  hexed::Gauss_legendre gl(row_size); // pde::Advection always uses the Legendre nodes (include/pde.hpp:281)
  for (int i = 0; i < row_size; ++i) out[i] = gl.node(i);
  #else
  static const double table [7][8] {
    #include "legendre_nodes.inc"
  };
  for (int i = 0; i < row_size; ++i) out[i] = table[row_size - 2][i];
  #endif
}

//! the packed basis table of hexed_b200_create (layout: include/hexed_b200.h "Basis tables")
std::vector<double> pack_basis(const hexed::Basis& b)
{
  const int rs = b.row_size;
  std::vector<double> p;
  for (int i = 0; i < rs; ++i) p.push_back(b.node(i));
  auto w = b.node_weights();
  for (int i = 0; i < rs; ++i) p.push_back(w(i));
  auto d = b.diff_mat();
  for (int i = 0; i < rs; ++i) for (int j = 0; j < rs; ++j) p.push_back(d(i, j));
  auto bd = b.boundary();
  for (int s = 0; s < 2; ++s) for (int j = 0; j < rs; ++j) p.push_back(bd(s, j));
  for (int deg = 0; deg < rs; ++deg) {
    auto o = b.orthogonal(deg);
    for (int i = 0; i < rs; ++i) p.push_back(o(i));
  }
  auto f = b.filter();
  for (int i = 0; i < rs; ++i) for (int j = 0; j < rs; ++j) p.push_back(f(i, j));
  for (int transfer = 0; transfer < 2; ++transfer) {
    for (int h = 0; h < 2; ++h) {
      bool have = true;
      decltype(b.prolong(0)) m;
      try {m = transfer ? b.restrict(h) : b.prolong(h);}
      catch (...) {have = false;} // Gauss_lobatto does not implement them (include/Gauss_lobatto.hpp)
      for (int i = 0; i < rs; ++i) for (int j = 0; j < rs; ++j) p.push_back(have ? m(i, j) : 0.);
    }
  }
  p.push_back((b.*(&Basis_access::min_eig_convection))());
  p.push_back(b.min_eig_diffusion());
  p.push_back((b.*(&Basis_access::quadratic_safety))());
  std::vector<double> gl(rs);
  legendre_nodes(rs, gl.data());
  p.insert(p.end(), gl.begin(), gl.end());
  return p;
}

/* Transport_model hides its five coefficients (include/Transport_model.hpp:16-27). They are read here by object layout:
 * five doubles in declaration order followed by the `is_viscous` flag. LAYOUT-DEPENDENT -- the static_assert and the
 * coefficient() cross-check below catch a changed class. */
hexed_b200_transport transport(const hexed::Transport_model& t)
{
  static_assert(sizeof(hexed::Transport_model) == 6*sizeof(double), "Transport_model layout changed: update hexed_b200::transport");
  double v [5];
  std::memcpy(v, &t, sizeof(v));
  hexed_b200_transport out {v[0], v[1], v[2], v[3], v[4], t.is_viscous ? 1 : 0};
  const double probe = 17.25; // sqrt(temperature) of the cross-check
  double r = probe/out.sqrt_ref_temp;
  double expect = out.const_val + out.ref_val*r*r*r*(out.ref_temp + out.temp_offset)/(probe*probe + out.temp_offset);
  double got = t.coefficient(probe);
  if (std::abs(expect - got) > 1e-12*std::max(std::abs(got), 1e-300)) throw std::runtime_error("hexed_b200: Transport_model layout is not the expected one");
  return out;
}

//! one device's share of the mesh: a context plus the host addresses behind its element, face and normal slots
struct Rank
{
  hexed_b200_ctx* ctx = nullptr;
  int device = 0;
  int n_car = 0, n_def = 0, n_face_slot = 0, n_normal_slot = 0;
  std::vector<hexed::Kernel_element*> elem; // car then def, local order
  std::vector<int> global_elem;             // local element -> position in Kernel_mesh::elems
  std::vector<double*> face_ptr;            // local face slot -> host storage (halo slots point at the remote element's face)
  std::vector<char> face_owned;             // this rank's copy is the one written back to the host
  std::vector<double*> normal_ptr;
  std::vector<int> def_con;                 // local [n][7]
  std::vector<int> boundary_con;            // local def_con rows that are boundary connections
  int side_list [2] {-1, -1};               // face list ids: inside faces / ghost faces of the boundary connections
  std::vector<int> side_slots [2];
  std::vector<double> staging;
};

struct Mirror
{
  int n_dim = 0, row_size = 0;
  std::vector<Rank> ranks;
  hexed_b200_group* group = nullptr; // more than one rank: halo exchange + allreduce (include/hexed_b200.h "device group")
  Flat_tables tab;                   // the undivided tables
  std::vector<int> owner;            // element -> rank
  bool have_mesh = false;
  bool explicitly_invalidated = false;
  std::vector<const void*> fingerprint;
  std::vector<std::array<int, 3>> coords; // optional integer element coordinates for the Morton split (set_element_coordinates)
  std::vector<double> packed_basis;
  bool device_bcs_ran = false; // set by apply_state_bcs / apply_flux_bcs; lets the flux_bc thunk see what its callback did
  bool prefetch_inside = false; // resident mode, host-applied state BCs: stage drivers start the download of the inside faces themselves
  bool prefetch_valid = false;  // nothing has touched the device faces since the last prefetch was started
  bool prefetch_collected = true; // the last prefetch was picked up by boundary_faces_to_host (if not, the host has stopped asking: stop prefetching)
  hexed_b200_ctx* ctx0() {return ranks.empty() ? nullptr : ranks[0].ctx;}
  void destroy_device_side()
  {
    if (group) hexed_b200_group_destroy(group);
    group = nullptr;
    for (Rank& r : ranks) if (r.ctx) hexed_b200_destroy(r.ctx);
    ranks.clear();
  }
  ~Mirror() {destroy_device_side();}
};

Sync_mode g_mode = sync_every_call;
int g_host_threads = 0; // 0: every hardware thread (NOT omp_get_max_threads(): launchers such as torchrun export OMP_NUM_THREADS=1, and a Hexed
                        // process that drives several GPUs wants its copy loops on all cores)
int host_threads()
{
  if (g_host_threads > 0) return g_host_threads;
  static const int hw = std::max(1u, std::thread::hardware_concurrency());
  return hw;
}
std::vector<int> g_devices {0};
//! keyed by the identity of the mesh (the address of its `elems` view, which belongs to one Solver / Accessible_mesh) and its shape
std::map<std::tuple<const void*, int, int>, std::unique_ptr<Mirror>> g_mirrors;

void check(Mirror* m, int rc, hexed_b200_ctx* ctx = nullptr)
{
  if (!rc) return;
  if (rc == HEXED_B200_INVALID_KERNEL) throw std::runtime_error("demand for invalid kernel"); // include/kernel_factory.hpp:114-116
  if (rc == HEXED_B200_NOT_FINITE) throw std::runtime_error("state is not finite"); // HEXED_ASSERT of src/thermo.cpp:14
  if (!ctx && m) ctx = m->ctx0();
  throw std::runtime_error(std::string("hexed_b200: ") + hexed_b200_last_error(ctx));
}

void check_group(Mirror& m, int rc)
{
  if (!rc) return;
  if (rc == HEXED_B200_INVALID_KERNEL) throw std::runtime_error("demand for invalid kernel");
  if (rc == HEXED_B200_NOT_FINITE) throw std::runtime_error("state is not finite");
  throw std::runtime_error(std::string("hexed_b200: ") + hexed_b200_group_last_error(m.group));
}

/* What identifies a mesh epoch. `thorough` (sync_every_call mode, where a walk over the views is small against the PCIe traffic of the
 * call) folds EVERY element's and connection's storage addresses into a rolling hash, so that a same-size mesh edit that moves any
 * object is noticed; resident mode samples three objects per view and relies on invalidate() after mesh changes (adapter.hpp). */
std::vector<const void*> make_fingerprint(Kernel_mesh& km, bool thorough)
{
  std::vector<const void*> f;
  auto num = [&](size_t n) {f.push_back(reinterpret_cast<const void*>(n));};
  num(km.n_dim); num(km.row_size);
  f.push_back(&km.basis);
  size_t hash = 1469598103934665603ull;
  auto mix = [&](const void* p) {hash = (hash ^ reinterpret_cast<size_t>(p))*1099511628211ull;};
  auto elems = [&](hexed::Sequence<hexed::Kernel_element&>& s) {
    int n = s.size(); num(n);
    for (int i : {0, n/2, n - 1}) if (n) {auto& e = s[i]; f.push_back(e.state()); f.push_back(e.face(0, false));}
    if (thorough) for (int i = 0; i < n; ++i) {auto& e = s[i]; mix(&e); mix(e.state()); mix(e.face(0, false));}
  };
  auto cons = [&](hexed::Sequence<hexed::Kernel_connection&>& s) {
    int n = s.size(); num(n);
    for (int i : {0, n/2, n - 1}) if (n) {auto& c = s[i]; f.push_back(c.state(0, false)); f.push_back(c.state(1, false));}
    if (thorough) for (int i = 0; i < n; ++i) {auto& c = s[i]; mix(c.state(0, false)); mix(c.state(1, false));}
  };
  elems(km.car_elems); elems(km.def_elems); cons(km.car_cons); cons(km.def_cons);
  int nr = km.ref_faces.size(); num(nr);
  for (int i : {0, nr/2, nr - 1}) if (nr) f.push_back(km.ref_faces[i].coarse);
  if (thorough) for (int i = 0; i < nr; ++i) {auto& r = km.ref_faces[i]; mix(r.coarse); for (double* p : r.fine) mix(p);}
  num(thorough ? hash : 0);
  return f;
}

} // namespace

Flat_tables flatten(Kernel_mesh km)
{
  Flat_tables t;
  const int nd = km.n_dim, nf = 2*nd;
  t.n_dim = nd; t.row_size = km.row_size;
  t.n_car = km.car_elems.size(); t.n_def = km.def_elems.size();
  const int ne = t.n_car + t.n_def;
  if (km.elems.size() != ne) throw std::runtime_error("hexed_b200: Kernel_mesh::elems is not car_elems + def_elems");
  t.elem.resize(ne);
  t.face_ptr.assign(size_t(nf)*ne, nullptr);
  t.normal_ptr.assign(size_t(nf)*t.n_def, nullptr);
  std::unordered_map<const double*, int> face_slot, normal_slot;
  face_slot.reserve(size_t(nf)*ne*2);
  for (int e = 0; e < ne; ++e) {
    hexed::Kernel_element& el = e < t.n_car ? km.car_elems[e] : km.def_elems[e - t.n_car];
    if (el.deformed() != (e >= t.n_car)) throw std::runtime_error("hexed_b200: element in the wrong Kernel_mesh view");
    t.elem[e] = &el;
    for (int f = 0; f < nf; ++f) {
      double* p = el.face(f, false);
      if (!p) continue;
      t.face_ptr[size_t(e)*nf + f] = p;
      face_slot[p] = e*nf + f;
      if (e >= t.n_car) {
        double* n = el.kernel_face_normal(f);
        int s = (e - t.n_car)*nf + f;
        t.normal_ptr[s] = n;
        if (n) normal_slot[n] = s;
      }
    }
  }
  auto face_of = [&](double* p) -> int { // faces that no element aliases (ghosts, mortar faces) get the slots after the elements'
    auto it = face_slot.find(p);
    if (it != face_slot.end()) return it->second;
    int s = int(t.face_ptr.size());
    t.face_ptr.push_back(p); face_slot[p] = s;
    return s;
  };
  int n_cc = km.car_cons.size();
  t.car_con.resize(size_t(n_cc)*3);
  for (int i = 0; i < n_cc; ++i) {
    auto& c = km.car_cons[i];
    auto dir = c.get_direction();
    t.car_con[size_t(i)*3] = face_of(c.state(0, false));
    t.car_con[size_t(i)*3 + 1] = face_of(c.state(1, false));
    t.car_con[size_t(i)*3 + 2] = dir.i_dim[0];
  }
  int n_dc = km.def_cons.size();
  t.def_con.resize(size_t(n_dc)*7);
  std::vector<int> fresh_side1; // connections whose side 1 is not an element face: boundary or coarse-side mortar
  for (int i = 0; i < n_dc; ++i) {
    auto& c = km.def_cons[i];
    auto dir = c.get_direction();
    int* row = t.def_con.data() + size_t(i)*7;
    row[0] = face_of(c.state(0, false));
    int before = int(t.face_ptr.size());
    row[1] = face_of(c.state(1, false));
    if (row[1] >= before && row[0] < nf*ne) fresh_side1.push_back(i);
    row[2] = dir.i_dim[0]; row[3] = dir.i_dim[1]; row[4] = dir.face_sign[0]; row[5] = dir.face_sign[1];
    double* n = c.normal();
    if (!n) throw std::runtime_error("hexed_b200: deformed connection without a normal");
    auto it = normal_slot.find(n);
    if (it != normal_slot.end()) row[6] = it->second;
    else {
      row[6] = int(t.normal_ptr.size());
      t.normal_ptr.push_back(n); normal_slot[n] = row[6];
    }
  }
  int n_rf = km.ref_faces.size();
  t.ref_face.assign(size_t(n_rf)*7, -1);
  for (int i = 0; i < n_rf; ++i) {
    auto& r = km.ref_faces[i];
    int* row = t.ref_face.data() + size_t(i)*7;
    row[0] = face_of(r.coarse);
    int n_fine = ipow(2, nd - 1);
    for (int k = 0; k < nd - 1; ++k) if (r.stretch[k]) n_fine /= 2;
    for (int k = 0; k < n_fine; ++k) row[1 + k] = face_of(r.fine[k]);
    row[5] = r.stretch[0]; row[6] = nd > 2 ? int(r.stretch[1]) : 0;
  }
  // a connection whose side 1 belongs to no element is a boundary connection unless that face is a mortar face of a refined face
  std::vector<char> is_mortar(t.face_ptr.size(), 0);
  for (int i = 0; i < n_rf; ++i) for (int k = 1; k < 5; ++k) if (t.ref_face[size_t(i)*7 + k] >= 0) is_mortar[t.ref_face[size_t(i)*7 + k]] = 1;
  for (int i : fresh_side1) if (!is_mortar[t.def_con[size_t(i)*7 + 1]]) t.boundary_con.push_back(i);
  t.n_face_slot = int(t.face_ptr.size());
  t.n_normal_slot = int(t.normal_ptr.size());
  return t;
}

namespace
{

// ---- data movement between the host objects and the device mirror ----

struct Slot_range {int first, n;};

std::vector<Slot_range> slot_ranges(const Mirror& m, unsigned groups)
{
  const int nv = m.n_dim + 2, rs = m.row_size;
  std::vector<Slot_range> r; // element slots in reference order (src/Element.cpp:114-142,187-189)
  if (groups & state) r.push_back({0, nv});
  if (groups & tss) r.push_back({nv, 1});
  if (groups & art_visc) r.push_back({nv + 1, 6});
  if (groups & advection) r.push_back({nv + 7, rs});
  if (groups & res_cache) r.push_back({nv + 7 + rs, std::max(nv, rs)});
  std::vector<Slot_range> merged; // adjacent groups travel together
  for (auto s : r) {
    if (!merged.empty() && merged.back().first + merged.back().n == s.first) merged.back().n += s.n;
    else merged.push_back(s);
  }
  return merged;
}

void move_elements(Mirror& m, Rank& k, unsigned groups, bool up)
{
  const int nq = ipow(m.row_size, m.n_dim);
  const int ne = int(k.elem.size());
  const int chunk = 8192;
  for (auto range : slot_ranges(m, groups)) {
    const size_t per_elem = size_t(range.n)*nq;
    k.staging.resize(per_elem*std::min(chunk, std::max(ne, 1)));
    for (int first = 0; first < ne; first += chunk) {
      const int n = std::min(chunk, ne - first);
      if (up) {
        #pragma omp parallel for num_threads(host_threads())
        for (int i = 0; i < n; ++i) std::memcpy(k.staging.data() + per_elem*i, k.elem[first + i]->state() + size_t(range.first)*nq, per_elem*sizeof(double));
        check(&m, hexed_b200_upload_elem_slots(k.ctx, k.staging.data(), per_elem, range.first, range.n, first, n), k.ctx);
      } else {
        check(&m, hexed_b200_download_elem_slots(k.ctx, k.staging.data(), per_elem, range.first, range.n, first, n), k.ctx);
        #pragma omp parallel for num_threads(host_threads())
        for (int i = 0; i < n; ++i) std::memcpy(k.elem[first + i]->state() + size_t(range.first)*nq, k.staging.data() + per_elem*i, per_elem*sizeof(double));
      }
    }
  }
}

//! one face array (`which` = HEXED_B200_FACE_STATE / _LDG / _WIDE) for all slots; `offset` doubles into each host face.
//! Downloads write only the copies this rank owns (halo slots and replicated connection faces belong to another rank).
void move_face_array(Mirror& m, Rank& k, int which, size_t width, size_t offset, bool up)
{
  const int ns = k.n_face_slot;
  const int chunk = 1 << 16;
  k.staging.resize(width*std::min(chunk, std::max(ns, 1)));
  for (int first = 0; first < ns; first += chunk) {
    const int n = std::min(chunk, ns - first);
    if (up) {
      #pragma omp parallel for num_threads(host_threads())
      for (int i = 0; i < n; ++i) {
        double* p = k.face_ptr[first + i];
        if (p) std::memcpy(k.staging.data() + width*i, p + offset, width*sizeof(double));
        else std::memset(k.staging.data() + width*i, 0, width*sizeof(double));
      }
      check(&m, hexed_b200_upload(k.ctx, which, k.staging.data(), first, n), k.ctx);
    } else {
      check(&m, hexed_b200_download(k.ctx, which, k.staging.data(), first, n), k.ctx);
      #pragma omp parallel for num_threads(host_threads())
      for (int i = 0; i < n; ++i) {
        double* p = k.face_ptr[first + i];
        if (p && k.face_owned[first + i]) std::memcpy(p + offset, k.staging.data() + width*i, width*sizeof(double));
      }
    }
  }
}

void move_faces(Mirror& m, Rank& k, unsigned groups, bool up)
{
  const int nfq = ipow(m.row_size, m.n_dim - 1), nv = m.n_dim + 2;
  if (groups & faces) { // `face(i, is_ldg)` = storage + is_ldg*n_var*nfq (src/Element.cpp:189, include/connection.hpp:66)
    move_face_array(m, k, HEXED_B200_FACE_STATE, size_t(nv)*nfq, 0, up);
    move_face_array(m, k, HEXED_B200_FACE_LDG, size_t(nv)*nfq, size_t(nv)*nfq, up);
  }
  if (groups & faces_wide) move_face_array(m, k, HEXED_B200_FACE_WIDE, size_t(m.n_dim + m.row_size)*nfq, 0, up);
}

void upload_vertex_tss(Mirror& m, Rank& k)
{ // vertex_time_step_scale is rewritten by Solver::set_local_tss between calls; it is tiny
  const int n_vert = ipow(2, m.n_dim), ne = int(k.elem.size());
  std::vector<double> buf(size_t(ne)*n_vert);
  for (int e = 0; e < ne; ++e) for (int v = 0; v < n_vert; ++v) buf[size_t(e)*n_vert + v] = k.elem[e]->vertex_time_step_scale(v);
  check(&m, hexed_b200_upload(k.ctx, HEXED_B200_VERTEX_TSS, buf.data(), 0, ne), k.ctx);
}

void upload_geometry(Mirror& m, Rank& k)
{
  const int nd = m.n_dim, nq = ipow(m.row_size, nd), nfq = nq/m.row_size, nf = 2*nd;
  const int ne = int(k.elem.size()), n_def = k.n_def, n_car = k.n_car;
  std::vector<double> buf(ne);
  for (int e = 0; e < ne; ++e) buf[e] = k.elem[e]->nominal_size();
  check(&m, hexed_b200_upload(k.ctx, HEXED_B200_NOMINAL_SIZE, buf.data(), 0, ne), k.ctx);
  upload_vertex_tss(m, k);
  const int chunk = 8192;
  for (int first = 0; first < n_def; first += chunk) {
    const int n = std::min(chunk, n_def - first);
    buf.resize(size_t(n)*nd*nd*nq);
    #pragma omp parallel for num_threads(host_threads())
    for (int i = 0; i < n; ++i) std::memcpy(buf.data() + size_t(i)*nd*nd*nq, k.elem[n_car + first + i]->reference_level_normals(), sizeof(double)*nd*nd*nq);
    check(&m, hexed_b200_upload(k.ctx, HEXED_B200_REF_NORMALS, buf.data(), first, n), k.ctx);
    #pragma omp parallel for num_threads(host_threads())
    for (int i = 0; i < n; ++i) std::memcpy(buf.data() + size_t(i)*nq, k.elem[n_car + first + i]->jacobian_determinant(), sizeof(double)*nq);
    check(&m, hexed_b200_upload(k.ctx, HEXED_B200_JAC_DET, buf.data(), first, n), k.ctx);
  }
  const int nn = k.n_normal_slot;
  buf.assign(size_t(nn)*nd*nfq, 0.);
  for (int s = 0; s < nn; ++s) {
    double* dst = buf.data() + size_t(s)*nd*nfq;
    if (k.normal_ptr[s]) std::memcpy(dst, k.normal_ptr[s], sizeof(double)*nd*nfq);
    else { // deformed element face in a Cartesian connection: unit normal of the face's dimension (include/Spatial.hpp:331-339,366)
      const int i_dim = (s % nf)/2;
      for (int q = 0; q < nfq; ++q) dst[size_t(i_dim)*nfq + q] = 1.;
    }
  }
  if (nn) check(&m, hexed_b200_upload(k.ctx, HEXED_B200_NORMALS, buf.data(), 0, nn), k.ctx);
}

void download_uncert(Mirror& m, Rank& k)
{
  const int ne = int(k.elem.size());
  std::vector<double> buf(ne);
  check(&m, hexed_b200_download(k.ctx, HEXED_B200_UNCERT, buf.data(), 0, ne), k.ctx);
  for (int e = 0; e < ne; ++e) k.elem[e]->uncert() = buf[e];
}

void move(Mirror& m, unsigned groups, bool up)
{
  for (Rank& k : m.ranks) {
    move_elements(m, k, groups, up);
    move_faces(m, k, groups, up);
    if (up && (groups & geometry)) upload_geometry(m, k);
    if (!up && (groups & uncert)) download_uncert(m, k);
  }
}

/* boundary faces between the host objects and the device. side 0 = inside faces, 1 = ghost faces; half 0 = state, 1 = LDG.
 * The two combinations a state boundary-condition loop needs, (inside, state, down) and (ghost, state, up), take the asynchronous
 * route of include/hexed_b200.h "asynchronous variants" in resident mode. */
void move_boundary(Mirror& m, bool up, unsigned sides = both_sides, unsigned halves = both_halves)
{
  const int nfq = ipow(m.row_size, m.n_dim - 1), nv = m.n_dim + 2;
  const size_t width = size_t(nv)*nfq;
  for (Rank& k : m.ranks) for (int side = 0; side < 2; ++side) {
    if (!(sides & (1u << side)) || k.side_slots[side].empty()) continue;
    const std::vector<int>& slots = k.side_slots[side];
    const size_t n = slots.size();
    for (int half = 0; half < 2; ++half) {
      if (!(halves & (1u << half))) continue;
      const bool fast = g_mode == resident && half == 0 && ((!up && side == 0) || (up && side == 1));
      if (up) {
        double* buf;
        if (fast) check(&m, hexed_b200_face_list_staging(k.ctx, k.side_list[side], &buf), k.ctx);
        else {k.staging.resize(width*n); buf = k.staging.data();}
        #pragma omp parallel for num_threads(host_threads())
        for (size_t i = 0; i < n; ++i) std::memcpy(buf + width*i, k.face_ptr[slots[i]] + half*width, width*sizeof(double));
        if (fast) check(&m, hexed_b200_face_list_upload_deferred(k.ctx, k.side_list[side], half), k.ctx);
        else check(&m, hexed_b200_face_list_upload(k.ctx, k.side_list[side], half, buf), k.ctx);
      } else {
        const double* buf;
        if (fast) {
          if (!m.prefetch_valid) check(&m, hexed_b200_face_list_prefetch(k.ctx, k.side_list[side], half), k.ctx); // what is in flight (if anything) is stale
          check(&m, hexed_b200_face_list_prefetched(k.ctx, k.side_list[side], half, &buf), k.ctx);
        }
        else {
          k.staging.resize(width*n);
          check(&m, hexed_b200_face_list_download(k.ctx, k.side_list[side], half, k.staging.data()), k.ctx);
          buf = k.staging.data();
        }
        #pragma omp parallel for num_threads(host_threads())
        for (size_t i = 0; i < n; ++i) std::memcpy(k.face_ptr[slots[i]] + half*width, buf + width*i, width*sizeof(double));
      }
    }
  }
  if (!up && g_mode == resident && (sides & inside) && (halves & state_half)) { m.prefetch_inside = true; m.prefetch_collected = true; }
}

//! resident mode with host-applied state BCs: start the download of the inside boundary faces the stage has just produced
void prefetch_boundary(Mirror& m)
{
  if (!m.prefetch_inside || g_mode != resident) return;
  if (!m.prefetch_collected) { // a whole stage went by without the host collecting the faces (e.g. the conditions moved to the device): no more copies
    m.prefetch_inside = false;
    return;
  }
  for (Rank& k : m.ranks) if (!k.side_slots[0].empty()) check(&m, hexed_b200_face_list_prefetch(k.ctx, k.side_list[0], 0), k.ctx);
  m.prefetch_valid = true;
  m.prefetch_collected = false;
}

Mesh_graph graph_of(const Flat_tables& t)
{
  Mesh_graph g;
  g.n_dim = t.n_dim; g.n_car = t.n_car; g.n_def = t.n_def; g.n_face_slot = t.n_face_slot; g.n_normal_slot = t.n_normal_slot;
  g.car_con = t.car_con; g.def_con = t.def_con; g.ref_face = t.ref_face; g.boundary_con = t.boundary_con;
  return g;
}

//! (re)builds the device side of a mirror from its freshly flattened tables: one rank per device, the group when there are several
void build_device_side(Mirror& m)
{
  const int n_ranks = int(g_devices.size());
  if (int(m.ranks.size()) != n_ranks) {
    m.destroy_device_side();
    m.ranks.resize(n_ranks);
    for (int r = 0; r < n_ranks; ++r) {
      m.ranks[r].device = g_devices[r];
      check(nullptr, hexed_b200_create(&m.ranks[r].ctx, g_devices[r], m.n_dim, m.row_size, m.packed_basis.data(), int(m.packed_basis.size())));
    }
    if (n_ranks > 1) {
      std::vector<hexed_b200_ctx*> ctxs;
      for (Rank& k : m.ranks) ctxs.push_back(k.ctx);
      int rc = hexed_b200_group_create(&m.group, n_ranks, ctxs.data());
      if (rc) throw std::runtime_error(std::string("hexed_b200: ") + hexed_b200_group_last_error(nullptr));
    }
  }
  const Flat_tables& t = m.tab;
  std::vector<Rank_mesh> parts;
  if (n_ranks == 1) { // the undivided mesh is the one rank's mesh
    Rank_mesh whole;
    whole.graph = graph_of(t);
    parts.push_back(std::move(whole));
    m.owner.assign(t.elem.size(), 0);
  } else {
    Mesh_graph g = graph_of(t);
    if (m.coords.size() == t.elem.size()) m.owner = owners_by_morton(m.coords, t.n_dim, t.n_car, n_ranks);
    else m.owner = owners_by_graph(g, n_ranks);
    parts = partition(g, m.owner, n_ranks);
  }
  for (int r = 0; r < n_ranks; ++r) {
    Rank& k = m.ranks[r];
    Rank_mesh& pm = parts[r];
    const Mesh_graph& lg = pm.graph;
    k.n_car = lg.n_car; k.n_def = lg.n_def; k.n_face_slot = lg.n_face_slot; k.n_normal_slot = lg.n_normal_slot;
    if (n_ranks == 1) {
      k.elem = t.elem; k.face_ptr = t.face_ptr; k.normal_ptr = t.normal_ptr;
      k.global_elem.resize(t.elem.size());
      for (size_t i = 0; i < t.elem.size(); ++i) k.global_elem[i] = int(i);
      k.face_owned.assign(t.face_ptr.size(), 1);
    } else {
      k.global_elem = pm.global_elem;
      k.elem.resize(pm.global_elem.size());
      for (size_t i = 0; i < k.elem.size(); ++i) k.elem[i] = t.elem[pm.global_elem[i]];
      k.face_ptr.resize(pm.global_face.size());
      for (size_t i = 0; i < k.face_ptr.size(); ++i) k.face_ptr[i] = t.face_ptr[pm.global_face[i]];
      k.normal_ptr.resize(pm.global_normal.size());
      for (size_t i = 0; i < k.normal_ptr.size(); ++i) k.normal_ptr[i] = t.normal_ptr[pm.global_normal[i]];
      k.face_owned = pm.face_owned;
    }
    // local def_con order: [other interior connections | boundary connections | cut connections]. Boundary connections count as
    // "late" like the cut ones (hexed_b200_set_partition), so that ghost faces uploaded by a host boundary-condition loop may still be
    // on their way while the Neighbor kernels of the other connections run.
    const size_t n_dc = lg.def_con.size()/7;
    const size_t n_interior = n_dc - size_t(pm.n_cut_def);
    std::vector<char> is_bnd(n_dc, 0);
    for (int i : lg.boundary_con) is_bnd[i] = 1;
    std::vector<size_t> order;
    for (size_t i = 0; i < n_interior; ++i) if (!is_bnd[i]) order.push_back(i);
    const size_t first_bnd = order.size();
    for (size_t i = 0; i < n_interior; ++i) if (is_bnd[i]) order.push_back(i);
    const int n_bnd = int(order.size() - first_bnd);
    for (size_t i = n_interior; i < n_dc; ++i) order.push_back(i);
    k.def_con.resize(lg.def_con.size());
    for (size_t i = 0; i < n_dc; ++i) std::copy(lg.def_con.begin() + order[i]*7, lg.def_con.begin() + order[i]*7 + 7, k.def_con.begin() + i*7);
    k.boundary_con.resize(n_bnd);
    for (int i = 0; i < n_bnd; ++i) k.boundary_con[i] = int(first_bnd) + i;
    hexed_b200_mesh_desc d {};
    d.n_car = lg.n_car; d.n_def = lg.n_def; d.n_face_slot = lg.n_face_slot; d.n_normal_slot = lg.n_normal_slot;
    d.n_car_con = int(lg.car_con.size()/3); d.n_def_con = int(n_dc); d.n_ref = int(lg.ref_face.size()/7);
    d.car_con = lg.car_con.data(); d.def_con = k.def_con.data(); d.ref_face = lg.ref_face.data();
    check(&m, hexed_b200_mesh_create(k.ctx, &d), k.ctx);
    check(&m, hexed_b200_set_partition(k.ctx, pm.n_cut_car, pm.n_cut_def + n_bnd, int(pm.pre_prolong.size()), pm.pre_prolong.data()), k.ctx);
    if (n_ranks > 1) {
      std::vector<int> n_send, n_recv, send, recv;
      for (size_t i = 0; i < pm.peers.size(); ++i) {
        n_send.push_back(int(pm.send_slots[i].size())); n_recv.push_back(int(pm.recv_slots[i].size()));
        send.insert(send.end(), pm.send_slots[i].begin(), pm.send_slots[i].end());
        recv.insert(recv.end(), pm.recv_slots[i].begin(), pm.recv_slots[i].end());
      }
      check_group(m, hexed_b200_group_set_halo(m.group, r, int(pm.peers.size()), pm.peers.data(), n_send.data(), send.data(), n_recv.data(), recv.data()));
    }
    for (int side = 0; side < 2; ++side) {
      k.side_slots[side].clear();
      for (int i : k.boundary_con) k.side_slots[side].push_back(k.def_con[size_t(i)*7 + side]);
      k.side_list[side] = -1;
      if (n_bnd) check(&m, hexed_b200_face_list_create(k.ctx, k.side_slots[side].data(), n_bnd, &k.side_list[side]), k.ctx);
    }
    upload_geometry(m, k);
  }
}

//! finds (or builds) the device mirror of `km`; a new mesh epoch uploads geometry and, in resident mode, all data
Mirror& mirror(Kernel_mesh& km)
{
  auto key = std::make_tuple(static_cast<const void*>(&km.elems), km.n_dim, km.row_size);
  auto& slot = g_mirrors[key];
  if (!slot) {
    std::unique_ptr<Mirror> m(new Mirror);
    m->n_dim = km.n_dim; m->row_size = km.row_size;
    if (km.n_dim >= 1 && km.n_dim <= 3 && km.row_size >= 2 && km.row_size <= 8) m->packed_basis = pack_basis(km.basis);
    else m->packed_basis.assign(1, 0.); // the library answers "demand for invalid kernel" before it looks at the table
    slot = std::move(m);
  }
  Mirror& m = *slot;
  auto fp = make_fingerprint(km, g_mode == sync_every_call);
  if (m.have_mesh && fp == m.fingerprint) return m;
  if (m.have_mesh && g_mode == resident && !m.explicitly_invalidated) {
    // the device copy is authoritative in resident mode: re-flattening would upload host data over results only the device holds
    throw std::runtime_error("hexed_b200: the mesh changed while its state is resident on the device; call to_host() before changing the "
                             "mesh and invalidate() after (adapter.hpp)");
  }
  m.tab = flatten(km);
  build_device_side(m);
  m.fingerprint = fp;
  m.have_mesh = true;
  m.explicitly_invalidated = false;
  if (g_mode == resident) move(m, all_elem | faces | faces_wide, true);
  return m;
}

hexed_b200_options options(const Kernel_options& o) {return {o.dt, o.i_stage, o.compute_residual, o.use_filter};}

//! RAII for one entry point: sync-in on construction, sync-out in finish() (not in the destructor: downloads can throw)
struct Call
{
  Mirror& m;
  unsigned out;
  Call(Kernel_mesh& km, unsigned in, unsigned out_groups) : m{mirror(km)}, out{out_groups}
  {
    m.prefetch_valid = false; // this entry point may write faces: a download started earlier no longer describes them
    if (g_mode == sync_every_call) {
      // kernels write their outputs only partly (e.g. write_face leaves ghost and mortar faces alone), so everything that will be
      // downloaded must first hold the host's values
      in |= out & ~unsigned(uncert);
      // "always correct with no Solver change" includes the metric terms: Solver::calc_jacobian, vertex relaxation and snap_faces
      // rewrite normals, determinants and nominal data IN PLACE (same pointers, same sizes), so this mode re-reads them with the state
      if (in & (state | tss)) in |= geometry;
      move(m, in, true);
    }
  }
  void finish() {if (g_mode == sync_every_call) move(m, out, false);}
  //! `f(ctx)` on every rank
  template <typename F> void each(F f) {for (Rank& k : m.ranks) check(&m, f(k.ctx), k.ctx);}
  bool multi() const {return m.group != nullptr;}
};

void single_device_only(Mirror& m, const char* what)
{
  if (m.group) throw std::runtime_error(std::string("hexed_b200: ") + what + " is not available on more than one device (set_devices)");
}

// the Stopwatch_tree side-contract (include/kernel_factory.hpp:32-45): every kernel call adds sequence.size() work units to its
// child. Device work of one stage is a single asynchronous C call, so its host time is charged to the category stopwatches
// as a whole; per-kernel device time is available from hexed_b200_kernel_stats.
struct Watch
{
  std::vector<std::unique_ptr<hexed::Stopwatch::Operator>> running;
  void start(hexed::Stopwatch_tree& t) {if (!t.stopwatch.running()) running.emplace_back(new hexed::Stopwatch::Operator(t.stopwatch));}
};

void count(hexed::Stopwatch_tree& category, const char* name, int units) {category.children.at(name).work_units_completed += units;}

void count_convective(Kernel_mesh& km, Kernel_options& o)
{
  count(o.sw_car, "neighbor", km.car_cons.size()); count(o.sw_def, "neighbor", km.def_cons.size());
  o.sw_pr.work_units_completed += km.ref_faces.size();
  count(o.sw_car, "local", km.car_elems.size()); count(o.sw_def, "local", km.def_elems.size());
  o.sw_pr.work_units_completed += km.ref_faces.size();
}

void count_diffusive(Kernel_mesh& km, Kernel_options& o)
{ // src/kernels_diffusive.cpp:8-26
  count(o.sw_car, "neighbor", km.car_cons.size()); count(o.sw_def, "neighbor", km.def_cons.size());
  o.sw_pr.work_units_completed += 2*km.ref_faces.size();
  count(o.sw_car, "local", km.car_elems.size()); count(o.sw_def, "local", km.def_elems.size());
  if (!o.i_stage) {
    o.sw_pr.work_units_completed += km.ref_faces.size();
    count(o.sw_car, "neighbor", km.car_cons.size()); count(o.sw_def, "neighbor", km.def_cons.size());
    o.sw_pr.work_units_completed += km.ref_faces.size();
    count(o.sw_car, "reconcile LDG flux", km.car_elems.size()); count(o.sw_def, "reconcile LDG flux", km.def_elems.size());
  }
  o.sw_pr.work_units_completed += km.ref_faces.size();
}

struct Flux_bc_thunk
{
  Mirror* m;
  std::function<void()>* fun;
  static void call(void* user)
  {
    auto* self = static_cast<Flux_bc_thunk*>(user);
    if (!*self->fun) return;
    // the host callback reads and writes boundary faces (Solver::apply_flux_bcs, src/Solver.cpp:69-81): sync_every_call mode brackets it with
    // the face traffic. In resident mode the callback owns that traffic, exactly like the patched apply_state_bcs:
    // a host loop calls boundary_faces_to_host(mesh, inside, ldg_half) / ghost_faces_to_device(mesh, ghost, ldg_half) around itself, a
    // callback that applies its conditions on the device (hexed_b200::apply_flux_bcs) moves nothing -- an unconditional download here cost
    // that path 141 MB of PCIe traffic and a host pass over every boundary object per step (26.1 ms against 14.4 ms of device time at 262 k
    // elements, visit r02y)
    if (g_mode == sync_every_call) {for (Rank& k : self->m->ranks) move_faces(*self->m, k, faces, false);}
    self->m->device_bcs_ran = false;
    (*self->fun)();
    if (g_mode == sync_every_call) {for (Rank& k : self->m->ranks) move_faces(*self->m, k, faces, true);}
  }
};

//! the five max_dt_* entry points: pde 0 Euler, 1 Navier-Stokes, 2 advection, 3 smooth AV, 4 fix therm admis
double max_dt_call(Kernel_mesh& km, Kernel_options& o, unsigned in, int pde, double sc, double sd, bool local_time,
                   hexed_b200_transport visc, hexed_b200_transport cond, double advect_length)
{
  Call call(km, in | tss, tss);
  Watch w; w.start(o.sw_car); w.start(o.sw_def);
  double dt = 0.;
  if (call.multi()) check_group(call.m, hexed_b200_group_max_dt(call.m.group, pde, sc, sd, local_time, visc, cond, advect_length, &dt));
  else {
    hexed_b200_ctx* c = call.m.ctx0();
    int rc = 0;
    switch (pde) {
      case 0: rc = hexed_b200_max_dt_euler(c, options(o), sc, sd, local_time, &dt); break;
      case 1: rc = hexed_b200_max_dt_navier_stokes(c, options(o), sc, sd, local_time, visc, cond, &dt); break;
      case 2: rc = hexed_b200_max_dt_advection(c, options(o), sc, sd, local_time, advect_length, &dt); break;
      case 3: rc = hexed_b200_max_dt_smooth_av(c, options(o), sc, sd, local_time, &dt); break;
      default: rc = hexed_b200_max_dt_fix_therm_admis(c, options(o), sc, sd, local_time, &dt); break;
    }
    check(&call.m, rc);
  }
  count(o.sw_car, "compute time step", km.car_elems.size()); count(o.sw_def, "compute time step", km.def_elems.size());
  call.finish();
  return dt;
}

const hexed_b200_transport no_transport {0., 0., 1., 1., 1., 0};

class Face_permutation_host : public hexed::Face_permutation_dynamic
{
  std::vector<int> table; // matched[p] = original[table[p]]
  int n_var;
  double* data;
  std::vector<double> tmp;
  public:
  Face_permutation_host(int n_dim, int row_size, hexed::Connection_direction dir, double* d) : n_var{n_dim + 2}, data{d}
  {
    table.resize(ipow(row_size, n_dim - 1));
    int dr [4] {dir.i_dim[0], dir.i_dim[1], dir.face_sign[0], dir.face_sign[1]};
    int rc = hexed_b200_face_permutation_indices(n_dim, row_size, dr, table.data());
    if (rc == HEXED_B200_INVALID_KERNEL) throw std::runtime_error("demand for invalid kernel");
    if (rc) throw std::runtime_error("hexed_b200: bad connection direction");
    tmp.resize(table.size());
  }
  void match_faces() override
  {
    const size_t n = table.size();
    for (int v = 0; v < n_var; ++v) {
      for (size_t p = 0; p < n; ++p) tmp[p] = data[v*n + table[p]];
      std::memcpy(data + v*n, tmp.data(), n*sizeof(double));
    }
  }
  void restore() override
  {
    const size_t n = table.size();
    for (int v = 0; v < n_var; ++v) {
      for (size_t p = 0; p < n; ++p) tmp[table[p]] = data[v*n + p];
      std::memcpy(data + v*n, tmp.data(), n*sizeof(double));
    }
  }
};

} // namespace

void set_sync_mode(Sync_mode mode) {g_mode = mode;}
void set_host_threads(int n) {g_host_threads = n;}
Sync_mode sync_mode() {return g_mode;}
void set_device(int d) {set_devices({d});}
void set_devices(const std::vector<int>& devices)
{
  if (devices.empty()) throw std::runtime_error("hexed_b200: set_devices needs at least one device");
  if (devices == g_devices) return;
  g_devices = devices;
  for (auto& kv : g_mirrors) if (kv.second) { // existing mirrors are rebuilt on the new device set at their next call
    if (kv.second->have_mesh && g_mode == resident) throw std::runtime_error("hexed_b200: set_devices while a mesh is resident; call to_host() and release() first");
    kv.second->destroy_device_side();
    kv.second->have_mesh = false;
  }
}
int n_devices() {return int(g_devices.size());}
void set_element_coordinates(Kernel_mesh km, const std::vector<std::array<int, 3>>& coords)
{
  auto key = std::make_tuple(static_cast<const void*>(&km.elems), km.n_dim, km.row_size);
  auto& slot = g_mirrors[key];
  if (!slot) {
    slot.reset(new Mirror);
    slot->n_dim = km.n_dim; slot->row_size = km.row_size;
    if (km.n_dim >= 1 && km.n_dim <= 3 && km.row_size >= 2 && km.row_size <= 8) slot->packed_basis = pack_basis(km.basis);
    else slot->packed_basis.assign(1, 0.);
  }
  if (int(coords.size()) != km.elems.size()) throw std::runtime_error("hexed_b200: set_element_coordinates: one coordinate triple per element of Kernel_mesh::elems");
  slot->coords = coords;
  if (slot->have_mesh && g_devices.size() > 1) {
    if (g_mode == resident) throw std::runtime_error("hexed_b200: set_element_coordinates while a mesh is resident");
    slot->have_mesh = false;
  }
}
std::vector<int> element_owners(Kernel_mesh km) {return mirror(km).owner;}
void invalidate() {for (auto& kv : g_mirrors) if (kv.second) {kv.second->have_mesh = false; kv.second->explicitly_invalidated = true;}}
void release() {g_mirrors.clear();}
void to_host(Kernel_mesh km, unsigned groups) {move(mirror(km), groups & ~unsigned(geometry), false);}
void to_device(Kernel_mesh km, unsigned groups) {Mirror& m = mirror(km); m.prefetch_valid = false; move(m, groups & ~unsigned(uncert), true);}
void boundary_faces_to_host(Kernel_mesh km) {move_boundary(mirror(km), false);}
void ghost_faces_to_device(Kernel_mesh km) {move_boundary(mirror(km), true);}
void boundary_faces_to_host(Kernel_mesh km, unsigned sides, unsigned halves) {move_boundary(mirror(km), false, sides, halves);}
void ghost_faces_to_device(Kernel_mesh km, unsigned sides, unsigned halves) {move_boundary(mirror(km), true, sides, halves);}
void synchronize(Kernel_mesh km)
{
  Mirror& m = mirror(km);
  if (m.group) check_group(m, hexed_b200_group_synchronize(m.group));
  else check(&m, hexed_b200_synchronize(m.ctx0()));
}
std::string transport_description(Kernel_mesh km)
{
  Mirror& m = mirror(km);
  if (!m.group) return "single device";
  int v = 0; long long ex = 0, bytes = 0;
  hexed_b200_group_info(m.group, &v, &ex, &bytes);
  return "NCCL " + std::to_string(v/10000) + "." + std::to_string(v/100%100) + "." + std::to_string(v%100) + " send/recv over " + std::to_string(m.ranks.size())
         + " devices, " + std::to_string(ex) + " exchanges, " + std::to_string(bytes) + " bytes sent";
}

int add_device_bc(Kernel_mesh km, int kind, const std::vector<double*>& inside_faces, const std::vector<double>& params)
{
  Mirror& m = mirror(km);
  int id = -1;
  // every rank registers the condition for the boundary connections it owns (possibly none), so that ids agree across ranks
  std::vector<char> found(inside_faces.size(), 0);
  for (Rank& k : m.ranks) {
    std::unordered_map<const double*, int> con_of_inside; // inside face of a boundary connection -> its row in the rank's def_con
    for (int i : k.boundary_con) con_of_inside[k.face_ptr[k.def_con[size_t(i)*7]]] = i;
    std::vector<int> inside, ghost, normal;
    for (size_t j = 0; j < inside_faces.size(); ++j) {
      auto it = con_of_inside.find(inside_faces[j]);
      if (it == con_of_inside.end()) continue;
      found[j] = 1;
      const int* row = k.def_con.data() + size_t(it->second)*7;
      inside.push_back(row[0]); ghost.push_back(row[1]); normal.push_back(row[6]);
    }
    int rank_id = -1;
    check(&m, hexed_b200_bc_create(k.ctx, kind, int(inside.size()), inside.data(), ghost.data(), normal.data(), params.data(), int(params.size()), &rank_id), k.ctx);
    if (id >= 0 && rank_id != id) throw std::runtime_error("hexed_b200: boundary condition ids diverged between devices");
    id = rank_id;
  }
  for (char f : found) if (!f) throw std::runtime_error("hexed_b200: add_device_bc: not the inside face of a boundary connection");
  return id;
}

void set_device_bc_params(Kernel_mesh km, int id, const std::vector<double>& params)
{
  Mirror& m = mirror(km);
  for (Rank& k : m.ranks) check(&m, hexed_b200_bc_set_params(k.ctx, id, params.data(), int(params.size())), k.ctx);
}

void apply_state_bcs(Kernel_mesh km)
{
  Call call(km, faces, faces);
  call.each([&](hexed_b200_ctx* c) {return hexed_b200_apply_state_bcs(c);});
  call.m.device_bcs_ran = true;
  call.finish();
}

double update_euler(Kernel_mesh km, double safety, int n_steps, double* last_dt, bool use_graph, int n_cheby)
{ // the flow loop of Solver::update (src/Solver.cpp:834-886) for the inviscid case, n_cheby_flow = 1, device boundary conditions
  Call call(km, state | tss | res_cache | faces, state | tss | res_cache | faces);
  single_device_only(call.m, "update_euler (the CUDA-graph flow loop)");
  double dt = 0, t = 0;
  check(&call.m, hexed_b200_update_euler(call.m.ctx0(), safety, n_cheby, n_steps, use_graph, &dt, &t));
  call.m.device_bcs_ran = true;
  call.finish();
  if (last_dt) *last_dt = dt;
  return t;
}

bool is_admissible(Kernel_mesh km, std::vector<int>* record)
{ // Solver::is_admissible, src/Solver.cpp:921-958
  Call call(km, state | faces, 0);
  // a Solver that checks admissibility wants it after every stage: from now on the pipelined Local kernels leave the bits of what
  // they write and the check right after a compute_euler reduces those (any other write to the state or faces, e.g. the uploads of
  // sync_every_call mode, falls back to the full scan)
  bool all_ok = true;
  if (record) record->assign(call.m.tab.elem.size(), 0);
  for (Rank& k : call.m.ranks) { // every rank checks its elements, faces and the mortar faces it holds: started everywhere before any wait
    check(&call.m, hexed_b200_set_option(k.ctx, HEXED_B200_OPT_FUSED_ADMIS, 1), k.ctx);
    check(&call.m, hexed_b200_is_admissible_begin(k.ctx), k.ctx);
  }
  for (Rank& k : call.m.ranks) { // the answer is the AND
    int ok = 0;
    check(&call.m, hexed_b200_is_admissible_finish(k.ctx, &ok), k.ctx);
    all_ok = all_ok && ok != 0;
    if (record && !ok && !k.elem.empty()) { // (all of an admissible rank's records are 0, which `record` already holds: nothing to fetch)
      std::vector<int> local(k.elem.size());
      check(&call.m, hexed_b200_download_record(k.ctx, local.data(), 0, int(local.size())), k.ctx);
      for (size_t i = 0; i < local.size(); ++i) (*record)[k.global_elem[i]] = local[i];
    }
  }
  call.finish();
  return all_ok;
}

// ---- pointwise loops of the artificial-viscosity pipelines ----
namespace
{
std::vector<double> weights_of(const hexed::Basis& b)
{
  auto w = b.node_weights();
  std::vector<double> out(b.row_size);
  for (int i = 0; i < b.row_size; ++i) out[i] = w(i);
  return out;
}
}

void av_scale_velocity(Kernel_mesh km, bool restore)
{ // src/Solver.cpp:467-478 | :567-571
  Call call(km, state, state);
  call.each([&](hexed_b200_ctx* c) {return hexed_b200_av_scale_velocity(c, restore);});
  call.finish();
}

void av_project_forcing(Kernel_mesh km)
{ // src/Solver.cpp:527-541
  Call call(km, state | advection | art_visc, art_visc);
  auto w = weights_of(km.basis);
  auto o = km.basis.orthogonal(km.row_size - 1);
  std::vector<double> orth(km.row_size);
  for (int i = 0; i < km.row_size; ++i) orth[i] = o(i);
  call.each([&](hexed_b200_ctx* c) {return hexed_b200_av_project_forcing(c, w.data(), orth.data());});
  call.finish();
}

double av_finish(Kernel_mesh km, double mult, double us_max, int n_real)
{ // src/Solver.cpp:551-573
  Call call(km, state | art_visc, state | art_visc);
  auto w = weights_of(km.basis);
  double resid = 0;
  single_device_only(call.m, "av_finish");
  check(&call.m, hexed_b200_av_finish(call.m.ctx0(), mult, us_max, n_real, w.data(), &resid));
  call.finish();
  return resid;
}

void interp_vertices(Kernel_mesh km, int target, const std::vector<double>& vertex_values)
{ // math::hypercube_matvec(interp, .) of src/Solver.cpp:652-656, :1021-1031 with interp = [1 - node, node]
  Call call(km, art_visc, art_visc);
  std::vector<double> interp(size_t(2)*km.row_size);
  for (int i = 0; i < km.row_size; ++i) {interp[2*i] = 1. - km.basis.node(i); interp[2*i + 1] = km.basis.node(i);}
  if (vertex_values.size() != call.m.tab.elem.size()*size_t(ipow(2, km.n_dim))) throw std::runtime_error("hexed_b200: interp_vertices: one value per element vertex expected");
  single_device_only(call.m, "interp_vertices");
  check(&call.m, hexed_b200_interp_vertices(call.m.ctx0(), target, vertex_values.data(), interp.data()));
  call.finish();
}

void av_swap(Kernel_mesh km)
{ // src/Solver.cpp:1032-1038
  Call call(km, art_visc, art_visc);
  call.each([&](hexed_b200_ctx* c) {return hexed_b200_av_swap(c);});
  call.finish();
}

void vertex_topology(Kernel_mesh km, const std::vector<int>& elem_vertex, int n_vertex, const std::vector<int>& matchers)
{
  Mirror& m = mirror(km);
  single_device_only(m, "vertex_topology");
  if (elem_vertex.size() != m.tab.elem.size()*size_t(ipow(2, km.n_dim)) || matchers.size() % 8) throw std::runtime_error("hexed_b200: vertex_topology: bad table size");
  check(&m, hexed_b200_vertex_topology(m.ctx0(), elem_vertex.data(), n_vertex, matchers.data(), int(matchers.size()/8)));
}

void av_elwise_ramp(Kernel_mesh km, double scale)
{ // src/Solver.cpp:590-601
  Call call(km, 0, uncert);
  for (Rank& k : call.m.ranks) {
    const int ne = int(k.elem.size());
    std::vector<double> buf(ne);
    for (int e = 0; e < ne; ++e) buf[e] = k.elem[e]->uncert();
    if (ne) check(&call.m, hexed_b200_upload(k.ctx, HEXED_B200_UNCERT, buf.data(), 0, ne), k.ctx);
    check(&call.m, hexed_b200_av_elwise_ramp(k.ctx, scale), k.ctx);
    download_uncert(call.m, k); // the reference leaves the ramped value in Element::uncertainty, whatever the coherence mode
  }
}

void av_elwise_forcing(Kernel_mesh km, bool restore)
{ // src/Solver.cpp:603-612 | :614-619
  Call call(km, art_visc, art_visc);
  call.each([&](hexed_b200_ctx* c) {return hexed_b200_av_elwise_forcing(c, restore);});
  call.finish();
}

void av_elwise_vertices(Kernel_mesh km)
{ // src/Solver.cpp:620-632 with interp = Gauss_lobatto(2).interpolate(nodes) = [1 - node, node]
  Call call(km, art_visc, art_visc);
  std::vector<double> interp(size_t(2)*km.row_size);
  for (int i = 0; i < km.row_size; ++i) {interp[2*i] = 1. - km.basis.node(i); interp[2*i + 1] = km.basis.node(i);}
  single_device_only(call.m, "av_elwise_vertices");
  check(&call.m, hexed_b200_av_elwise_vertices(call.m.ctx0(), interp.data()));
  call.finish();
}

void apply_aux_bcs(Kernel_mesh km, int mode)
{
  const unsigned grp = mode == HEXED_B200_BC_MODE_ADVECTION ? unsigned(faces_wide) : unsigned(faces);
  Call call(km, grp, grp);
  call.each([&](hexed_b200_ctx* c) {return hexed_b200_apply_aux_bcs(c, mode);});
  call.m.device_bcs_ran = true;
  call.finish();
}

void apply_flux_bcs(Kernel_mesh km)
{
  Call call(km, faces, faces);
  call.each([&](hexed_b200_ctx* c) {return hexed_b200_apply_flux_bcs(c);});
  call.m.device_bcs_ran = true;
  call.finish();
}

} // namespace hexed_b200

// ---------------------------------------------------------------------------------------------------------------------
// the reference's entry points (include/kernels.hpp:22-42, include/stabilizing_art_visc.hpp:13)
// ---------------------------------------------------------------------------------------------------------------------
namespace hexed
{

using namespace hexed_b200;

void compute_euler(Kernel_mesh mesh, Kernel_options opts)
{ // src/kernels_convective.cpp:18
  Call call(mesh, state | tss | res_cache | faces, state | res_cache | faces);
  Watch w; w.start(opts.sw_car); w.start(opts.sw_def);
  if (call.multi()) check_group(call.m, hexed_b200_group_compute_euler(call.m.group, options(opts)));
  else check(&call.m, hexed_b200_compute_euler(call.m.ctx0(), options(opts)));
  prefetch_boundary(call.m);
  count_convective(mesh, opts);
  call.finish();
}

void compute_advection(Kernel_mesh mesh, Kernel_options opts, double advect_length)
{ // src/kernels_convective.cpp:19
  Call call(mesh, state | tss | advection | res_cache | faces_wide, advection | res_cache | faces_wide);
  Watch w; w.start(opts.sw_car); w.start(opts.sw_def);
  single_device_only(call.m, "compute_advection");
  check(&call.m, hexed_b200_compute_advection(call.m.ctx0(), options(opts), advect_length));
  count_convective(mesh, opts);
  call.finish();
}

void compute_navier_stokes(Kernel_mesh mesh, Kernel_options opts, std::function<void()> flux_bc, Transport_model visc, Transport_model therm_cond)
{ // src/kernels_diffusive.cpp:28-29
  Call call(mesh, state | tss | art_visc | res_cache | faces, state | res_cache | faces);
  Watch w; w.start(opts.sw_car); w.start(opts.sw_def);
  Flux_bc_thunk thunk {&call.m, &flux_bc};
  if (call.multi()) check_group(call.m, hexed_b200_group_compute_navier_stokes(call.m.group, options(opts), &Flux_bc_thunk::call, &thunk, transport(visc), transport(therm_cond)));
  else check(&call.m, hexed_b200_compute_navier_stokes(call.m.ctx0(), options(opts), &Flux_bc_thunk::call, &thunk, transport(visc), transport(therm_cond)));
  prefetch_boundary(call.m);
  count_diffusive(mesh, opts);
  call.finish();
}

void compute_smooth_av(Kernel_mesh mesh, Kernel_options opts, std::function<void()> flux_bc, double diff_time, double chebyshev_step)
{ // src/kernels_diffusive.cpp:30-31
  Call call(mesh, state | tss | art_visc | res_cache | faces, art_visc | res_cache | faces);
  Watch w; w.start(opts.sw_car); w.start(opts.sw_def);
  Flux_bc_thunk thunk {&call.m, &flux_bc};
  single_device_only(call.m, "compute_smooth_av");
  check(&call.m, hexed_b200_compute_smooth_av(call.m.ctx0(), options(opts), &Flux_bc_thunk::call, &thunk, diff_time, chebyshev_step));
  count_diffusive(mesh, opts);
  call.finish();
}

void compute_fix_therm_admis(Kernel_mesh mesh, Kernel_options opts, std::function<void()> flux_bc)
{ // src/kernels_diffusive.cpp:32
  Call call(mesh, state | tss | art_visc | res_cache | faces, state | res_cache | faces);
  Watch w; w.start(opts.sw_car); w.start(opts.sw_def);
  Flux_bc_thunk thunk {&call.m, &flux_bc};
  single_device_only(call.m, "compute_fix_therm_admis");
  check(&call.m, hexed_b200_compute_fix_therm_admis(call.m.ctx0(), options(opts), &Flux_bc_thunk::call, &thunk));
  count_diffusive(mesh, opts);
  call.finish();
}

double max_dt_euler(Kernel_mesh mesh, Kernel_options opts, double convective_safety, double diffusive_safety, bool local_time)
{ // src/kernels_max_dt.cpp:14
  return max_dt_call(mesh, opts, state, 0, convective_safety, diffusive_safety, local_time, no_transport, no_transport, 0.);
}

double max_dt_navier_stokes(Kernel_mesh mesh, Kernel_options opts, double convective_safety, double diffusive_safety, bool local_time,
                            Transport_model visc, Transport_model therm_cond)
{ // src/kernels_max_dt.cpp:15-17
  return max_dt_call(mesh, opts, state | art_visc, 1, convective_safety, diffusive_safety, local_time, transport(visc), transport(therm_cond), 0.);
}

double max_dt_advection(Kernel_mesh mesh, Kernel_options opts, double convective_safety, double diffusive_safety, bool local_time, double advect_length)
{ // src/kernels_max_dt.cpp:18-19
  return max_dt_call(mesh, opts, state | advection, 2, convective_safety, diffusive_safety, local_time, no_transport, no_transport, advect_length);
}

double max_dt_smooth_av(Kernel_mesh mesh, Kernel_options opts, double convective_safety, double diffusive_safety, bool local_time)
{ // src/kernels_max_dt.cpp:20
  return max_dt_call(mesh, opts, state | art_visc, 3, convective_safety, diffusive_safety, local_time, no_transport, no_transport, 0.);
}

double max_dt_fix_therm_admis(Kernel_mesh mesh, Kernel_options opts, double convective_safety, double diffusive_safety, bool local_time)
{ // src/kernels_max_dt.cpp:21
  return max_dt_call(mesh, opts, state | art_visc, 4, convective_safety, diffusive_safety, local_time, no_transport, no_transport, 0.);
}

void compute_prolong(Kernel_mesh mesh, bool scale, bool offset)
{ // src/kernels_convective.cpp:23-26
  Call call(mesh, faces, faces);
  if (call.multi()) check_group(call.m, hexed_b200_group_exchange(call.m.group, offset ? 1 : 0)); // a remote coarse face must be current before it is prolonged here
  call.each([&](hexed_b200_ctx* c) {return hexed_b200_compute_prolong(c, scale, offset);});
  call.finish();
}

void compute_restrict(Kernel_mesh mesh, bool scale, bool offset)
{ // src/kernels_convective.cpp:28-31
  Call call(mesh, faces, faces);
  call.each([&](hexed_b200_ctx* c) {return hexed_b200_compute_restrict(c, scale, offset);});
  call.finish();
}

void compute_prolong_advection(Kernel_mesh mesh)
{ // src/kernels_convective.cpp:33-36
  Call call(mesh, faces_wide, faces_wide);
  call.each([&](hexed_b200_ctx* c) {return hexed_b200_compute_prolong_advection(c);});
  call.finish();
}

std::unique_ptr<Face_permutation_dynamic> face_permutation(int n_dim, int row_size, Connection_direction dir, double* data)
{ // src/kernels_convective.cpp:38-41. `data` is one host face ([n_var][nfq]); the integer table comes from the library.
  return std::unique_ptr<Face_permutation_dynamic>(new Face_permutation_host(n_dim, row_size, dir, data));
}

void compute_write_face(Kernel_mesh mesh)
{ // src/kernels_convective.cpp:43-46
  Call call(mesh, state, faces);
  call.each([&](hexed_b200_ctx* c) {return hexed_b200_compute_write_face(c);});
  call.finish();
}

void compute_write_face_advection(Kernel_mesh mesh)
{ // src/kernels_convective.cpp:48-51
  Call call(mesh, state | advection, faces_wide);
  call.each([&](hexed_b200_ctx* c) {return hexed_b200_compute_write_face_advection(c);});
  call.finish();
}

void compute_write_face_smooth_av(Kernel_mesh mesh)
{ // src/kernels_convective.cpp:53-56
  Call call(mesh, state | art_visc, faces);
  call.each([&](hexed_b200_ctx* c) {return hexed_b200_compute_write_face_smooth_av(c);});
  call.finish();
}

void stabilizing_art_visc(Kernel_mesh mesh, double char_speed)
{ // src/stabilizing_art_visc.cpp:8-66
  Call call(mesh, state, uncert);
  call.each([&](hexed_b200_ctx* c) {return hexed_b200_stabilizing_art_visc(c, char_speed);});
  call.finish();
}

}
