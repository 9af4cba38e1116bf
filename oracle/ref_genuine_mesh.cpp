/* ref_genuine_mesh.cpp -- a mesh made of the REFERENCE'S GENUINE storage classes, driven through include/kernels.hpp.
 *
 * TEST INFRASTRUCTURE ONLY (nothing under hexed_b200/ links or calls it). Compiled twice by oracle/Makefile.ref:
 *   oracle/_ref/libhexed_ref.so               + the reference's own src/kernels_*.cpp   -> hexed::compute_euler etc. ARE the reference
 *   oracle/_ref/libhexed_adapter_genuine*.so  + hexed_b200/host/adapter.cpp built with -DHEXED_B200_WITH_HEXED_HEADERS against the
 *                                               genuine include/kernels.hpp -> the same calls land in the B200 library
 * Both stand up the SAME object graph (deterministic construction), so running the same call sequence in each and comparing the
 * exported data is the drop-in test SURVEY section 8b asks for: genuine `Element` / `Deformed_element` (src/Element.cpp,
 * src/Deformed_element.cpp), `Element_face_connection`, `Refined_connection` (+ its `Fine_connection`s and `Refined_face`),
 * `Typed_bound_connection` (include/connection.hpp:52-400), `Vertex` merging, the reference's layouts and aliasing -- none of it
 * restated. The `Kernel_mesh` views follow the reference's ordering (include/Mesh_by_type.hpp:70-95, src/Accessible_mesh.cpp:126-147):
 * car_cons = conformal Cartesian connections then the fine connections of Cartesian hanging faces; def_cons = conformal deformed
 * connections, fine connections of deformed hanging faces, then ALL boundary connections (Cartesian first); elems = Cartesian then
 * deformed; ref_faces = Cartesian then deformed.
 *
 * The one thing restated here is the connection pass of `Solver::calc_jacobian` (src/Solver.cpp:288-364; Solver.cpp itself needs
 * HDF5 and the whole mesh layer), used to give the mesh consistent metric terms after the genuine `set_jacobian` of every element.
 */
#include <kernels.hpp>
#include <stabilizing_art_visc.hpp>
#include <Gauss_legendre.hpp>
#include <connection.hpp>

#include <algorithm>
#include <cstring>
#include <map>
#include <memory>
#include <vector>

#include "flat_mesh.h"

#ifdef HR_WITH_ADAPTER
#include "../hexed_b200/host/adapter.hpp"
#endif

namespace
{
using namespace hexed;

int ipow(int b, int e) {int r = 1; for (int i = 0; i < e; ++i) r *= b; return r;}

template <typename T>
class Ptr_seq : public Sequence<T&>
{
  public:
  std::vector<T*> v;
  int size() override {return int(v.size());}
  T& operator[](int i) override {return *v[i];}
};

Stopwatch_tree make_tree()
{
  return Stopwatch_tree("element", {{"neighbor", Stopwatch_tree("connection")}, {"local", Stopwatch_tree("element")},
                                    {"reconcile LDG flux", Stopwatch_tree("element")}, {"compute time step", Stopwatch_tree("element")}});
}

Transport_model transport(ho_transport t)
{
  if (!t.is_viscous) return Transport_model::inviscid();
  if (t.ref_val == 0.) return Transport_model::constant(t.const_val);
  return Transport_model::sutherland(t.ref_val, t.ref_temp, t.temp_offset);
}

struct Genuine_mesh
{
  int nd, rs, nq, nfq, nv;
  Storage_params params;
  Gauss_legendre basis;
  // owners, in creation order
  std::vector<std::unique_ptr<Element>> car;
  std::vector<std::unique_ptr<Deformed_element>> def;
  std::vector<std::unique_ptr<Element_face_connection<Element>>> car_con;
  std::vector<std::unique_ptr<Element_face_connection<Deformed_element>>> def_con;
  std::vector<std::unique_ptr<Refined_connection<Element>>> car_ref;
  std::vector<std::unique_ptr<Refined_connection<Deformed_element>>> def_ref;
  std::vector<std::unique_ptr<Typed_bound_connection<Element>>> car_bc;
  std::vector<std::unique_ptr<Typed_bound_connection<Deformed_element>>> def_bc;
  // views
  Ptr_seq<Kernel_element> s_car, s_def, s_all;
  Ptr_seq<Kernel_connection> s_ccon, s_dcon;
  Ptr_seq<Refined_face> s_ref;
  std::vector<Element*> all_elems;               // car then def
  std::vector<Face_connection<Deformed_element>*> def_face_cons; // what Solver::calc_jacobian calls def_cons
  std::vector<Boundary_connection*> bcs;         // car then def
  Stopwatch_tree sw_car = make_tree(), sw_def = make_tree(), sw_pr {"refined face"};
  std::string error;

  Genuine_mesh(int n_dim, int row_size) : nd{n_dim}, rs{row_size}, params{2, n_dim + 2, n_dim, row_size}, basis{row_size}
  {nq = ipow(rs, nd); nfq = nq/rs; nv = nd + 2;}

  Kernel_mesh mesh() {return {nd, rs, basis, s_ccon, s_dcon, s_car, s_def, s_all, s_ref};}
  Kernel_options options(ho_options o) {return {sw_car, sw_def, sw_pr, o.dt, o.i_stage, bool(o.compute_residual), bool(o.use_filter)};}

  /*! `cell_kind[n^nd]` (row-major, last dimension fastest): bit 0 = deformed, bit 1 = refined into 2^nd children.
   *  Vertices of deformed elements are displaced by `warp`*h*sin(2 pi x)sin(2 pi y)... afterwards. */
  void build(int n, const int* cell_kind, double warp)
  {
    const double root = 1.;
    struct Key {int level; std::array<int, 3> pos; bool operator<(const Key& o) const {return std::tie(level, pos) < std::tie(o.level, o.pos);}};
    std::map<Key, Element*> index;
    std::map<Element*, bool> is_def;
    std::vector<std::pair<Key, Element*>> order;
    auto cell_id = [&](std::array<int, 3> c) {int id = 0; for (int d = 0; d < nd; ++d) id = id*n + c[d]; return id;};
    auto add = [&](int level, std::array<int, 3> pos, bool deformed) {
      std::vector<int> p(pos.begin(), pos.begin() + nd);
      // nominal size = mesh_size/2^ref_level (src/Element.cpp:11): the coarse cells are level `L0` of a unit root
      Element* e;
      if (deformed) {def.emplace_back(new Deformed_element(params, p, root/n, level)); e = def.back().get();}
      else {car.emplace_back(new Element(params, p, root/n, level)); e = car.back().get();}
      index[Key{level, pos}] = e; is_def[e] = deformed;
      order.push_back({Key{level, pos}, e});
    };
    const int n_cell = ipow(n, nd);
    for (int id = 0; id < n_cell; ++id) {
      std::array<int, 3> c {0, 0, 0};
      for (int d = nd - 1, r = id; d >= 0; --d) {c[d] = r%n; r /= n;}
      const bool deformed = cell_kind[id] & 1;
      if (cell_kind[id] & 2) {
        for (int k = 0; k < ipow(2, nd); ++k) {
          std::array<int, 3> p {0, 0, 0};
          for (int d = 0; d < nd; ++d) p[d] = 2*c[d] + ((k >> (nd - 1 - d)) & 1);
          add(1, p, deformed);
        }
      } else add(0, c, deformed);
    }
    auto fine_children = [&](std::array<int, 3> cell, int d, int side) { // children of `cell` touching its face (d, side), row-major in the tangential dims
      std::vector<Element*> fine;
      for (int k = 0; k < ipow(2, nd - 1); ++k) {
        std::array<int, 3> p {0, 0, 0};
        int bit = nd - 2;
        for (int t = 0; t < nd; ++t) {
          if (t == d) p[t] = 2*cell[t] + side;
          else {p[t] = 2*cell[t] + ((k >> bit) & 1); --bit;}
        }
        fine.push_back(index.at(Key{1, p}));
      }
      return fine;
    };
    for (auto& [key, e] : order) {
      const int cells = n*(key.level ? 2 : 1);
      for (int d = 0; d < nd; ++d) {
        // positive side
        if (key.pos[d] + 1 < cells) {
          auto nb = key.pos; nb[d] += 1;
          auto it = index.find(Key{key.level, nb});
          if (it != index.end()) connect(e, it->second, d, is_def[e] && is_def[it->second]);
          else if (key.level == 0) { // coarse below, children above: not reversed
            auto fine = fine_children(nb, d, 0);
            connect_refined(e, fine, d, false, is_def);
          }
        } else bound(e, d, true, is_def[e]);
        // negative side
        if (key.pos[d] == 0) bound(e, d, false, is_def[e]);
        else if (key.level == 0) {
          auto nb = key.pos; nb[d] -= 1;
          if (index.find(Key{0, nb}) == index.end()) { // refined neighbour below: reversed
            auto fine = fine_children(nb, d, 1);
            connect_refined(e, fine, d, true, is_def);
          }
        }
      }
    }
    // displace the vertices of deformed elements (merged vertices move once: the displacement is a function of position)
    if (warp != 0.) {
      for (auto& e : def) {
        for (int i = 0; i < ipow(2, nd); ++i) {
          Vertex& v = e->vertex(i);
          if (v.record.size()) continue; // already moved
          double s = warp*(root/n)/2.;
          for (int d = 0; d < nd; ++d) s *= std::sin(2*M_PI*(v.pos(d) + 0.13*(d + 1)));
          Mat<3> p = v.pos;
          for (int d = 0; d < nd; ++d) p(d) += s*(1. + 0.3*d);
          v.pos = p;
          v.record.push_back(1);
        }
      }
      for (auto& e : def) for (int i = 0; i < ipow(2, nd); ++i) e->vertex(i).record.clear();
    }
    // the reference leaves connection storage and Jacobian data uninitialised (Eigen::VectorXd of a given size): zero it so that
    // exported data is deterministic (element data is initialised by the Element constructor and left alone)
    {
      const int state_sz = nv*nfq, face_sz = std::max(2*state_sz, (nd + rs)*nfq);
      auto zero = [](double* p, int k) {for (int i = 0; i < k; ++i) p[i] = 0.;};
      for (auto& e : def) zero(e->reference_level_normals(), (nd*nd + 1)*nq);
      for (auto& c : car_con) zero(c->state(0, false), 2*face_sz);
      for (auto& c : def_con) zero(c->state(0, false), 2*face_sz + 2*nd*nfq);
      for (auto& r : car_ref) {zero(r->coarse_state(), 3*state_sz); for (int i = 0; i < r->n_fine_elements(); ++i) zero(r->connection(i).state(0, false), 2*face_sz);}
      for (auto& r : def_ref) {
        zero(r->coarse_state(), 3*state_sz);
        auto dir = r->direction();
        zero(r->coarse_element().face_normal(2*dir.i_dim[r->order_reversed()] + dir.face_sign[r->order_reversed()]), nd*nfq);
        for (int i = 0; i < r->n_fine_elements(); ++i) zero(r->connection(i).state(0, false), 2*face_sz + 2*nd*nfq);
      }
      for (auto& b : car_bc) zero(b->state(0, false), 2*face_sz + 2*nd*nfq);
      for (auto& b : def_bc) zero(b->state(0, false), 2*face_sz + 2*nd*nfq);
    }
    // views in the reference's order
    for (auto& e : car) {s_car.v.push_back(e.get()); s_all.v.push_back(e.get()); all_elems.push_back(e.get());}
    for (auto& e : def) {s_def.v.push_back(e.get()); s_all.v.push_back(e.get()); all_elems.push_back(e.get());}
    for (auto& c : car_con) s_ccon.v.push_back(c.get());
    for (int want : {1, 2, 4}) for (auto& r : car_ref) if (r->n_fine_elements() == want)
      for (int i = 0; i < want; ++i) s_ccon.v.push_back(&r->connection(i));
    for (auto& c : def_con) {s_dcon.v.push_back(c.get()); def_face_cons.push_back(c.get());}
    for (int want : {1, 2, 4}) for (auto& r : def_ref) if (r->n_fine_elements() == want)
      for (int i = 0; i < want; ++i) {s_dcon.v.push_back(&r->connection(i)); def_face_cons.push_back(&r->connection(i));}
    for (auto& b : car_bc) {s_dcon.v.push_back(b.get()); def_face_cons.push_back(b.get()); bcs.push_back(b.get());}
    for (auto& b : def_bc) {s_dcon.v.push_back(b.get()); def_face_cons.push_back(b.get()); bcs.push_back(b.get());}
    for (int want : {1, 2, 4}) for (auto& r : car_ref) if (r->n_fine_elements() == want) s_ref.v.push_back(&r->refined_face);
    for (int want : {1, 2, 4}) for (auto& r : def_ref) if (r->n_fine_elements() == want) s_ref.v.push_back(&r->refined_face);
  }

  void connect(Element* lo, Element* hi, int d, bool both_def)
  {
    if (both_def) {
      def_con.emplace_back(new Element_face_connection<Deformed_element>(
        {static_cast<Deformed_element*>(lo), static_cast<Deformed_element*>(hi)}, Con_dir<Deformed_element>{{d, d}, {true, false}}));
    } else car_con.emplace_back(new Element_face_connection<Element>({lo, hi}, Con_dir<Element>{d}));
  }
  void connect_refined(Element* coarse, std::vector<Element*> fine, int d, bool rev, std::map<Element*, bool>& is_def)
  {
    bool all_def = is_def[coarse];
    for (Element* f : fine) all_def = all_def && is_def[f];
    if (all_def) {
      std::vector<Deformed_element*> f;
      for (Element* e : fine) f.push_back(static_cast<Deformed_element*>(e));
      def_ref.emplace_back(new Refined_connection<Deformed_element>(static_cast<Deformed_element*>(coarse), f,
                                                                    Con_dir<Deformed_element>{{d, d}, {true, false}}, rev));
    } else car_ref.emplace_back(new Refined_connection<Element>(coarse, fine, Con_dir<Element>{d}, rev));
  }
  void bound(Element* e, int d, bool positive, bool deformed)
  {
    if (deformed) def_bc.emplace_back(new Typed_bound_connection<Deformed_element>(*static_cast<Deformed_element*>(e), d, positive, 0));
    else car_bc.emplace_back(new Typed_bound_connection<Element>(*e, d, positive, 0));
  }

  //! `Solver::calc_jacobian` without snapping (src/Solver.cpp:273-381): genuine set_jacobian, then the connection pass restated
  void calc_jacobian()
  {
    for (Element* e : all_elems) e->set_jacobian(basis);
    for (auto* con : def_face_cons) {double* n = con->normal(); for (int i = 0; i < nd*nfq; ++i) n[i] = 0.;}  // :291-296
    compute_prolong(mesh(), true);                                                                             // :299
    for (auto& ref : def_ref) {                                                                                // :300-318
      bool rev = ref->order_reversed();
      auto dir = ref->direction();
      int sign = 1 - 2*(dir.flip_normal(0) != dir.flip_normal(1));
      for (int i_fine = 0; i_fine < ref->n_fine_elements(); ++i_fine) {
        auto& fine = ref->connection(i_fine);
        double* face [2] {fine.state(rev, false), fine.state(!rev, false)};
        auto fp = face_permutation(nd, rs, dir, face[1]);
        fp->match_faces();
        for (int i = 0; i < nd*nfq; ++i) face[1][i] = sign*face[0][i];
        fp->restore();
      }
    }
    for (auto* bc : bcs) {                                                                                     // :319-326
      double* in_f = bc->inside_face(false); double* gh_f = bc->ghost_face(false);
      for (int i = 0; i < nd*nfq; ++i) gh_f[i] = in_f[i];
    }
    for (auto* con : def_face_cons) {                                                                          // :327-352
      double* elem_nrml [2] {con->state(0, false), con->state(1, false)};
      auto dir = con->direction();
      auto fp = face_permutation(nd, rs, dir, elem_nrml[1]);
      fp->match_faces();
      int sign [2];
      for (int i_side : {0, 1}) sign[i_side] = 1 - 2*dir.flip_normal(i_side);
      for (int i = 0; i < nd*nfq; ++i) {
        double n = 0;
        for (int i_side : {0, 1}) n += 0.5*sign[i_side]*elem_nrml[i_side][i];
        for (int i_side : {0, 1}) elem_nrml[i_side][i] = sign[i_side]*n;
      }
      fp->restore();
      for (int i_side = 0; i_side < 2; ++i_side) {
        double* n = con->normal(i_side);
        for (int i = 0; i < nd*nfq; ++i) n[i] = elem_nrml[i_side][i];
      }
    }
    for (auto& ref : def_ref) {                                                                                // :353-364
      bool rev = ref->order_reversed();
      auto& elem = ref->connection(0).element(rev);
      auto dir = ref->direction();
      int i_face = 2*dir.i_dim[rev] + dir.face_sign[rev];
      double* nrml = elem.face_normal(i_face);
      double* state = elem.face(i_face, false);
      for (int i = 0; i < nd*nfq; ++i) nrml[i] = state[i];
    }
  }

  // ---- every double the kernels can see, in one deterministic order (export / import between the two builds, and result comparison)
  template <typename F> void walk(F f)
  {
    const int n_vert = ipow(2, nd);
    for (Element* e : all_elems) {
      f(e->state(), params.n_dof_numeric());
      for (int i = 0; i < n_vert; ++i) f(&e->vertex_time_step_scale(i), 1);
      f(&e->uncertainty, 1);
      if (e->get_is_deformed()) f(e->reference_level_normals(), (nd*nd + 1)*nq);
    }
    const int state_sz = nv*nfq, face_sz = std::max(2*state_sz, (nd + rs)*nfq);
    for (auto& c : car_con) f(c->state(0, false), 2*face_sz);
    for (auto& c : def_con) f(c->state(0, false), 2*face_sz + 2*nd*nfq);
    for (auto& r : car_ref) {
      f(r->coarse_state(), 3*state_sz);
      for (int i = 0; i < r->n_fine_elements(); ++i) f(r->connection(i).state(0, false), 2*face_sz);
    }
    for (auto& r : def_ref) {
      f(r->coarse_state(), 3*state_sz);
      auto& coarse = r->coarse_element();
      auto dir = r->direction();
      f(coarse.face_normal(2*dir.i_dim[r->order_reversed()] + dir.face_sign[r->order_reversed()]), nd*nfq);
      for (int i = 0; i < r->n_fine_elements(); ++i) f(r->connection(i).state(0, false), 2*face_sz + 2*nd*nfq);
    }
    for (auto* b : bcs) f(b->state(0, false), 2*face_sz + 2*nd*nfq);
  }
};

template <typename F> int guarded(Genuine_mesh* g, F f)
{
  try {f(); return 0;}
  catch (const std::exception& e) {g->error = e.what(); return std::string(e.what()) == "demand for invalid kernel" ? 1 : 2;}
  catch (...) {g->error = "unknown exception"; return 3;}
}

} // namespace

extern "C" {

void* hr_gm_create(int n_dim, int row_size, int n, const int* cell_kind, double warp)
{
  try {
    auto* g = new Genuine_mesh(n_dim, row_size);
    g->build(n, cell_kind, warp);
    return g;
  } catch (const std::exception& e) {fprintf(stderr, "hr_gm_create: %s\n", e.what()); return nullptr;}
}
void hr_gm_destroy(void* h) {delete static_cast<Genuine_mesh*>(h);}
const char* hr_gm_error(void* h) {return static_cast<Genuine_mesh*>(h)->error.c_str();}
//! counts: n_car, n_def, car_cons, def_cons, ref_faces, boundary connections
void hr_gm_counts(void* h, int* out)
{
  auto* g = static_cast<Genuine_mesh*>(h);
  out[0] = int(g->car.size()); out[1] = int(g->def.size()); out[2] = g->s_ccon.size(); out[3] = g->s_dcon.size();
  out[4] = g->s_ref.size(); out[5] = int(g->bcs.size());
}
long hr_gm_blob_size(void* h) {long n = 0; static_cast<Genuine_mesh*>(h)->walk([&](double*, int k) {n += k;}); return n;}
void hr_gm_export(void* h, double* buf)
{static_cast<Genuine_mesh*>(h)->walk([&](double* p, int k) {std::memcpy(buf, p, sizeof(double)*k); buf += k;});}
void hr_gm_import(void* h, const double* buf)
{static_cast<Genuine_mesh*>(h)->walk([&](double* p, int k) {std::memcpy(p, buf, sizeof(double)*k); buf += k;});}
int hr_gm_calc_jacobian(void* h) {auto* g = static_cast<Genuine_mesh*>(h); return guarded(g, [&]() {g->calc_jacobian();});}
//! quadrature point positions [n_elem][n_dim][nq] from the genuine `position()`
void hr_gm_positions(void* h, double* out)
{
  auto* g = static_cast<Genuine_mesh*>(h);
  for (size_t e = 0; e < g->all_elems.size(); ++e) for (int q = 0; q < g->nq; ++q) {
    auto p = g->all_elems[e]->position(g->basis, q);
    for (int d = 0; d < g->nd; ++d) out[(e*g->nd + d)*g->nq + q] = p[d];
  }
}
//! element slots [first, first + n) of every element, [n_elem][n][nq]
void hr_gm_get_slots(void* h, int first, int n, double* out)
{
  auto* g = static_cast<Genuine_mesh*>(h);
  for (size_t e = 0; e < g->all_elems.size(); ++e) std::memcpy(out + e*n*g->nq, g->all_elems[e]->state() + size_t(first)*g->nq, sizeof(double)*n*g->nq);
}
void hr_gm_set_slots(void* h, int first, int n, const double* in)
{
  auto* g = static_cast<Genuine_mesh*>(h);
  for (size_t e = 0; e < g->all_elems.size(); ++e) std::memcpy(g->all_elems[e]->state() + size_t(first)*g->nq, in + e*n*g->nq, sizeof(double)*n*g->nq);
}
//! `Copy::apply_state` / `apply_flux` = copy_state (src/Boundary_condition.cpp:12-23,450-458) on every boundary connection
void hr_gm_bc_copy(void* h)
{
  auto* g = static_cast<Genuine_mesh*>(h);
  for (auto* b : g->bcs) std::memcpy(b->ghost_face(false), b->inside_face(false), sizeof(double)*2*g->nv*g->nfq);
}
//! `Freestream::apply_state` (src/Boundary_condition.cpp:66-76)
void hr_gm_bc_freestream(void* h, const double* fs)
{
  auto* g = static_cast<Genuine_mesh*>(h);
  for (auto* b : g->bcs) {
    double* gf = b->ghost_face(false);
    for (int v = 0; v < g->nv; ++v) for (int q = 0; q < g->nfq; ++q) gf[v*g->nfq + q] = fs[v];
  }
}

//! the tail of `Solver::initialize` (src/Solver.cpp:400-401): extrapolate the state to the faces, prolong it onto the mortar faces
int hr_gm_compute_write_face(void* h)
{auto* g = static_cast<Genuine_mesh*>(h); return guarded(g, [&]() {compute_write_face(g->mesh()); compute_prolong(g->mesh());});}
int hr_gm_compute_euler(void* h, ho_options o) {auto* g = static_cast<Genuine_mesh*>(h); return guarded(g, [&]() {compute_euler(g->mesh(), g->options(o));});}
int hr_gm_compute_navier_stokes(void* h, ho_options o, ho_transport visc, ho_transport cond)
{
  auto* g = static_cast<Genuine_mesh*>(h);
  return guarded(g, [&]() {
    compute_navier_stokes(g->mesh(), g->options(o), [g]() {
#ifdef HR_WITH_ADAPTER
      if (hexed_b200::sync_mode() == hexed_b200::resident) hexed_b200::boundary_faces_to_host(g->mesh());
#endif
      // `Copy::apply_flux` on the LDG halves; the host loop Solver::apply_flux_bcs runs inside the callback (src/Solver.cpp:69-81)
      for (auto* b : g->bcs) std::memcpy(b->ghost_face(true), b->inside_face(true), sizeof(double)*g->nv*g->nfq);
#ifdef HR_WITH_ADAPTER
      if (hexed_b200::sync_mode() == hexed_b200::resident) hexed_b200::ghost_faces_to_device(g->mesh());
#endif
    }, transport(visc), transport(cond));
  });
}
int hr_gm_max_dt(void* h, int pde, double sc, double sd, int local_time, ho_transport visc, ho_transport cond, double* dt_out)
{
  auto* g = static_cast<Genuine_mesh*>(h);
  return guarded(g, [&]() {
    ho_options none {1., 0, 0, 0};
    if (pde == HO_EULER) *dt_out = max_dt_euler(g->mesh(), g->options(none), sc, sd, local_time);
    else *dt_out = max_dt_navier_stokes(g->mesh(), g->options(none), sc, sd, local_time, transport(visc), transport(cond));
  });
}
int hr_gm_stabilizing_art_visc(void* h, double char_speed)
{auto* g = static_cast<Genuine_mesh*>(h); return guarded(g, [&]() {stabilizing_art_visc(g->mesh(), char_speed);});}
//! work units the drivers reported through the Stopwatch_tree side-contract (include/kernel_factory.hpp:32-45): car local, def local, car neighbor, def neighbor, prolong/restrict
void hr_gm_work_units(void* h, int* out)
{
  auto* g = static_cast<Genuine_mesh*>(h);
  out[0] = g->sw_car.children.at("local").work_units_completed; out[1] = g->sw_def.children.at("local").work_units_completed;
  out[2] = g->sw_car.children.at("neighbor").work_units_completed; out[3] = g->sw_def.children.at("neighbor").work_units_completed;
  out[4] = g->sw_pr.work_units_completed;
}

#ifdef HR_WITH_ADAPTER
int hr_gm_adapter_present(void) {return 1;}
void hr_gm_set_sync_mode(int resident) {hexed_b200::set_sync_mode(resident ? hexed_b200::resident : hexed_b200::sync_every_call);}
int hr_gm_to_host(void* h) {auto* g = static_cast<Genuine_mesh*>(h); return guarded(g, [&]() {hexed_b200::to_host(g->mesh());});}
int hr_gm_to_device(void* h) {auto* g = static_cast<Genuine_mesh*>(h); return guarded(g, [&]() {hexed_b200::to_device(g->mesh());});}
int hr_gm_boundary_faces_to_host(void* h) {auto* g = static_cast<Genuine_mesh*>(h); return guarded(g, [&]() {hexed_b200::boundary_faces_to_host(g->mesh());});}
int hr_gm_ghost_faces_to_device(void* h) {auto* g = static_cast<Genuine_mesh*>(h); return guarded(g, [&]() {hexed_b200::ghost_faces_to_device(g->mesh());});}
void hr_gm_release(void) {hexed_b200::release();}
//! the flattened tables of the genuine pointer graph (device-free): counts = n_car, n_def, n_face_slot, n_normal_slot, car cons, def cons, ref faces, boundary cons
int hr_gm_flatten_counts(void* h, int* out)
{
  auto* g = static_cast<Genuine_mesh*>(h);
  return guarded(g, [&]() {
    auto t = hexed_b200::flatten(g->mesh());
    out[0] = t.n_car; out[1] = t.n_def; out[2] = t.n_face_slot; out[3] = t.n_normal_slot;
    out[4] = int(t.car_con.size())/3; out[5] = int(t.def_con.size())/7; out[6] = int(t.ref_face.size())/7; out[7] = int(t.boundary_con.size());
  });
}
#else
int hr_gm_adapter_present(void) {return 0;}
#endif

} // extern "C"
