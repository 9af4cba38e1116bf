/* partition.cpp -- see partition.hpp. Pure integer host logic (no device, no hexed types): tests/test_partition_cpp.py checks it
 * table for table against the Python partitioner the torchrun bench uses (hexed_b200/partition.py). */
#include "partition.hpp"

#include <algorithm>
#include <map>
#include <numeric>
#include <queue>
#include <stdexcept>
#include <unordered_map>

namespace hexed_b200
{

std::uint64_t morton_key(const std::array<int, 3>& index, int n_dim)
{
  std::uint64_t key = 0;
  for (int b = 0; b < 21; ++b) for (int d = 0; d < n_dim; ++d) {
    key |= ((std::uint64_t(index[d]) >> b) & 1u) << (b*n_dim + (n_dim - 1 - d));
  }
  return key;
}

std::vector<int> split_by_curve(const std::vector<std::uint64_t>& keys, int n_parts, const std::vector<double>& weights)
{
  const size_t n = keys.size();
  std::vector<size_t> order(n);
  std::iota(order.begin(), order.end(), size_t(0));
  std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) {return keys[a] < keys[b];});
  auto w = [&](size_t e) {return weights.empty() ? 1. : weights[e];};
  double total = 0;
  for (size_t e = 0; e < n; ++e) total += w(e);
  std::vector<int> part(n, 0);
  double cum = 0;
  for (size_t i = 0; i < n; ++i) {
    const size_t e = order[i];
    cum += w(e);
    int p = int((cum - 0.5*w(e))*n_parts/total); // the element goes where its midpoint along the curve falls
    part[e] = std::min(std::max(p, 0), n_parts - 1);
  }
  return part;
}

namespace
{
std::vector<double> default_weights(int n_car, size_t n)
{ // weight 1 Cartesian / ~1.4 deformed (bytes per element-stage 67.4 KB vs 92.4 KB)
  std::vector<double> w(n, 1.);
  for (size_t e = size_t(n_car); e < n; ++e) w[e] = 1.4;
  return w;
}
}

std::vector<int> owners_by_morton(const std::vector<std::array<int, 3>>& coords, int n_dim, int n_car, int n_parts)
{
  std::vector<std::uint64_t> keys(coords.size());
  for (size_t e = 0; e < coords.size(); ++e) keys[e] = morton_key(coords[e], n_dim);
  return split_by_curve(keys, n_parts, default_weights(n_car, coords.size()));
}

std::vector<int> owners_by_graph(const Mesh_graph& g, int n_parts)
{
  const int nf = 2*g.n_dim, ne = g.n_car + g.n_def, n_elem_slots = nf*ne;
  std::vector<std::vector<int>> adj(ne);
  // mortar faces connect a fine element to the coarse element of their refined face
  std::unordered_map<int, int> mortar_coarse;
  for (size_t r = 0; r < g.ref_face.size()/7; ++r) for (int k = 1; k < 5; ++k) if (g.ref_face[r*7 + k] >= 0) mortar_coarse[g.ref_face[r*7 + k]] = g.ref_face[r*7];
  auto elem_of = [&](int slot) -> int {
    if (slot < n_elem_slots) return slot/nf;
    auto it = mortar_coarse.find(slot);
    return it == mortar_coarse.end() ? -1 : it->second/nf;
  };
  auto link = [&](int s0, int s1) {
    int a = elem_of(s0), b = elem_of(s1);
    if (a >= 0 && b >= 0 && a != b) {adj[a].push_back(b); adj[b].push_back(a);}
  };
  for (size_t i = 0; i < g.car_con.size()/3; ++i) link(g.car_con[i*3], g.car_con[i*3 + 1]);
  for (size_t i = 0; i < g.def_con.size()/7; ++i) link(g.def_con[i*7], g.def_con[i*7 + 1]);
  std::vector<std::uint64_t> rank_in_order(ne, 0);
  std::vector<char> seen(ne, 0);
  std::uint64_t next = 0;
  for (int seed = 0; seed < ne; ++seed) { // one sweep per connected component
    if (seen[seed]) continue;
    std::queue<int> q;
    q.push(seed); seen[seed] = 1;
    while (!q.empty()) {
      int e = q.front(); q.pop();
      rank_in_order[e] = next++;
      for (int nb : adj[e]) if (!seen[nb]) {seen[nb] = 1; q.push(nb);}
    }
  }
  return split_by_curve(rank_in_order, n_parts, default_weights(g.n_car, ne));
}

std::vector<Rank_mesh> partition(const Mesh_graph& g, const std::vector<int>& owner, int n_parts)
{
  const int nd = g.n_dim, nf = 2*nd, ne = g.n_car + g.n_def, n_elem_slots = nf*ne;
  if (int(owner.size()) != ne) throw std::runtime_error("hexed_b200::partition: one owner per element expected");
  for (int o : owner) if (o < 0 || o >= n_parts) throw std::runtime_error("hexed_b200::partition: owner out of range");
  const size_t n_cc = g.car_con.size()/3, n_dc = g.def_con.size()/7, n_rf = g.ref_face.size()/7;
  auto owner_of_slot = [&](int s) {return s < n_elem_slots ? owner[s/nf] : -1;};
  // which refined face a mortar slot belongs to, and the ranks every refined face lives on
  std::unordered_map<int, int> mortar_ref;
  for (size_t r = 0; r < n_rf; ++r) for (int k = 1; k < 5; ++k) if (g.ref_face[r*7 + k] >= 0) mortar_ref[g.ref_face[r*7 + k]] = int(r);
  auto ref_of = [&](int s0, int s1) {
    auto it = mortar_ref.find(s0);
    if (it != mortar_ref.end()) return it->second;
    it = mortar_ref.find(s1);
    return it == mortar_ref.end() ? -1 : it->second;
  };
  std::vector<std::vector<char>> ref_on(n_rf, std::vector<char>(n_parts, 0));
  for (size_t r = 0; r < n_rf; ++r) ref_on[r][owner_of_slot(g.ref_face[r*7])] = 1;
  auto note_fine = [&](int s0, int s1) {
    int r = ref_of(s0, s1);
    if (r < 0) return;
    int fine_slot = mortar_ref.count(s0) ? s1 : s0;
    ref_on[r][owner_of_slot(fine_slot)] = 1;
  };
  for (size_t i = 0; i < n_cc; ++i) note_fine(g.car_con[i*3], g.car_con[i*3 + 1]);
  for (size_t i = 0; i < n_dc; ++i) note_fine(g.def_con[i*7], g.def_con[i*7 + 1]);
  std::vector<char> is_boundary(n_dc, 0);
  for (int i : g.boundary_con) is_boundary[i] = 1;

  std::vector<Rank_mesh> parts(n_parts);
  std::vector<std::map<int, std::vector<int>>> halo_recv_global(n_parts); // rank -> peer -> global slots received from it
  std::vector<std::unordered_map<int, int>> local_of_elem(n_parts);
  for (int p = 0; p < n_parts; ++p) {
    Rank_mesh& m = parts[p];
    Mesh_graph& lg = m.graph;
    lg.n_dim = nd;
    for (int e = 0; e < g.n_car; ++e) if (owner[e] == p) m.global_elem.push_back(e);
    lg.n_car = int(m.global_elem.size());
    for (int e = g.n_car; e < ne; ++e) if (owner[e] == p) m.global_elem.push_back(e);
    lg.n_def = int(m.global_elem.size()) - lg.n_car;
    const int n_local = int(m.global_elem.size());
    auto& local_of = local_of_elem[p];
    for (int i = 0; i < n_local; ++i) local_of[m.global_elem[i]] = i;
    m.global_face.resize(size_t(nf)*n_local);
    for (int i = 0; i < n_local; ++i) for (int f = 0; f < nf; ++f) m.global_face[size_t(i)*nf + f] = m.global_elem[i]*nf + f;
    m.global_normal.resize(size_t(nf)*lg.n_def);
    for (int i = 0; i < lg.n_def; ++i) for (int f = 0; f < nf; ++f) m.global_normal[size_t(i)*nf + f] = (m.global_elem[lg.n_car + i] - g.n_car)*nf + f;
    std::unordered_map<int, int> extra_slots, extra_normals;
    auto lslot = [&](int s) -> int {
      if (s < n_elem_slots && owner[s/nf] == p) return local_of[s/nf]*nf + s%nf;
      auto it = extra_slots.find(s);
      if (it != extra_slots.end()) return it->second;
      int l = int(m.global_face.size());
      extra_slots[s] = l;
      m.global_face.push_back(s);
      if (s < n_elem_slots) halo_recv_global[p][owner[s/nf]].push_back(s);
      return l;
    };
    auto lnormal = [&](int s) -> int {
      if (s < nf*g.n_def) {
        int e = g.n_car + s/nf;
        if (owner[e] == p) return (local_of[e] - lg.n_car)*nf + s%nf;
      }
      auto it = extra_normals.find(s);
      if (it != extra_normals.end()) return it->second;
      int l = int(m.global_normal.size());
      extra_normals[s] = l;
      m.global_normal.push_back(s);
      return l;
    };
    // decides whether rank p keeps a connection and whether it must wait for the exchange
    auto classify = [&](int s0, int s1, bool& cut) -> bool {
      const int r = ref_of(s0, s1);
      if (r >= 0) { // fine connection of a hanging-node face: kept where the fine element or the coarse element is local
        const int fine_owner = owner_of_slot(mortar_ref.count(s0) ? s1 : s0), coarse_owner = owner_of_slot(g.ref_face[size_t(r)*7]);
        if (p != fine_owner && p != coarse_owner) return false;
        cut = fine_owner != p || coarse_owner != p; // the mortar face is only valid once the coarse face has arrived and been prolonged
        return true;
      }
      const int o0 = owner_of_slot(s0), o1 = owner_of_slot(s1);
      if (o0 < 0 || o1 < 0) { // boundary connection: belongs to its element's rank
        cut = false;
        return (o0 < 0 ? o1 : o0) == p;
      }
      if (o0 != p && o1 != p) return false;
      cut = o0 != o1;
      return true;
    };
    std::vector<std::array<int, 3>> car_int, car_cut;
    for (size_t i = 0; i < n_cc; ++i) {
      const int* c = g.car_con.data() + i*3;
      bool cut = false;
      if (!classify(c[0], c[1], cut)) continue;
      std::array<int, 3> row {lslot(c[0]), lslot(c[1]), c[2]};
      (cut ? car_cut : car_int).push_back(row);
    }
    struct Def_row {std::array<int, 7> row; int global; bool boundary;};
    std::vector<Def_row> def_int, def_cut;
    for (size_t i = 0; i < n_dc; ++i) {
      const int* c = g.def_con.data() + i*7;
      bool cut = false;
      if (!classify(c[0], c[1], cut)) continue;
      Def_row d {{lslot(c[0]), lslot(c[1]), c[2], c[3], c[4], c[5], lnormal(c[6])}, int(i), bool(is_boundary[i])};
      (cut ? def_cut : def_int).push_back(d);
    }
    for (size_t r = 0; r < n_rf; ++r) {
      if (!ref_on[r][p]) continue;
      const int* row = g.ref_face.data() + r*7;
      if (owner_of_slot(row[0]) != p) m.pre_prolong.push_back(int(lg.ref_face.size()/7));
      lg.ref_face.push_back(lslot(row[0]));
      for (int k = 1; k < 5; ++k) lg.ref_face.push_back(row[k] >= 0 ? lslot(row[k]) : -1);
      lg.ref_face.push_back(row[5]); lg.ref_face.push_back(row[6]);
    }
    for (auto* rows : {&car_int, &car_cut}) for (auto& r : *rows) lg.car_con.insert(lg.car_con.end(), r.begin(), r.end());
    for (auto* rows : {&def_int, &def_cut}) for (auto& d : *rows) {
      if (d.boundary) lg.boundary_con.push_back(int(lg.def_con.size()/7));
      lg.def_con.insert(lg.def_con.end(), d.row.begin(), d.row.end());
      m.global_def_con.push_back(d.global);
    }
    m.n_cut_car = int(car_cut.size()); m.n_cut_def = int(def_cut.size());
    lg.n_face_slot = int(m.global_face.size());
    lg.n_normal_slot = int(m.global_normal.size());
  }
  // which copy of a face is written back to the host: element faces by their owner; mortar faces by the rank that owns the coarse
  // element (the only one whose Prolong always starts from a current coarse face -- elsewhere the copy is refreshed by pre_prolong
  // just before use and may be stale in between); other connection-owned faces (boundary ghosts) by the one rank that holds them
  std::unordered_map<int, int> first_holder;
  for (int p = 0; p < n_parts; ++p) for (int s : parts[p].global_face) if (s >= n_elem_slots && !first_holder.count(s)) first_holder[s] = p;
  for (int p = 0; p < n_parts; ++p) {
    Rank_mesh& m = parts[p];
    m.face_owned.resize(m.global_face.size());
    for (size_t l = 0; l < m.global_face.size(); ++l) {
      const int s = m.global_face[l];
      if (s < n_elem_slots) m.face_owned[l] = owner[s/nf] == p;
      else {
        auto it = mortar_ref.find(s);
        m.face_owned[l] = (it != mortar_ref.end() ? owner_of_slot(g.ref_face[size_t(it->second)*7]) : first_holder[s]) == p;
      }
    }
  }
  // halo lists: what q receives from p is what p sends to q, both in ascending global slot order
  for (int p = 0; p < n_parts; ++p) for (auto& kv : halo_recv_global[p]) std::sort(kv.second.begin(), kv.second.end());
  for (int p = 0; p < n_parts; ++p) {
    Rank_mesh& m = parts[p];
    std::vector<char> is_peer(n_parts, 0);
    for (auto& kv : halo_recv_global[p]) is_peer[kv.first] = 1;
    for (int q = 0; q < n_parts; ++q) if (halo_recv_global[q].count(p)) is_peer[q] = 1;
    std::unordered_map<int, int> local_extra;
    for (size_t l = size_t(nf)*m.global_elem.size(); l < m.global_face.size(); ++l) local_extra[m.global_face[l]] = int(l);
    for (int q = 0; q < n_parts; ++q) {
      if (!is_peer[q]) continue;
      m.peers.push_back(q);
      std::vector<int> send, recv;
      auto it = halo_recv_global[q].find(p);
      if (it != halo_recv_global[q].end()) for (int s : it->second) send.push_back(local_of_elem[p][s/nf]*nf + s%nf);
      auto jt = halo_recv_global[p].find(q);
      if (jt != halo_recv_global[p].end()) for (int s : jt->second) recv.push_back(local_extra[s]);
      m.send_slots.push_back(send); m.recv_slots.push_back(recv);
    }
  }
  return parts;
}

}
