/* partition.hpp -- domain decomposition of a flattened Kernel_mesh across the GPUs of one box (device-free, integer-exact).
 *
 * The reference is single-process and hands its kernels ONE Kernel_mesh (include/Kernel_mesh.hpp:14-25); to use more than one GPU
 * behind that boundary the adapter splits the flattened tables (`Flat_tables`, adapter.hpp) itself. Rules (they keep every kernel
 * unchanged and the result identical to the undivided run):
 *   * an element belongs to exactly one rank; its 2*n_dim faces live on that rank;
 *   * a connection whose two faces belong to different ranks is REPLICATED on both with the remote face replaced by a HALO slot;
 *     both ranks compute the same flux from the same two inputs, so nothing is sent back. Cut connections are the LAST rows of
 *     `car_con` / `def_con` (`n_cut_car`, `n_cut_def`) so that the interior ones run while the exchange is in flight;
 *   * a hanging-node face (Refined_face: coarse face + mortar faces + fine connections, include/Refined_face.hpp:9-15,
 *     include/connection.hpp:218-269) is replicated on every rank that owns one of its participants; where the coarse element is
 *     remote its face arrives through a halo slot and is prolonged locally before the fine connections are evaluated (`pre_prolong`);
 *   * connection-owned storage (boundary ghosts, mortar faces, connection normals) is copied to every rank that uses it.
 * Ownership comes from a space-filling curve: Morton / Z-order over integer element coordinates when the caller supplies them
 * (`Element::nominal_position()` scaled to the finest level; a Kernel_mesh itself carries no coordinates), otherwise a breadth-first
 * ordering of the connection graph cut into contiguous equal-weight ranges.
 */
#ifndef HEXED_B200_PARTITION_HPP_
#define HEXED_B200_PARTITION_HPP_

#include <array>
#include <cstdint>
#include <vector>

namespace hexed_b200
{

//! the integer part of `Flat_tables` (adapter.hpp), so that the partitioner needs no hexed types
struct Mesh_graph
{
  int n_dim = 0, n_car = 0, n_def = 0, n_face_slot = 0, n_normal_slot = 0;
  std::vector<int> car_con;  //!< [n][3]
  std::vector<int> def_con;  //!< [n][7]
  std::vector<int> ref_face; //!< [n][7]
  std::vector<int> boundary_con; //!< rows of def_con that are boundary connections
};

struct Rank_mesh
{
  Mesh_graph graph;              //!< local tables (slots renumbered); interior connections first, cut connections last
  std::vector<int> global_elem;  //!< local element -> global element (Cartesian first, then deformed, each in global order)
  std::vector<int> global_face;  //!< local face slot -> global face slot
  std::vector<int> global_normal; //!< local normal slot -> global normal slot
  std::vector<char> face_owned;  //!< local face slot: this rank's copy is the one written back to the host (halo and replicated copies are not)
  int n_cut_car = 0, n_cut_def = 0;
  std::vector<int> pre_prolong;  //!< rows of graph.ref_face whose coarse face is a halo slot
  std::vector<int> peers;        //!< ranks this one exchanges with, ascending
  std::vector<std::vector<int>> send_slots, recv_slots; //!< per peer: local face slots sent / halo slots filled, both ordered by global slot
  std::vector<int> global_def_con; //!< local def_con row -> global def_con row
};

//! Z-order key of integer coordinates (up to 21 bits per dimension), last dimension least significant
std::uint64_t morton_key(const std::array<int, 3>& index, int n_dim);
//! contiguous equal-weight ranges along a curve: owner per element (weights: 1 Cartesian / 1.4 deformed when empty and `n_car` given)
std::vector<int> split_by_curve(const std::vector<std::uint64_t>& keys, int n_parts, const std::vector<double>& weights = {});
std::vector<int> owners_by_morton(const std::vector<std::array<int, 3>>& coords, int n_dim, int n_car, int n_parts);
//! no coordinates: breadth-first ordering of the element connection graph, cut into contiguous equal-weight ranges
std::vector<int> owners_by_graph(const Mesh_graph& g, int n_parts);

std::vector<Rank_mesh> partition(const Mesh_graph& g, const std::vector<int>& owner, int n_parts);

}
#endif
