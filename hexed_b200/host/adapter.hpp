/* adapter.hpp -- what the B200 replacement of Hexed's kernel drivers offers BEYOND include/kernels.hpp.
 *
 * adapter.cpp defines the reference's own free functions (hexed::compute_euler(Kernel_mesh, Kernel_options), ...; same
 * signatures as include/kernels.hpp:22-42 and include/stabilizing_art_visc.hpp:13), so swapping it for
 * src/kernels_convective.cpp, src/kernels_diffusive.cpp, src/kernels_max_dt.cpp and src/stabilizing_art_visc.cpp is the whole
 * integration. The reference API has no begin/end-of-epoch call and its callers read and write element and
 * face storage between kernel calls, so coherence is a policy the caller picks here:
 *
 *   sync_every_call (default)  every entry point uploads what it reads from the host objects and downloads what it wrote:
 *                              always correct, no Solver change, PCIe-bound.
 *   resident                   the device copy is authoritative between calls; entry points move nothing except their return
 *                              value. The caller moves data explicitly with to_host / to_device (whole groups) or
 *                              boundary_faces_to_host / ghost_faces_to_device (the per-stage traffic of host-applied
 *                              boundary conditions, src/Solver.cpp:56-81).
 */
#ifndef HEXED_B200_ADAPTER_HPP_
#define HEXED_B200_ADAPTER_HPP_

#include <array>
#include <string>
#include <vector>
#ifdef HEXED_B200_WITH_HEXED_HEADERS
#include <kernels.hpp>
#include <stabilizing_art_visc.hpp>
#else
#include "hexed_standin.hpp"
#endif

namespace hexed_b200
{

enum Sync_mode {sync_every_call = 0, resident = 1};

//! groups of host data (bit mask) for to_host / to_device
enum Data_group : unsigned {
  state = 1,       //!< element slots [0, n_var)
  tss = 2,         //!< time_step_scale()
  art_visc = 4,    //!< bulk + laplacian AV coefficient and the 4 AV forcing slots
  advection = 8,   //!< the row_size advection-state slots
  res_cache = 16,  //!< residual_cache()
  faces = 32,      //!< face(i, false) and face(i, true) of every face slot (elements, ghosts, mortar faces)
  faces_wide = 64, //!< the same storage viewed as (n_dim + row_size) variables (pde::Advection)
  geometry = 128,  //!< normals, Jacobian determinant, face normals, nominal size, vertex time step scale (to_device only)
  uncert = 256,    //!< Kernel_element::uncert() (to_host only)
  all_elem = state | tss | art_visc | advection | res_cache,
  everything = all_elem | faces | uncert
};

void set_sync_mode(Sync_mode);
Sync_mode sync_mode();
void set_device(int cuda_device); //!< run on this one device (default 0); same as set_devices({cuda_device})
/*! \brief run every Kernel_mesh on SEVERAL GPUs of the box, below the kernels.hpp boundary: the Solver still hands
 * over one Kernel_mesh and still is one process. The flattened mesh is split by a space-filling curve (Morton order of the integer
 * element coordinates given with `set_element_coordinates`; without coordinates a breadth-first ordering of the connection graph),
 * each device gets a self-contained sub-mesh whose cut faces are halo slots, and `compute_euler` / `compute_navier_stokes` /
 * `max_dt_*` run as: NCCL send/recv of the cut faces overlapped with the interior Neighbor kernels, ncclAllReduce(min) of the time
 * step (include/hexed_b200.h "device group", hexed_b200/csrc/group.cu). Results are those of one device (same kernels, same operands).
 * Entry points without a multi-device version (the artificial-viscosity PDE drivers, update_euler) throw when more than one device is set. */
void set_devices(const std::vector<int>& cuda_devices);
int n_devices();
/*! integer coordinates of every element of `Kernel_mesh::elems` in units of the finest element size (e.g. `Element::nominal_position()`
 * shifted to its refinement level), used for the Morton split. Optional. Must be called before the first kernel call of the epoch. */
void set_element_coordinates(hexed::Kernel_mesh, const std::vector<std::array<int, 3>>& coords);
std::vector<int> element_owners(hexed::Kernel_mesh); //!< device rank of every element of `Kernel_mesh::elems` (all 0 on one device)
std::string transport_description(hexed::Kernel_mesh); //!< e.g. "NCCL 2.27.3 send/recv over 8 devices, ..."
/*! forget the flattened mesh: the next entry point re-walks the Sequence views (call after mesh adaptation).
 * sync_every_call mode needs no such call: it fingerprints every element's and connection's storage address on each call and re-reads
 * the metric terms with the state, so in-place edits (`calc_jacobian`, vertex relaxation, `snap_faces`) and same-size mesh changes are
 * picked up by themselves. resident mode samples three objects per view; after a mesh change there, call to_host() first, change the
 * mesh, then invalidate() -- a changed fingerprint without invalidate() throws rather than overwrite device-only results. Metric terms
 * rewritten in place in resident mode travel with to_device(mesh, geometry). */
void invalidate();
void release(); //!< destroy every device context (also done at exit)

void to_host(hexed::Kernel_mesh, unsigned groups = everything);
void to_device(hexed::Kernel_mesh, unsigned groups = everything | geometry);
void boundary_faces_to_host(hexed::Kernel_mesh);  //!< both sides of every boundary connection, state + LDG halves
void ghost_faces_to_device(hexed::Kernel_mesh);   //!< the same faces back
/*! \brief the same traffic, only as much of it as a boundary-condition loop needs, and asynchronous where that pays.
 * `Solver::apply_state_bcs` (src/Solver.cpp:56-67) reads `inside_face(false)` and writes `ghost_face(false)`: use
 * `boundary_faces_to_host(mesh, inside, state_half)` before the loop and `ghost_faces_to_device(mesh, ghost, state_half)` after it. In
 * resident mode the download of the inside faces is then STARTED by every stage driver as soon as its kernels are enqueued (pinned
 * memory, copy stream) and only collected here, and the upload of the ghost faces is started here and lands in the face storage
 * inside the next stage driver after its Neighbor kernels on the non-boundary connections, i.e. both PCIe trips overlap device work.
 * `apply_flux_bcs` (:69-81) uses the LDG halves: (inside, ldg_half) / (ghost, ldg_half). */
enum Face_side : unsigned {inside = 1, ghost = 2, both_sides = 3};
enum Face_half : unsigned {state_half = 1, ldg_half = 2, both_halves = 3};
void boundary_faces_to_host(hexed::Kernel_mesh, unsigned sides, unsigned halves);
void ghost_faces_to_device(hexed::Kernel_mesh, unsigned sides, unsigned halves);
void synchronize(hexed::Kernel_mesh);
//! host threads of the adapter's own copy loops (0 = every hardware thread, the default; the OpenMP environment is deliberately not consulted)
void set_host_threads(int n);

/*! \brief device-resident boundary conditions: removes the per-stage PCIe round trip of the boundary faces.
 * \details `kind` / `params` as in `hexed_b200_bc_create` (include/hexed_b200.h: Freestream, Copy, Nonpenetration, Outflow,
 * Pressure_outflow, No_slip, Riemann_invariants); `inside_faces[i]` = `Boundary_connection::inside_face(false)` of the i-th face the condition applies to
 * (= `Kernel_connection::state(0, false)` of its boundary connection). Registrations belong to the current mesh epoch.
 * `apply_state_bcs` / `apply_flux_bcs` then stand in for the loops of `Solver::apply_state_bcs` / `apply_flux_bcs`
 * (src/Solver.cpp:56-81) for the registered faces. */
int add_device_bc(hexed::Kernel_mesh, int kind, const std::vector<double*>& inside_faces, const std::vector<double>& params);
void set_device_bc_params(hexed::Kernel_mesh, int id, const std::vector<double>& params); //!< e.g. a freestream state that changes between iterations
void apply_state_bcs(hexed::Kernel_mesh);
void apply_flux_bcs(hexed::Kernel_mesh);

/*! \brief the pointwise loops `Solver::update_art_visc_smoothness` (src/Solver.cpp:457-581), `fix_admissibility` (:1021-1038) and
 * `set_art_visc_admis` (:636-658) wrap around the kernels, on the device, so those pipelines run in `resident` mode
 * without moving the state: see include/hexed_b200.h for the arithmetic of each. `apply_aux_bcs(mesh, HEXED_B200_BC_MODE_*)` stands in
 * for the `flow_bc->apply_advection` loop (:505-510), `apply_avc_diff_bcs` / `apply_avc_diff_flux_bcs` (:83-101) and `apply_fta_flux_bcs`
 * (:103-115) for the boundary conditions registered with `add_device_bc`. */
void av_scale_velocity(hexed::Kernel_mesh, bool restore);
void av_project_forcing(hexed::Kernel_mesh);
double av_finish(hexed::Kernel_mesh, double mult, double us_max, int n_real); //!< returns `art_visc_residual`
void interp_vertices(hexed::Kernel_mesh, int target, const std::vector<double>& vertex_values); //!< target 0 bulk / 1 laplacian AV coefficient
void av_swap(hexed::Kernel_mesh);
/*! \brief `Solver::update_art_visc_elwise` (src/Solver.cpp:584-633) after its `set_uncertainty` call. `av_elwise_ramp` sends `Element::uncertainty`
 * up, applies the ramp of :590-601 (`scale` = width/(row_size - 1)*(freestream_speed + freestream_sound_speed)) and writes it back to the host
 * objects as the reference leaves it; `av_elwise_forcing(false)` / `(true)` are the point loops before / after `diffuse_art_visc` in the
 * PDE-based branch (:603-619); `av_elwise_vertices` is the vertex-based branch (:620-632; needs the vertex topology, single device) and, after
 * `hexed::stabilizing_art_visc`, all that is left of `Solver::set_art_visc_admis` (:636-658). */
/*! \brief the vertex connectivity `Solver::share_vertex_data` walks (src/Solver.cpp:35-54), which a Kernel_mesh does not carry: `elem_vertex`
 * [n_elem][2^n_dim] = an id in [0, n_vertex) for `Element::vertex(i)` of every element in `Kernel_mesh::elems` order, `matchers` [n][8] = one row
 * {i_dim, is_positive, stretch0, stretch1, fine element 0..3 (-1 unused)} per `Hanging_vertex_matcher`. Belongs to the current mesh epoch; single device. */
void vertex_topology(hexed::Kernel_mesh, const std::vector<int>& elem_vertex, int n_vertex, const std::vector<int>& matchers);
void av_elwise_ramp(hexed::Kernel_mesh, double scale);
void av_elwise_forcing(hexed::Kernel_mesh, bool restore);
void av_elwise_vertices(hexed::Kernel_mesh);
void apply_aux_bcs(hexed::Kernel_mesh, int mode);

/*! \brief `n_steps` passes of the flow loop of `Solver::update` (src/Solver.cpp:834-886) for the inviscid case (`n_cheby` = `n_cheby_flow`) and
 * every boundary condition registered with `add_device_bc`: `max_dt_euler(safety)` + two stages, each `apply_state_bcs` + `compute_euler`.
 * The time step never leaves the device and the step is replayed from a CUDA graph (`use_graph`), which is what matters on the small
 * meshes of the 2-D sample cases. Returns the flow time advanced (add it to `status.flow_time`); `last_dt` receives `status.time_step`. */
double update_euler(hexed::Kernel_mesh, double safety, int n_steps, double* last_dt = nullptr, bool use_graph = true, int n_cheby = 1);

/*! \brief `Solver::is_admissible` (src/Solver.cpp:921-958) on the device: the check Solver::update makes after every
 * stage. Returns what the reference returns; `record`, if given, receives `Element::record` of every element in `Kernel_mesh::elems` order
 * (1 = inadmissible), which `fix_admissibility` spreads to the vertices. Throws `std::runtime_error("state is not finite")` like the
 * reference's HEXED_ASSERT. One 8-byte read-back instead of a pass over the whole state on the host. */
bool is_admissible(hexed::Kernel_mesh, std::vector<int>* record = nullptr);

/*! \brief the pointer graph of a Kernel_mesh turned into slot tables (layout: include/hexed_b200.h). Device-free. */
struct Flat_tables
{
  int n_dim = 0, row_size = 0;
  int n_car = 0, n_def = 0, n_face_slot = 0, n_normal_slot = 0;
  std::vector<int> car_con;  //!< [n][3]
  std::vector<int> def_con;  //!< [n][7]
  std::vector<int> ref_face; //!< [n][7]
  std::vector<hexed::Kernel_element*> elem; //!< car_elems then def_elems
  std::vector<double*> face_ptr;   //!< [n_face_slot] host address of `face(i, false)` / `state(side, false)`; nullptr = unconnected
  std::vector<double*> normal_ptr; //!< [n_normal_slot] host address; nullptr = unit normal of the face's dimension
  std::vector<int> boundary_con;   //!< indices into def_con of the boundary connections (side 1 = ghost face of no element)
};
Flat_tables flatten(hexed::Kernel_mesh);

}
#endif
