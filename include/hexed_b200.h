/* hexed_b200.h -- C ABI of the B200-native implementation of Hexed's per-stage DG residual update.
 *
 * This is the drop-in boundary for the hot path behind `hexed::Solver::update`: the functions below are what a
 * replacement of the reference's four kernel-driver files (src/kernels_convective.cpp, src/kernels_diffusive.cpp,
 * src/kernels_max_dt.cpp, src/stabilizing_art_visc.cpp) binds to. Each compute entry point cites the reference
 * declaration it replaces (paths relative to the Hexed source tree). The C++ adapter that implements the genuine
 * `hexed::compute_euler(Kernel_mesh, Kernel_options)` signatures on top of this ABI is hexed_b200/host/ (see
 * INTEGRATION.md).
 *
 * Conventions
 *   - plain pointers and sizes only; every function returns 0 on success or a HEXED_B200_* error code, and
 *     `hexed_b200_last_error` gives the message. The C++ adapter turns code 1 into
 *     std::runtime_error("demand for invalid kernel") exactly like include/kernel_factory.hpp:114-116.
 *   - there is NO CPU fallback: without a CUDA device every compute call fails with HEXED_B200_NO_DEVICE.
 *   - data pointers passed to upload/download calls may be host (pageable or pinned) or device memory.
 *   - all floating point data is IEEE double; all index tables are 32-bit int.
 *
 * Flattened mesh (one "mesh epoch"; the reference's pointer graph include/connection.hpp:111-123 turned into slots)
 *   nq = row_size^n_dim, nfq = row_size^(n_dim-1), nv = n_dim + 2
 *   elements [0, n_car) Cartesian, [n_car, n_car + n_def) deformed (order of Kernel_mesh::car_elems / def_elems)
 *   element "slots" of nq doubles, reference order (src/Element.cpp:114-142,187-189, src/Storage_params.cpp:32-35):
 *       state nv | time-step scale 1 | bulk AV 1 | laplacian AV 1 | AV forcing 4 | advection row_size | residual cache max(nv,row_size)
 *   face slot of element e face f (= 2*i_dim + sign): e*2*n_dim + f; connection-owned faces (boundary ghosts, hanging-node
 *       mortar faces) use slots >= 2*n_dim*n_elem.  A face slot holds [nv][nfq] doubles per "kind":
 *       kind 0 = `face(i, false)` / `state(side, false)`, kind 1 = the LDG half `(…, true)`,
 *       kind 2 = the (n_dim+row_size)-variable view used by pde::Advection.
 *   normal slot of deformed element d (= e - n_car) face f: d*2*n_dim + f, [n_dim][nfq] doubles
 *       (`Kernel_element::kernel_face_normal`, unit normal when the reference returns nullptr, include/Spatial.hpp:366);
 *       connection-owned normals (`Kernel_connection::normal()`) that alias no element face use later slots.
 *   car_con[n][3] = {slot side 0, slot side 1, i_dim}
 *   def_con[n][7] = {slot side 0, slot side 1, i_dim0, i_dim1, face_sign0, face_sign1, normal slot}
 *       (`Connection_direction`, include/Kernel_connection.hpp:7-37; boundary connections are included, as in
 *        src/Accessible_mesh.cpp:136-147)
 *   ref_face[n][7] = {coarse slot, fine slot 0..3 (-1 = unused), stretch0, stretch1}   (include/Refined_face.hpp:9-15)
 *
 * Basis tables (argument `basis` of hexed_b200_create): packed doubles, rs = row_size, matrices row-major M[i][j]
 *   node[rs] weight[rs] diff_mat[rs][rs] boundary[2][rs] orthogonal[rs][rs] filter[rs][rs] prolong[2][rs][rs]
 *   restrict[2][rs][rs] min_eig_convection min_eig_diffusion quadratic_safety      (include/Basis.hpp:16-66)
 *   legendre_node[rs]   nodes of Gauss_legendre(row_size), which pde::Advection uses whatever the basis (include/pde.hpp:281)
 */
#ifndef HEXED_B200_H_
#define HEXED_B200_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct hexed_b200_ctx hexed_b200_ctx;

enum {
  HEXED_B200_OK = 0,
  HEXED_B200_INVALID_KERNEL = 1, /* n_dim not in 1..3 or row_size not in 2..8: "demand for invalid kernel" */
  HEXED_B200_NO_DEVICE = 2,
  HEXED_B200_CUDA_ERROR = 3,
  HEXED_B200_BAD_ARGUMENT = 4,
  HEXED_B200_NO_MESH = 5,
  HEXED_B200_NOT_IMPLEMENTED = 6,
  HEXED_B200_NOT_FINITE = 7      /* a state value is NaN/Inf: the reference's HEXED_ASSERT("state is not finite"), src/thermo.cpp:14 */
};

/* element arrays addressable by hexed_b200_upload / hexed_b200_download (item = one element unless noted) */
enum {
  HEXED_B200_NOMINAL_SIZE = 0, /* [n_elem]                 Kernel_element::nominal_size() */
  HEXED_B200_VERTEX_TSS = 1,   /* [n_elem][2^n_dim]        Kernel_element::vertex_time_step_scale(i) */
  HEXED_B200_REF_NORMALS = 2,  /* [n_def][n_dim*n_dim][nq] Kernel_element::reference_level_normals() */
  HEXED_B200_JAC_DET = 3,      /* [n_def][nq]              Kernel_element::jacobian_determinant() */
  HEXED_B200_FACE_STATE = 4,   /* [n_face_slot][nv*nfq]    item = face slot */
  HEXED_B200_FACE_LDG = 5,     /* [n_face_slot][nv*nfq]    item = face slot */
  HEXED_B200_FACE_WIDE = 6,    /* [n_face_slot][(n_dim+row_size)*nfq] item = face slot */
  HEXED_B200_NORMALS = 7,      /* [n_normal_slot][n_dim*nfq] item = normal slot */
  HEXED_B200_UNCERT = 8,       /* [n_elem]                 Kernel_element::uncert() */
  HEXED_B200_VERTEX_SCRATCH = 9 /* [n_elem][2^n_dim]       Element::vertex_fix_admis_coef(i) / vertex_elwise_av(i) */
};

enum { HEXED_B200_BC_FREESTREAM = 0, HEXED_B200_BC_COPY = 1, HEXED_B200_BC_NONPENETRATION = 2,
       HEXED_B200_BC_OUTFLOW = 3, HEXED_B200_BC_PRESSURE_OUTFLOW = 4, HEXED_B200_BC_NO_SLIP = 5,
       HEXED_B200_BC_RIEMANN_INVARIANTS = 6 };

typedef struct {
  int n_car, n_def;
  int n_face_slot, n_normal_slot;
  int n_car_con, n_def_con, n_ref;
  const int* car_con;
  const int* def_con;
  const int* ref_face;
} hexed_b200_mesh_desc;

/* mirrors hexed::Kernel_options minus the stopwatches (include/kernels.hpp:11-20) */
typedef struct {
  double dt;
  int i_stage;
  int compute_residual;
  int use_filter;
} hexed_b200_options;

/* the five doubles + flag of hexed::Transport_model (include/Transport_model.hpp:16-31) */
typedef struct {
  double const_val, ref_val, ref_temp, sqrt_ref_temp, temp_offset;
  int is_viscous;
} hexed_b200_transport;

typedef void (*hexed_b200_callback)(void* user);

/* per-kernel work-unit counters and device time, keeps the reference's Stopwatch_tree side-contract alive
 * (include/kernel_factory.hpp:32-45); names match the reference's children: "neighbor", "local",
 * "reconcile LDG flux", "compute time step", "prolong/restrict", "boundary conditions" */
typedef struct {
  const char* name;
  int deformed;           /* 0 = cartesian tree, 1 = deformed tree, 2 = prolong/restrict / other */
  long long work_units;
  long long launches;
  double device_seconds;  /* only accumulated while timing is enabled */
} hexed_b200_kernel_stat;

/* ---- life cycle ---- */
int hexed_b200_device_count(int* count);
int hexed_b200_create(hexed_b200_ctx** ctx, int device, int n_dim, int row_size, const double* basis, int n_basis);
int hexed_b200_destroy(hexed_b200_ctx* ctx);
const char* hexed_b200_last_error(const hexed_b200_ctx* ctx); /* ctx may be NULL: error of the last failed create */
int hexed_b200_synchronize(hexed_b200_ctx* ctx);
int hexed_b200_cuda_stream(hexed_b200_ctx* ctx, void** stream); /* the cudaStream_t all kernels of this context run on */

/* ---- mesh epoch: (re)builds the device mirror; previous mesh storage is released ---- */
int hexed_b200_mesh_create(hexed_b200_ctx* ctx, const hexed_b200_mesh_desc* desc);
int hexed_b200_upload(hexed_b200_ctx* ctx, int which, const double* src, size_t first_item, size_t n_items);
int hexed_b200_download(hexed_b200_ctx* ctx, int which, double* dst, size_t first_item, size_t n_items);
/* element slots in the reference's per-element layout; `elem_stride` = doubles between consecutive elements in the host array */
int hexed_b200_upload_elem_slots(hexed_b200_ctx* ctx, const double* src, size_t elem_stride, int first_slot, int n_slots, int first_elem, int n_elem);
int hexed_b200_download_elem_slots(hexed_b200_ctx* ctx, double* dst, size_t elem_stride, int first_slot, int n_slots, int first_elem, int n_elem);
/* registered lists of face slots (e.g. all boundary faces) moved as one packed block [n][width(kind)] */
int hexed_b200_face_list_create(hexed_b200_ctx* ctx, const int* slots, int n, int* list_id);
int hexed_b200_face_list_download(hexed_b200_ctx* ctx, int list_id, int kind, double* dst);
int hexed_b200_face_list_upload(hexed_b200_ctx* ctx, int list_id, int kind, const double* src);
/* asynchronous variants for the per-stage traffic of HOST-applied boundary conditions (Solver::apply_state_bcs, src/Solver.cpp:56-67):
 *   prefetch         gather the list's faces and start their device-to-host copy into a pinned buffer of the list, on a copy stream
 *                    (call it right after the stage that produced the faces; returns at once)
 *   prefetched       wait for that copy (starting it now if none is in flight) and return the pinned buffer [n][width(kind)]
 *   staging          the list's pinned upload buffer [n][widest kind]: the caller writes the faces there ...
 *   upload_deferred  ... and this starts the host-to-device copy at once; the faces are scattered into the face storage inside the next
 *                    stage driver AFTER its Neighbor kernels on the connections not declared late by hexed_b200_set_partition (declare
 *                    the boundary connections late and the copy overlaps the interior flux work). Any other entry point that could read
 *                    the faces completes the upload first. */
int hexed_b200_face_list_prefetch(hexed_b200_ctx* ctx, int list_id, int kind);
int hexed_b200_face_list_prefetched(hexed_b200_ctx* ctx, int list_id, int kind, const double** host);
int hexed_b200_face_list_staging(hexed_b200_ctx* ctx, int list_id, double** host);
int hexed_b200_face_list_upload_deferred(hexed_b200_ctx* ctx, int list_id, int kind);
/* integer table the kernels use for `Face_permutation::match_faces` (include/Spatial.hpp:85-129): out[nfq] */
int hexed_b200_face_permutation_table(hexed_b200_ctx* ctx, const int dir[4], int* out);
/* the same table without a context or a device (pure integer host logic, used by the C++ adapter's `face_permutation`):
 * dir = {i_dim0, i_dim1, face_sign0, face_sign1}; out[row_size^(n_dim-1)]; matched[p] = original[out[p]] */
int hexed_b200_face_permutation_indices(int n_dim, int row_size, const int dir[4], int* out);

/* ---- stage drivers ---- */
/* void compute_euler(Kernel_mesh, Kernel_options)                                  include/kernels.hpp:22, src/kernels_convective.cpp:18 */
int hexed_b200_compute_euler(hexed_b200_ctx* ctx, hexed_b200_options opts);
/* double max_dt_euler(Kernel_mesh, Kernel_options, double, double, bool)           include/kernels.hpp:29, src/kernels_max_dt.cpp:14 */
/* The five max_dt_* assume an ADMISSIBLE state (positive density and energy, finite), which Solver::update guarantees by calling
 * is_admissible after every stage: the reduction keeps positive finite local time steps only, so a negative or NaN local value -- which the
 * reference's std::min would carry through as an obviously wrong dt -- is ignored here. Call hexed_b200_is_admissible first where that matters. */
int hexed_b200_max_dt_euler(hexed_b200_ctx* ctx, hexed_b200_options opts, double convective_safety, double diffusive_safety, int local_time, double* dt);
/* void compute_write_face(Kernel_mesh)                                             include/kernels.hpp:40, src/kernels_convective.cpp:43-46 */
int hexed_b200_compute_write_face(hexed_b200_ctx* ctx);
/* void compute_prolong(Kernel_mesh, bool scale, bool offset)                       include/kernels.hpp:36, src/kernels_convective.cpp:23-26 */
int hexed_b200_compute_prolong(hexed_b200_ctx* ctx, int scale, int offset);
/* void compute_restrict(Kernel_mesh, bool scale, bool offset)                      include/kernels.hpp:37, src/kernels_convective.cpp:28-31 */
int hexed_b200_compute_restrict(hexed_b200_ctx* ctx, int scale, int offset);
/* std::unique_ptr<Face_permutation_dynamic> face_permutation(int, int, Connection_direction, double*)
 *                                                                                  include/kernels.hpp:39, src/kernels_convective.cpp:38-41
 * applies match_faces (restore = 0) or restore (restore = 1) to nv variables of one face held in `data` */
int hexed_b200_face_permutation(hexed_b200_ctx* ctx, const int dir[4], int restore, double* data);

/* void compute_advection(Kernel_mesh, Kernel_options, double advect_length)        include/kernels.hpp:23, src/kernels_convective.cpp:19 */
int hexed_b200_compute_advection(hexed_b200_ctx* ctx, hexed_b200_options opts, double advect_length);
/* void compute_navier_stokes(Kernel_mesh, Kernel_options, std::function<void()> flux_bc, Transport_model visc, Transport_model therm_cond)
 *                                                                                  include/kernels.hpp:24-25, src/kernels_diffusive.cpp:28-29
 * `flux_bc(user)` is called on the calling thread between Prolong and Neighbor_reconcile of stage 0 (src/kernels_diffusive.cpp:18);
 * it may be NULL, and it may call hexed_b200_apply_flux_bcs / face_list transfers on this context. */
int hexed_b200_compute_navier_stokes(hexed_b200_ctx* ctx, hexed_b200_options opts, hexed_b200_callback flux_bc, void* user,
                                     hexed_b200_transport visc, hexed_b200_transport therm_cond);
/* void compute_smooth_av(Kernel_mesh, Kernel_options, std::function<void()> flux_bc, double diff_time, double chebyshev_step)
 *                                                                                  include/kernels.hpp:26, src/kernels_diffusive.cpp:30-31 */
int hexed_b200_compute_smooth_av(hexed_b200_ctx* ctx, hexed_b200_options opts, hexed_b200_callback flux_bc, void* user, double diff_time, double chebyshev_step);
/* void compute_fix_therm_admis(Kernel_mesh, Kernel_options, std::function<void()> flux_bc)   include/kernels.hpp:27, src/kernels_diffusive.cpp:32 */
int hexed_b200_compute_fix_therm_admis(hexed_b200_ctx* ctx, hexed_b200_options opts, hexed_b200_callback flux_bc, void* user);
/* double max_dt_navier_stokes / _advection / _smooth_av / _fix_therm_admis(...)    include/kernels.hpp:30-34, src/kernels_max_dt.cpp:15-21 */
int hexed_b200_max_dt_navier_stokes(hexed_b200_ctx* ctx, hexed_b200_options opts, double convective_safety, double diffusive_safety, int local_time,
                                    hexed_b200_transport visc, hexed_b200_transport therm_cond, double* dt);
int hexed_b200_max_dt_advection(hexed_b200_ctx* ctx, hexed_b200_options opts, double convective_safety, double diffusive_safety, int local_time, double advect_length, double* dt);
int hexed_b200_max_dt_smooth_av(hexed_b200_ctx* ctx, hexed_b200_options opts, double convective_safety, double diffusive_safety, int local_time, double* dt);
int hexed_b200_max_dt_fix_therm_admis(hexed_b200_ctx* ctx, hexed_b200_options opts, double convective_safety, double diffusive_safety, int local_time, double* dt);
/* void compute_prolong_advection(Kernel_mesh)                                      include/kernels.hpp:38, src/kernels_convective.cpp:33-36 */
int hexed_b200_compute_prolong_advection(hexed_b200_ctx* ctx);
/* void compute_write_face_advection / _smooth_av(Kernel_mesh)                      include/kernels.hpp:41-42, src/kernels_convective.cpp:48-56 */
int hexed_b200_compute_write_face_advection(hexed_b200_ctx* ctx);
int hexed_b200_compute_write_face_smooth_av(hexed_b200_ctx* ctx);
/* void stabilizing_art_visc(Kernel_mesh, double char_speed)                        include/stabilizing_art_visc.hpp:13, src/stabilizing_art_visc.cpp:8-66
 * result in the HEXED_B200_UNCERT array */
int hexed_b200_stabilizing_art_visc(hexed_b200_ctx* ctx, double char_speed);

/* individual kernels of the other PDEs, for unit-level parity. pde: 1 Navier-Stokes, 2 advection, 3 smooth AV, 4 fix therm admis;
 * which: 0 Neighbor, 1 Local, 2 Neighbor_reconcile, 3 Reconcile_ldg_flux (include/Spatial.hpp:613-704,326-509,716-759,543-594);
 * p0, p1: advect_length | diff_time, chebyshev_step */
int hexed_b200_pde_kernel(hexed_b200_ctx* ctx, int pde, int which, int deformed, hexed_b200_options opts,
                          hexed_b200_transport visc, hexed_b200_transport therm_cond, double p0, double p1);

/* ---- individual kernels of the Euler sequence (for unit-level parity; deformed: 0 = Cartesian set, 1 = deformed set) ---- */
int hexed_b200_neighbor_euler(hexed_b200_ctx* ctx, int deformed);   /* Spatial<..>::Neighbor   include/Spatial.hpp:613-704 */
int hexed_b200_local_euler(hexed_b200_ctx* ctx, int deformed, hexed_b200_options opts); /* Spatial<..>::Local include/Spatial.hpp:326-509 */

/* ---- device-resident ghost-state boundary conditions (Solver::apply_state_bcs, src/Solver.cpp:56-67) ---- */
/* Freestream src/Boundary_condition.cpp:66-76 (params = nv doubles), Copy :450-453, Nonpenetration :301-327, Outflow :465-477,
 * Pressure_outflow :184-223 (params = {specified pressure}), No_slip :367-418 (params = {thermal kind, a, b, c, heat flux coercion,
 * stefan_boltzmann}: kind 0 Prescribed_heat_flux(a), 1 Prescribed_energy(a), 2 Thermal_equilibrium(emissivity a, heat transfer
 * coefficient b, temperature c), include/Boundary_condition.hpp:140-200; its state cache lives on the device),
 * Riemann_invariants :97-182 (params = nv doubles of freestream state; characteristic decomposition include/pde.hpp:181-256 with the
 * 3x3 column-pivoted Householder QR done per face point in registers; state cache on the device) */
int hexed_b200_bc_create(hexed_b200_ctx* ctx, int kind, int n, const int* inside_slot, const int* ghost_slot,
                         const int* normal_slot, const double* params, int n_params, int* bc_id);
/* new values for the parameter block of a registered condition, e.g. a freestream state that HIL changes between iterations
 * (`Freestream::fs`, include/Boundary_condition.hpp): stream-ordered, read by every apply_*_bcs enqueued after this call */
int hexed_b200_bc_set_params(hexed_b200_ctx* ctx, int bc_id, const double* params, int n_params);
int hexed_b200_apply_state_bcs(hexed_b200_ctx* ctx);
/* Solver::apply_flux_bcs (src/Solver.cpp:69-81) for the same boundary conditions: Freestream/Copy::apply_flux = copy_state
 * (src/Boundary_condition.cpp:12-23,304-305,455-458), Nonpenetration::apply_flux (:329-341). The flux cache copy of :75-76 is host-side. */
int hexed_b200_apply_flux_bcs(hexed_b200_ctx* ctx);

/* ---- the flow loop of Solver::update (src/Solver.cpp:834-886) with every boundary condition registered on the device:
 * n_steps x [nominal_dt = max_dt(safety/max_cheby, safety); dt = nominal_dt*chebyshev_step(n_cheby, i_cheby); 2 x (apply_state_bcs + stage)],
 * i_cheby cycling through 0..n_cheby-1 (n_cheby = n_cheby_flow, default 1), with the time step kept on the device (no synchronisation per
 * step); with use_graph one Chebyshev cycle is captured in a CUDA graph and replayed. Bit-identical to the same calls made one by one.
 * Returns the last time step and the flow time advanced. For launch-bound (small) meshes.
 * update_euler: both stages compute_euler. update_navier_stokes (use_ldg(), :857-865): stage 0 compute_navier_stokes with the flux
 * boundary conditions on the device, stage 1 compute_euler.
 * PRECONDITIONS (what the loop of the reference does and these two do not; a caller that needs one of them uses the call-by-call entry points):
 *   - no `min(nominal_dt, max_time_step)` clamp (:851): `max_time_step` must not bind;
 *   - no `fix_admissibility` after each stage and therefore no early exit from the Chebyshev cycle (:866-868): check
 *     hexed_b200_is_admissible after the call (free with HEXED_B200_OPT_FUSED_ADMIS) and repair / redo on the host side if it fails;
 *   - global time stepping (`local_time` false) and `use_filter` 0. ---- */
int hexed_b200_update_euler(hexed_b200_ctx* ctx, double safety, int n_cheby, int n_steps, int use_graph, double* last_dt, double* time_advanced);
int hexed_b200_update_navier_stokes(hexed_b200_ctx* ctx, double safety, hexed_b200_transport visc, hexed_b200_transport therm_cond,
                                    int n_cheby, int n_steps, int use_graph, double* last_dt, double* time_advanced);

/* ---- thermodynamic admissibility: Solver::is_admissible (src/Solver.cpp:921-958, src/thermo.cpp:6-18), which
 * Solver::update runs after EVERY stage through fix_admissibility (:864-868). *admissible = 1 iff mass > 0 and energy > 0 at every
 * point of every element's state, of its 2*n_dim faces and of the fine mortar faces of every refined face; the per-element result
 * (Element::record, 1 = inadmissible) stays on the device for hexed_b200_download_record. One 8-byte read-back per call.
 * Returns HEXED_B200_NOT_FINITE where the reference throws "state is not finite". ---- */
int hexed_b200_is_admissible(hexed_b200_ctx* ctx, int* admissible);
/* the same in two parts -- enqueue the check and the read-back of its flags / wait for them -- so that one host thread driving several
 * devices starts the check on all of them before it waits for any (nothing else may be enqueued on the context in between) */
int hexed_b200_is_admissible_begin(hexed_b200_ctx* ctx);
int hexed_b200_is_admissible_finish(hexed_b200_ctx* ctx, int* admissible);
int hexed_b200_download_record(hexed_b200_ctx* ctx, int* dst, int first_elem, int n_elem);
/* Solver::share_vertex_data (src/Solver.cpp:35-54): every mesh vertex takes the min (op 0) or max (op 1) over the elements that share
 * it, then the Hanging_vertex_matchers interpolate onto hanging vertices (src/Hanging_vertex_matcher.cpp:13-41). The vertex connectivity
 * is outside Kernel_mesh, so it is given once per mesh epoch: elem_vertex [n_elem][2^n_dim] = id in [0, n_vertex) of Element::vertex(i);
 * matchers [n_match][8] = {i_dim, is_positive, stretch0, stretch1, fine element 0..3 (-1 = unused)} in the matcher's element order.
 * which = HEXED_B200_VERTEX_TSS (calc_jacobian's last step, :380) or HEXED_B200_VERTEX_SCRATCH. */
int hexed_b200_vertex_topology(hexed_b200_ctx* ctx, const int* elem_vertex, int n_vertex, const int* matchers, int n_match);
int hexed_b200_share_vertex_data(hexed_b200_ctx* ctx, int which, int op);
/* the spreading step of Solver::fix_admissibility (:1000-1038): record -> element vertices, share max, element-wise max, share max,
 * interpolation to laplacian_av_coef with interp[row_size][2] = {1 - node, node}, swap with bulk_av_coef. With max_dt_fix_therm_admis,
 * apply_aux_bcs and compute_fix_therm_admis the whole repair iteration then runs without touching the host objects. */
int hexed_b200_fix_admis_spread(hexed_b200_ctx* ctx, const double* interp);

/* ---- pointwise loops around the artificial-viscosity kernels: the host loops of
 * Solver::update_art_visc_smoothness (src/Solver.cpp:457-581), fix_admissibility (:1021-1038,1080-1086), set_art_visc_admis (:636-658)
 * and update_art_visc_elwise (:625-633), so that those pipelines can run with the state resident on the device.
 *   av_scale_velocity(restore = 0 | 1)   momentum /= | *= sqrt(2*mass*energy)                      (:467-478 | :567-571)
 *   av_project_forcing(w, orth)          forcing[0] = (sum_i advection_state_i*w_i*orth_i)^2*2*energy/mass, w = Basis::node_weights(),
 *                                        orth = Basis::orthogonal(row_size - 1)                   (:527-541)
 *   av_finish(mult, us_max, n_real, w, &residual)  f = mult*forcing[n_real]; bulk_av_coef = us_max*f/(us_max + f); velocity restored;
 *                                        residual = sqrt(sum (old - new)^2 * quadrature weight * nominal_size^n_dim)  (:551-573)
 *   interp_vertices(target, vertex_values[n_elem][2^n_dim], interp[row_size][2])  math::hypercube_matvec(interp, vertex values) into
 *                                        bulk_av_coef (target 0) or laplacian_av_coef (target 1)   (:652-656, :1021-1031)
 *   av_swap()                            swap bulk_av_coef and laplacian_av_coef                    (:1032-1038) ---- */
int hexed_b200_av_scale_velocity(hexed_b200_ctx* ctx, int restore);
int hexed_b200_av_project_forcing(hexed_b200_ctx* ctx, const double* node_weights, const double* orthogonal);
int hexed_b200_av_finish(hexed_b200_ctx* ctx, double mult, double us_max, int n_real, const double* node_weights, double* residual);
int hexed_b200_interp_vertices(hexed_b200_ctx* ctx, int target, const double* vertex_values, const double* interp);
int hexed_b200_av_swap(hexed_b200_ctx* ctx);
/* Solver::update_art_visc_elwise (src/Solver.cpp:584-633) after its set_uncertainty call, on HEXED_B200_UNCERT [n_elem]:
 *   av_elwise_ramp(scale)        u <- ramp(2*log10(u)) * scale: 0 below, 1 above, half a sine period across a window of width 1 centred at
 *                                -4 - 4.25*log10(row_size - 1); scale = width/(row_size - 1)*(freestream speed + sound speed)   (:590-601)
 *   av_elwise_forcing(restore)   0: art_visc_forcing[0] = u, art_visc_forcing[1] = laplacian_av_coef (:603-612, before diffuse_art_visc);
 *                                1: laplacian_av_coef = art_visc_forcing[1] (:614-619, after it)
 *   av_elwise_vertices(interp)   the other branch (:620-632): u -> vertex_elwise_av of every vertex of the element, share_vertex_data(max)
 *                                (needs hexed_b200_vertex_topology), laplacian_av_coef = hypercube_matvec(interp[row_size][2], vertex values) */
int hexed_b200_av_elwise_ramp(hexed_b200_ctx* ctx, double scale);
int hexed_b200_av_elwise_forcing(hexed_b200_ctx* ctx, int restore);
int hexed_b200_av_elwise_vertices(hexed_b200_ctx* ctx, const double* interp);
/* the boundary loops of the same pipelines for the boundary conditions registered with hexed_b200_bc_create:
 *   MODE_ADVECTION    Flow_bc::apply_advection and its overrides (src/Boundary_condition.cpp:24-41,346-369,429-448,460-463), Solver.cpp:505-510
 *   MODE_COPY_STATE   Flow_bc::apply_diffusion (:43-52; Solver::apply_avc_diff_bcs, Solver.cpp:83-91) and the ghost copy of :1063-1068
 *   MODE_NEGATE_FLUX  Flow_bc::flux_diffusion (:54-60; Solver::apply_avc_diff_flux_bcs, Solver.cpp:93-101) and apply_fta_flux_bcs (:103-115) */
enum { HEXED_B200_BC_MODE_ADVECTION = 0, HEXED_B200_BC_MODE_COPY_STATE = 1, HEXED_B200_BC_MODE_NEGATE_FLUX = 2 };
int hexed_b200_apply_aux_bcs(hexed_b200_ctx* ctx, int mode);

/* ---- metric terms on the device: the element loop of Solver::calc_jacobian (src/Solver.cpp:281-286), i.e.
 * Deformed_element::set_jacobian (src/Deformed_element.cpp:60-136, positions by :15-58 including the face-warping node adjustments)
 * for every deformed element and Element::set_jacobian (src/Element.cpp:99-112) for every Cartesian one. Host inputs:
 * vertex_pos [n_def][2^n_dim][n_dim] (vertex index row-major, dimension 0 slowest; Vertex::pos), node_adj [n_def][2*n_dim][nfq]
 * (Deformed_element::node_adjustments(), may be NULL = all zero); nominal sizes must have been uploaded. Writes reference-level normals,
 * determinant, vertex time-step scale of the deformed elements and, like the reference, each element's face normals into the first
 * n_dim*nfq doubles of its face storage (HEXED_B200_FACE_STATE), where the shared-normal pass of calc_jacobian (:287-357, host) reads them. ---- */
int hexed_b200_set_jacobian(hexed_b200_ctx* ctx, const double* vertex_pos, const double* node_adj);
/* the connection passes of the same function (:287-369) on what set_jacobian left in the face storage: coarse normals prolonged to the
 * mortar faces, fine sides of fine connections, ghost faces, the sign-aware average of the two element normals of every deformed
 * connection (stored as Kernel_connection::normal() and as kernel_face_normal() of the deformed elements on either side), coarse
 * element face normals. The last step of calc_jacobian, share_vertex_data(vertex_time_step_scale, min) (:380), is
 * hexed_b200_share_vertex_data(ctx, HEXED_B200_VERTEX_TSS, 0); the boundary surface positions (:370-379) are host-side geometry. */
int hexed_b200_calc_shared_normals(hexed_b200_ctx* ctx);

/* ---- domain decomposition (new in this implementation: the reference is single-process) ----
 * A rank's mesh is self-contained: faces of remote elements are HALO face slots (ordinary slots >= 2*n_dim*n_elem). The host
 * registers the slots it sends / receives as face lists and moves them with face_list_gather / face_list_scatter (device
 * buffers, asynchronous on the context's stream; the transport - NCCL send/recv - is the caller's). Cut connections must be
 * the last `n_cut_car` / `n_cut_def` rows of the connection tables; `pre_prolong_ref` lists the refined faces whose coarse face
 * is a halo slot (it is prolonged onto the local mortar faces after the exchange). */
int hexed_b200_set_partition(hexed_b200_ctx* ctx, int n_cut_car, int n_cut_def, int n_pre_prolong, const int* pre_prolong_ref);
int hexed_b200_face_list_gather(hexed_b200_ctx* ctx, int list_id, int kind, double* device_dst);
int hexed_b200_face_list_scatter(hexed_b200_ctx* ctx, int list_id, int kind, const double* device_src);
/* compute_euler split around the exchange: begin = Neighbor on interior connections; finish = everything else */
int hexed_b200_compute_euler_begin(hexed_b200_ctx* ctx);
int hexed_b200_compute_euler_finish(hexed_b200_ctx* ctx, hexed_b200_options opts);

/* compute_navier_stokes split around its TWO exchanges (the viscous stage also couples ranks through the LDG faces):
 *   begin : Neighbor (numerical flux + LDG average state) on the connections that touch no halo face
 *           -- meanwhile the caller exchanges the state faces (kind 0) of the cut connections --
 *   middle: pre-prolong, Neighbor on the cut connections, Restrict x2, Local, and for stage 0 Prolong of the viscous flux, flux_bc,
 *           Neighbor_reconcile on the interior connections
 *           -- meanwhile the caller exchanges the LDG faces (kind 1) of the cut connections --
 *   finish: (stage 0) pre-prolong of the LDG halves, Neighbor_reconcile on the cut connections, Restrict, Reconcile_ldg_flux; Prolong
 * Same kernel order as src/kernels_diffusive.cpp:8-26 on every connection / element. */
int hexed_b200_compute_navier_stokes_begin(hexed_b200_ctx* ctx, hexed_b200_options opts, hexed_b200_transport visc, hexed_b200_transport therm_cond);
int hexed_b200_compute_navier_stokes_middle(hexed_b200_ctx* ctx, hexed_b200_options opts, hexed_b200_callback flux_bc, void* user,
                                            hexed_b200_transport visc, hexed_b200_transport therm_cond);
int hexed_b200_compute_navier_stokes_finish(hexed_b200_ctx* ctx, hexed_b200_options opts, hexed_b200_transport visc, hexed_b200_transport therm_cond);

/* the two halves of `middle` (before / after the flux_bc callback), and max_dt of any PDE with the result left on the device at
 * `device_out` (pde: 0 Euler, 1 Navier-Stokes, 2 advection, 3 smooth AV, 4 fix therm admis); used by the device group below */
int hexed_b200_compute_navier_stokes_middle_local(hexed_b200_ctx* ctx, hexed_b200_options opts, hexed_b200_transport visc, hexed_b200_transport therm_cond);
int hexed_b200_compute_navier_stokes_middle_reconcile(hexed_b200_ctx* ctx, hexed_b200_options opts, hexed_b200_transport visc, hexed_b200_transport therm_cond);
int hexed_b200_max_dt_device(hexed_b200_ctx* ctx, int pde, double convective_safety, double diffusive_safety, int local_time,
                             hexed_b200_transport visc, hexed_b200_transport therm_cond, double advect_length, double* device_out);

/* ---- device group: ONE mesh on several GPUs driven by one host thread (what the C++ adapter uses when the Solver process owns more
 * than one device; the reference hands over a single Kernel_mesh, include/Kernel_mesh.hpp:14-25, so the split happens below the
 * drop-in boundary). The caller partitions the flattened tables (hexed_b200::partition in hexed_b200/host/partition.hpp), creates one
 * context per device, uploads each rank's mesh + hexed_b200_set_partition, registers each rank's halo and then calls the group
 * versions of the stage drivers. Transport: NCCL (bound at run time from libnccl.so.2; ncclCommInitAll over the contexts' devices),
 * ncclSend/ncclRecv of the cut faces on a communication stream per rank, overlapped with the interior Neighbor kernels, and
 * ncclAllReduce(min) of the time step. Results are those of the undivided mesh (same kernels, same operands; both owners of a cut
 * connection evaluate it). ---- */
typedef struct hexed_b200_group hexed_b200_group;
int hexed_b200_group_create(hexed_b200_group** group, int n, hexed_b200_ctx* const* contexts); /* contexts on n distinct devices */
int hexed_b200_group_destroy(hexed_b200_group* group);                                       /* the contexts stay alive */
const char* hexed_b200_group_last_error(const hexed_b200_group* group);                      /* group may be NULL: error of the last failed create */
int hexed_b200_group_size(const hexed_b200_group* group);
hexed_b200_ctx* hexed_b200_group_ctx(hexed_b200_group* group, int rank);
int hexed_b200_group_info(const hexed_b200_group* group, int* nccl_version, long long* exchanges, long long* bytes_sent);
/* rank's halo: for each of its n_peers peers, the n_send[i] local face slots it sends (concatenated in send_slots, ordered as the peer's
 * receive list) and the n_recv[i] halo slots it fills (recv_slots). Call after the rank's hexed_b200_mesh_create. */
int hexed_b200_group_set_halo(hexed_b200_group* group, int rank, int n_peers, const int* peers, const int* n_send, const int* send_slots,
                              const int* n_recv, const int* recv_slots);
int hexed_b200_group_exchange(hexed_b200_group* group, int kind); /* one face kind, e.g. after compute_write_face at initialisation */
int hexed_b200_group_synchronize(hexed_b200_group* group);
int hexed_b200_group_compute_euler(hexed_b200_group* group, hexed_b200_options opts);
int hexed_b200_group_compute_navier_stokes(hexed_b200_group* group, hexed_b200_options opts, hexed_b200_callback flux_bc, void* user,
                                           hexed_b200_transport visc, hexed_b200_transport therm_cond);
int hexed_b200_group_max_dt(hexed_b200_group* group, int pde, double convective_safety, double diffusive_safety, int local_time,
                            hexed_b200_transport visc, hexed_b200_transport therm_cond, double advect_length, double* dt);

/* ---- profiling side-contract ---- */
int hexed_b200_set_timing(hexed_b200_ctx* ctx, int enabled);
/* implementation switches (for A/B measurements and tests): HEXED_B200_OPT_PIPELINED_LOCAL = use the persistent TMA-pipelined
 * Local kernel where it applies (3-D row size 4 or 6, 2-D row size 4, 6 or 8, no modal filter); default 1. Value 2 = pipelined, but the
 * 3-D deformed kernel keeps its earlier shared-memory layout (everything staged, two resident CTAs per SM) instead of the lean one; 3 = additionally the lean layout with four
 * resident CTAs for 3-D Cartesian elements (measured slower on B200); 4 = the default kernels, but the stage buffer of an
 * element is handed back through an mbarrier (every thread arrives, only the refilling thread waits) instead of a CTA barrier (measured equal)
 * HEXED_B200_OPT_CFL_CACHE = the stage-1 Local kernel leaves min(spacing/char_speed) per element behind and the next global-time-step
 * max_dt_euler re-evaluates only the near-minimum elements instead of re-reading the whole state (any other write to the state
 * invalidates the screen); default 0 -- on B200 the extra work in the Local kernel costs what the saved pass over the state gains
 * HEXED_B200_OPT_FUSED_ADMIS = the pipelined Euler Local kernels also leave, per element, whether the state and faces they have just written
 * are admissible / finite; hexed_b200_is_admissible right after hexed_b200_compute_euler then reduces 4 bytes per element instead of scanning
 * the state (same answer and record; anything else that writes element state or faces in between falls back to the full scan). Default 0. */
enum { HEXED_B200_OPT_PIPELINED_LOCAL = 0, HEXED_B200_OPT_CFL_CACHE = 1, HEXED_B200_OPT_FUSED_ADMIS = 2,
       HEXED_B200_OPT_NS_LOCAL_LAYOUT = 3 /* 3-D row-size-6 Navier-Stokes Local, three variants with bit-identical results: 0 = ns_local_line_kernel (dense
                                             shared-memory fields, one bulk copy per array); 1 = ns_local_pad_kernel<38> (plane pitch 38: bank-conflict-free,
                                             one bulk copy per field plane); 2 = ns_local_pad_kernel<36> (dense fields with the task dealing, 16-byte accesses
                                             and early global loads of variant 1). The default is the one measured fastest on B200 */ };
int hexed_b200_set_option(hexed_b200_ctx* ctx, int option, int value);
int hexed_b200_kernel_stats(hexed_b200_ctx* ctx, hexed_b200_kernel_stat* out, int capacity, int* n_out);
int hexed_b200_reset_stats(hexed_b200_ctx* ctx);
long long hexed_b200_launch_count(const hexed_b200_ctx* ctx); /* kernels launched by this context so far */

#ifdef __cplusplus
}
#endif
#endif
