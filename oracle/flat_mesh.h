/* flat_mesh.h -- flattened description of a Hexed kernel mesh (oracle side).
 *
 * TEST INFRASTRUCTURE. This header describes the plain-array view of the
 * reference's `Kernel_mesh` (reference include/Kernel_mesh.hpp:14-25) that the
 * CPU oracle works on. The reference keeps one heap vector per element and per
 * connection and lets elements alias connection storage through raw pointers
 * (include/connection.hpp:111-123); here that pointer graph is flattened to
 * "slots" in a few big arrays so the same inputs can be handed to the oracle
 * and to the CUDA library.
 *
 * Conventions (3 = n_dim, 6 = row_size in the headline configuration):
 *   nq  = row_size^n_dim, nfq = row_size^(n_dim-1), nv = n_dim + 2
 *   elements [0, n_car) are Cartesian, [n_car, n_car+n_def) deformed
 *   elem_data[e][slot][nq] uses the reference's slot order (src/Element.cpp:114-142,187-189):
 *       state nv | tss 1 | bulk_av 1 | laplacian_av 1 | forcing 4 | advection rs | residual cache max(nv, rs)
 *   face slot of element e, face f (= 2*i_dim + sign) is e*2*n_dim + f; ghost, mortar and
 *       other connection-owned faces take slots after 2*n_dim*n_elem.
 *   face_state[slot][nv*nfq]  : extrapolated state / numerical flux   (`face(i, false)`)
 *   face_ldg  [slot][nv*nfq]  : LDG half of the face storage          (`face(i, true)`)
 *   face_wide [slot][(n_dim+row_size)*nfq] : face storage as seen by pde::Advection
 *       (in the reference this aliases the two arrays above; kept separate here)
 *   normal slot of deformed element d (= e - n_car), face f is d*2*n_dim + f;
 *       connection-owned normals that alias no element face come after.
 */
#ifndef HEXED_ORACLE_FLAT_MESH_H_
#define HEXED_ORACLE_FLAT_MESH_H_

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  int row_size;
  double node[8];
  double weight[8];
  double diff_mat[8][8];   /* M[i][j] = d(basis_j)/dx at node i */
  double boundary[2][8];
  double orthogonal[8][8]; /* [degree][node] */
  double filter[8][8];
  double prolong[2][8][8];
  double restrict_[2][8][8];
  double min_eig_convection;
  double min_eig_diffusion;
  double quadratic_safety;
} ho_basis;

typedef struct {
  double const_val, ref_val, ref_temp, sqrt_ref_temp, temp_offset;
  int is_viscous;
} ho_transport;

typedef struct {
  int n_dim, row_size;
  int n_car, n_def;
  int n_slot;            /* slots per element in elem_data */
  double* elem_data;     /* [n_elem][n_slot][nq] */
  double* nom_size;      /* [n_elem] */
  double* vertex_tss;    /* [n_elem][2^n_dim] */
  double* uncert;        /* [n_elem] */
  double* ref_normals;   /* [n_def][n_dim*n_dim][nq], layout [i_dim][j_dim][q] */
  double* det;           /* [n_def][nq] */
  int n_face_slot;
  double* face_state;    /* [n_face_slot][nv*nfq] */
  double* face_ldg;      /* [n_face_slot][nv*nfq] (may be null if unused) */
  double* face_wide;     /* [n_face_slot][(n_dim+row_size)*nfq] (may be null if unused) */
  int n_normal_slot;
  double* normals;       /* [n_normal_slot][n_dim][nfq] */
  int n_car_con;
  int* car_con;          /* [n_car_con][3]: slot0, slot1, i_dim */
  int n_def_con;
  int* def_con;          /* [n_def_con][7]: slot0, slot1, i_dim0, i_dim1, sign0, sign1, normal_slot */
  int n_ref;
  int* ref_face;         /* [n_ref][7]: coarse slot, fine slot 0..3 (-1 = none), stretch0, stretch1 */
} ho_mesh;

typedef struct {
  double dt;
  int i_stage;
  int compute_residual;
  int use_filter;
} ho_options;

/* which PDE a kernel is instantiated for */
enum { HO_EULER = 0, HO_NAVIER_STOKES = 1, HO_ADVECTION = 2, HO_SMOOTH_AV = 3, HO_FIX_THERM_ADMIS = 4 };

typedef void (*ho_callback)(void*);

/* stage drivers: same sequences as reference src/kernels_convective.cpp:8-19 and src/kernels_diffusive.cpp:8-32 */
int ho_compute_euler(const ho_basis*, ho_mesh*, ho_options);
int ho_compute_advection(const ho_basis*, ho_mesh*, ho_options, double advect_length);
int ho_compute_navier_stokes(const ho_basis*, ho_mesh*, ho_options, ho_callback flux_bc, void* user, ho_transport visc, ho_transport therm_cond);
int ho_compute_smooth_av(const ho_basis*, ho_mesh*, ho_options, ho_callback flux_bc, void* user, double diff_time, double cheby_step);
int ho_compute_fix_therm_admis(const ho_basis*, ho_mesh*, ho_options, ho_callback flux_bc, void* user);
/* reference src/kernels_max_dt.cpp:14-21 */
int ho_max_dt(int pde, const ho_basis*, ho_mesh*, double safety_conv, double safety_diff, int local_time,
              ho_transport visc, ho_transport therm_cond, double advect_length, double* dt_out);
/* reference src/kernels_convective.cpp:23-56 */
int ho_compute_write_face(int pde, const ho_basis*, ho_mesh*);
int ho_compute_prolong(int pde, const ho_basis*, ho_mesh*, int scale, int offset);
int ho_compute_restrict(int pde, const ho_basis*, ho_mesh*, int scale, int offset);
int ho_face_permutation(int n_dim, int row_size, int n_var, const int dir[4], int restore, double* data);
/* reference src/stabilizing_art_visc.cpp:8-66 */
int ho_stabilizing_art_visc(const ho_basis*, ho_mesh*, double char_speed);
/* individual kernels (for unit-level parity tests); which: 0 = cartesian set, 1 = deformed set */
int ho_neighbor(int pde, int deformed, ho_mesh*, int i_stage, ho_transport visc, ho_transport therm_cond, double p0, double p1);
int ho_local(int pde, int deformed, const ho_basis*, ho_mesh*, ho_options, ho_transport visc, ho_transport therm_cond, double p0, double p1);
int ho_neighbor_reconcile(int pde, int deformed, ho_mesh*);
int ho_reconcile_ldg_flux(int pde, int deformed, const ho_basis*, ho_mesh*, ho_options, ho_transport visc, ho_transport therm_cond, double p0, double p1);
/* 1-D operator exposed for the known-answer tests (reference include/Derivative.hpp:51-55) */
int ho_derivative(const ho_basis*, int n_var, const double* qpoint_vals, const double* boundary_vals, double* result);
/* ghost-state boundary conditions (reference src/Boundary_condition.cpp:66-76 Freestream, :450-453 Copy, :301-327 Nonpenetration) */
int ho_bc_freestream(ho_mesh*, int n_bc, const int* ghost_slot, const double* freestream);
int ho_bc_copy(ho_mesh*, int n_bc, const int* inside_slot, const int* ghost_slot);
int ho_bc_nonpenetration(ho_mesh*, int n_bc, const int* inside_slot, const int* ghost_slot, const int* normal_slot);
/* Riemann_invariants (src/Boundary_condition.cpp:97-182) and the Characteristics decomposition it uses (include/pde.hpp:181-256) */
int ho_characteristics(int n_dim, const double* state, const double* direction, const double* state1, double* eigvals, double* decomp);
int ho_bc_riemann_state(ho_mesh*, int n_bc, const int* inside_slot, const int* ghost_slot, const int* normal_slot, const double* freestream, double* cache);
int ho_bc_riemann_flux(ho_mesh*, int n_bc, const int* inside_slot, const int* ghost_slot, const int* normal_slot, const double* cache);
int ho_num_threads(void);

#ifdef __cplusplus
}
#endif
#endif
