/* oracle_call.h -- internal call record between oracle_api.cpp and the per-(n_dim, PDE) parts.
 * TEST INFRASTRUCTURE (see oracle_impl.hpp). The oracle is compiled as 15 translation units
 * (3 dimensionalities x 5 PDEs) so that `make -j` finishes in about a minute. */
#ifndef HEXED_ORACLE_CALL_H_
#define HEXED_ORACLE_CALL_H_
#include "flat_mesh.h"

enum { HO_OP_CONV_STAGE, HO_OP_DIFF_STAGE, HO_OP_MAX_DT, HO_OP_WRITE_FACE, HO_OP_NEIGHBOR, HO_OP_LOCAL,
       HO_OP_NEIGHBOR_RECONCILE, HO_OP_RECONCILE_LDG };

struct ho_call
{
  int op;
  const ho_basis* basis;
  ho_mesh* mesh;
  ho_options opts;
  ho_callback flux_bc;
  void* user;
  ho_transport visc, cond;
  double p0, p1;      /* PDE scalars: advect_length | diff_time, cheby_step */
  int deformed;
  int local_time;
  double safety_conv, safety_diff;
  double* dt_out;
};

typedef int (*ho_part_fn)(ho_call*);
#define HO_DECLARE_PART(ND, PDE) extern "C" int ho_part_##ND##_##PDE(ho_call*);
HO_DECLARE_PART(1, 0) HO_DECLARE_PART(1, 1) HO_DECLARE_PART(1, 2) HO_DECLARE_PART(1, 3) HO_DECLARE_PART(1, 4)
HO_DECLARE_PART(2, 0) HO_DECLARE_PART(2, 1) HO_DECLARE_PART(2, 2) HO_DECLARE_PART(2, 3) HO_DECLARE_PART(2, 4)
HO_DECLARE_PART(3, 0) HO_DECLARE_PART(3, 1) HO_DECLARE_PART(3, 2) HO_DECLARE_PART(3, 3) HO_DECLARE_PART(3, 4)
#endif
