"""ctypes binding of libhexed_b200.so plus a small host mirror of the reference's `kernels.hpp` entry points.

`Device` owns one `hexed_b200_ctx` (one GPU, one (n_dim, row_size)); `Device.load_mesh(FlatMesh)` creates the device
mirror of a flattened mesh and the methods `compute_euler`, `max_dt_euler`, `compute_write_face`, `compute_prolong`,
`compute_restrict`, `face_permutation` carry the same names, argument meaning and error behaviour as the reference's
free functions (include/kernels.hpp:22-42): invalid (n_dim, row_size) raises RuntimeError("demand for invalid kernel")
like include/kernel_factory.hpp:114-116.

There is no CPU path here. If the shared library is missing, or no CUDA device is present, construction raises.
"""
import ctypes as C
import os

import numpy as np

from .mesh import (BC_FREESTREAM, BC_COPY, BC_NONPENETRATION, BC_OUTFLOW, BC_PRESSURE_OUTFLOW, BC_NO_SLIP, BC_RIEMANN_INVARIANTS,
                   n_slot)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libhexed_b200.so")

(NOMINAL_SIZE, VERTEX_TSS, REF_NORMALS, JAC_DET, FACE_STATE, FACE_LDG, FACE_WIDE, NORMALS, UNCERT, VERTEX_SCRATCH) = range(10)

dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int)


class MeshDesc(C.Structure):
    _fields_ = [("n_car", C.c_int), ("n_def", C.c_int), ("n_face_slot", C.c_int), ("n_normal_slot", C.c_int),
                ("n_car_con", C.c_int), ("n_def_con", C.c_int), ("n_ref", C.c_int),
                ("car_con", ip), ("def_con", ip), ("ref_face", ip)]


class Options(C.Structure):
    _fields_ = [("dt", C.c_double), ("i_stage", C.c_int), ("compute_residual", C.c_int), ("use_filter", C.c_int)]


class Transport(C.Structure):
    _fields_ = [("const_val", C.c_double), ("ref_val", C.c_double), ("ref_temp", C.c_double),
                ("sqrt_ref_temp", C.c_double), ("temp_offset", C.c_double), ("is_viscous", C.c_int)]


class KernelStat(C.Structure):
    _fields_ = [("name", C.c_char_p), ("deformed", C.c_int), ("work_units", C.c_longlong), ("launches", C.c_longlong),
                ("device_seconds", C.c_double)]


# every symbol include/hexed_b200.h declares, with its argument types (restype is int unless noted)
SIGNATURES = {
    "hexed_b200_device_count": [ip],
    "hexed_b200_create": [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, dp, C.c_int],
    "hexed_b200_destroy": [C.c_void_p],
    "hexed_b200_last_error": [C.c_void_p],
    "hexed_b200_synchronize": [C.c_void_p],
    "hexed_b200_cuda_stream": [C.c_void_p, C.POINTER(C.c_void_p)],
    "hexed_b200_mesh_create": [C.c_void_p, C.POINTER(MeshDesc)],
    "hexed_b200_upload": [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_size_t],
    "hexed_b200_download": [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_size_t],
    "hexed_b200_upload_elem_slots": [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int],
    "hexed_b200_download_elem_slots": [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int],
    "hexed_b200_face_list_create": [C.c_void_p, ip, C.c_int, ip],
    "hexed_b200_face_list_download": [C.c_void_p, C.c_int, C.c_int, C.c_void_p],
    "hexed_b200_face_list_upload": [C.c_void_p, C.c_int, C.c_int, C.c_void_p],
    "hexed_b200_face_permutation_table": [C.c_void_p, ip, ip],
    "hexed_b200_face_permutation_indices": [C.c_int, C.c_int, ip, ip],
    "hexed_b200_set_partition": [C.c_void_p, C.c_int, C.c_int, C.c_int, ip],
    "hexed_b200_face_list_gather": [C.c_void_p, C.c_int, C.c_int, C.c_void_p],
    "hexed_b200_face_list_scatter": [C.c_void_p, C.c_int, C.c_int, C.c_void_p],
    "hexed_b200_face_list_prefetch": [C.c_void_p, C.c_int, C.c_int],
    "hexed_b200_face_list_prefetched": [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)],
    "hexed_b200_face_list_staging": [C.c_void_p, C.c_int, C.POINTER(C.c_void_p)],
    "hexed_b200_face_list_upload_deferred": [C.c_void_p, C.c_int, C.c_int],
    "hexed_b200_compute_euler_begin": [C.c_void_p],
    "hexed_b200_compute_euler_finish": [C.c_void_p, Options],
    "hexed_b200_compute_navier_stokes_begin": [C.c_void_p, Options, Transport, Transport],
    "hexed_b200_compute_navier_stokes_middle": [C.c_void_p, Options, C.c_void_p, C.c_void_p, Transport, Transport],
    "hexed_b200_compute_navier_stokes_middle_local": [C.c_void_p, Options, Transport, Transport],
    "hexed_b200_compute_navier_stokes_middle_reconcile": [C.c_void_p, Options, Transport, Transport],
    "hexed_b200_max_dt_device": [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_int, Transport, Transport, C.c_double, C.c_void_p],
    "hexed_b200_group_create": [C.POINTER(C.c_void_p), C.c_int, C.POINTER(C.c_void_p)],
    "hexed_b200_group_destroy": [C.c_void_p],
    "hexed_b200_group_last_error": [C.c_void_p],
    "hexed_b200_group_size": [C.c_void_p],
    "hexed_b200_group_ctx": [C.c_void_p, C.c_int],
    "hexed_b200_group_info": [C.c_void_p, ip, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)],
    "hexed_b200_group_set_halo": [C.c_void_p, C.c_int, C.c_int, ip, ip, ip, ip, ip],
    "hexed_b200_group_exchange": [C.c_void_p, C.c_int],
    "hexed_b200_group_synchronize": [C.c_void_p],
    "hexed_b200_group_compute_euler": [C.c_void_p, Options],
    "hexed_b200_group_compute_navier_stokes": [C.c_void_p, Options, C.c_void_p, C.c_void_p, Transport, Transport],
    "hexed_b200_group_max_dt": [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_int, Transport, Transport, C.c_double, dp],
    "hexed_b200_compute_navier_stokes_finish": [C.c_void_p, Options, Transport, Transport],
    "hexed_b200_compute_euler": [C.c_void_p, Options],
    "hexed_b200_max_dt_euler": [C.c_void_p, Options, C.c_double, C.c_double, C.c_int, dp],
    "hexed_b200_compute_write_face": [C.c_void_p],
    "hexed_b200_compute_prolong": [C.c_void_p, C.c_int, C.c_int],
    "hexed_b200_compute_restrict": [C.c_void_p, C.c_int, C.c_int],
    "hexed_b200_face_permutation": [C.c_void_p, ip, C.c_int, dp],
    "hexed_b200_compute_advection": [C.c_void_p, Options, C.c_double],
    "hexed_b200_compute_navier_stokes": [C.c_void_p, Options, C.c_void_p, C.c_void_p, Transport, Transport],
    "hexed_b200_compute_smooth_av": [C.c_void_p, Options, C.c_void_p, C.c_void_p, C.c_double, C.c_double],
    "hexed_b200_compute_fix_therm_admis": [C.c_void_p, Options, C.c_void_p, C.c_void_p],
    "hexed_b200_max_dt_navier_stokes": [C.c_void_p, Options, C.c_double, C.c_double, C.c_int, Transport, Transport, dp],
    "hexed_b200_max_dt_advection": [C.c_void_p, Options, C.c_double, C.c_double, C.c_int, C.c_double, dp],
    "hexed_b200_max_dt_smooth_av": [C.c_void_p, Options, C.c_double, C.c_double, C.c_int, dp],
    "hexed_b200_max_dt_fix_therm_admis": [C.c_void_p, Options, C.c_double, C.c_double, C.c_int, dp],
    "hexed_b200_compute_prolong_advection": [C.c_void_p],
    "hexed_b200_compute_write_face_advection": [C.c_void_p],
    "hexed_b200_compute_write_face_smooth_av": [C.c_void_p],
    "hexed_b200_stabilizing_art_visc": [C.c_void_p, C.c_double],
    "hexed_b200_pde_kernel": [C.c_void_p, C.c_int, C.c_int, C.c_int, Options, Transport, Transport, C.c_double, C.c_double],
    "hexed_b200_apply_flux_bcs": [C.c_void_p],
    "hexed_b200_set_jacobian": [C.c_void_p, dp, dp],
    "hexed_b200_calc_shared_normals": [C.c_void_p],
    "hexed_b200_av_scale_velocity": [C.c_void_p, C.c_int],
    "hexed_b200_av_project_forcing": [C.c_void_p, dp, dp],
    "hexed_b200_av_finish": [C.c_void_p, C.c_double, C.c_double, C.c_int, dp, dp],
    "hexed_b200_interp_vertices": [C.c_void_p, C.c_int, dp, dp],
    "hexed_b200_av_swap": [C.c_void_p],
    "hexed_b200_av_elwise_ramp": [C.c_void_p, C.c_double],
    "hexed_b200_av_elwise_forcing": [C.c_void_p, C.c_int],
    "hexed_b200_av_elwise_vertices": [C.c_void_p, dp],
    "hexed_b200_apply_aux_bcs": [C.c_void_p, C.c_int],
    "hexed_b200_update_euler": [C.c_void_p, C.c_double, C.c_int, C.c_int, C.c_int, dp, dp],
    "hexed_b200_update_navier_stokes": [C.c_void_p, C.c_double, Transport, Transport, C.c_int, C.c_int, C.c_int, dp, dp],
    "hexed_b200_is_admissible": [C.c_void_p, ip],
    "hexed_b200_is_admissible_begin": [C.c_void_p],
    "hexed_b200_is_admissible_finish": [C.c_void_p, C.POINTER(C.c_int)],
    "hexed_b200_vertex_topology": [C.c_void_p, ip, C.c_int, ip, C.c_int],
    "hexed_b200_share_vertex_data": [C.c_void_p, C.c_int, C.c_int],
    "hexed_b200_fix_admis_spread": [C.c_void_p, dp],
    "hexed_b200_download_record": [C.c_void_p, ip, C.c_int, C.c_int],
    "hexed_b200_neighbor_euler": [C.c_void_p, C.c_int],
    "hexed_b200_local_euler": [C.c_void_p, C.c_int, Options],
    "hexed_b200_bc_create": [C.c_void_p, C.c_int, C.c_int, ip, ip, ip, dp, C.c_int, ip],
    "hexed_b200_apply_state_bcs": [C.c_void_p],
    "hexed_b200_bc_set_params": [C.c_void_p, C.c_int, dp, C.c_int],
    "hexed_b200_set_timing": [C.c_void_p, C.c_int],
    "hexed_b200_set_option": [C.c_void_p, C.c_int, C.c_int],
    "hexed_b200_kernel_stats": [C.c_void_p, C.POINTER(KernelStat), C.c_int, ip],
    "hexed_b200_reset_stats": [C.c_void_p],
    "hexed_b200_launch_count": [C.c_void_p],
}
_RESTYPES = {"hexed_b200_last_error": C.c_char_p, "hexed_b200_launch_count": C.c_longlong,
             "hexed_b200_group_last_error": C.c_char_p, "hexed_b200_group_ctx": C.c_void_p}


CALLBACK = C.CFUNCTYPE(None, C.c_void_p)
NAVIER_STOKES, ADVECTION, SMOOTH_AV, FIX_THERM_ADMIS = 1, 2, 3, 4
K_NEIGHBOR, K_LOCAL, K_NEIGHBOR_RECONCILE, K_RECONCILE_LDG = 0, 1, 2, 3


def inviscid():
    """Transport_model::inviscid() (reference include/Transport_model.hpp:40)"""
    return Transport(0., 0., 1., 1., 1., 0)


def constant_transport(value):
    """Transport_model::constant (reference include/Transport_model.hpp:42)"""
    return Transport(value, 0., 1., 1., 1., 1)


def sutherland(ref_val, ref_temp, temp_offset):
    """Transport_model::sutherland (reference include/Transport_model.hpp:44-47)"""
    return Transport(0., ref_val, ref_temp, float(np.sqrt(ref_temp)), temp_offset, 1)


def load_library(path=None):
    """load the C-ABI library and attach prototypes; raises (loudly) if it has not been built"""
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise RuntimeError("%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(hexed_b200 has no CPU fallback)" % path)
    lib = C.CDLL(path)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, C.c_int)
    return lib


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _addr(a):
    """address of a numpy array or torch tensor (host or device)"""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        assert a.flags["C_CONTIGUOUS"] and a.dtype == np.float64
        return a.ctypes.data
    assert a.is_contiguous() and str(a.dtype) == "torch.float64"
    return a.data_ptr()


BC_MODE_ADVECTION, BC_MODE_COPY_STATE, BC_MODE_NEGATE_FLUX = 0, 1, 2  # include/hexed_b200.h
OPT_PIPELINED_LOCAL, OPT_CFL_CACHE, OPT_FUSED_ADMIS = 0, 1, 2      # hexed_b200_set_option


class Device:
    def __init__(self, n_dim, row_size, basis, device=0, lib_path=None):
        self.lib = load_library(lib_path)
        self.n_dim, self.row_size, self.basis = n_dim, row_size, basis
        self.ctx = C.c_void_p()
        packed = np.ascontiguousarray(basis.packed()) if basis is not None else np.zeros(1)
        rc = self.lib.hexed_b200_create(C.byref(self.ctx), device, n_dim, row_size, packed.ctypes.data_as(dp), packed.size)
        if rc:
            msg = self.lib.hexed_b200_last_error(None).decode()
            self.ctx = None
            raise RuntimeError(msg)
        self.mesh = None

    def close(self):
        if getattr(self, "ctx", None):
            self.lib.hexed_b200_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc:
            raise RuntimeError(self.lib.hexed_b200_last_error(self.ctx).decode())

    # ---- mesh epoch ----
    def load_mesh(self, m, upload_elem_data=True):
        """create the device mirror of FlatMesh `m` and upload everything the kernels read"""
        assert m.n_dim == self.n_dim and m.row_size == self.row_size
        car, dfc, ref = _i32(m.car_con), _i32(m.def_con), _i32(m.ref_face)
        d = MeshDesc(m.n_car, m.n_def, m.n_face_slot, m.n_normal_slot, car.shape[0], dfc.shape[0], ref.shape[0],
                     car.ctypes.data_as(ip), dfc.ctypes.data_as(ip), ref.ctypes.data_as(ip))
        self._check(self.lib.hexed_b200_mesh_create(self.ctx, C.byref(d)))
        self.mesh = m
        self.upload(NOMINAL_SIZE, m.nom_size)
        self.upload(VERTEX_TSS, m.vertex_tss)
        if m.n_def:
            self.upload(REF_NORMALS, m.ref_normals)
            self.upload(JAC_DET, m.det)
        if m.n_normal_slot:
            self.upload(NORMALS, m.normals)
        if upload_elem_data and m.elem_data is not None:
            self.upload_elements(m.elem_data)
        if m.face_state is not None:
            self.upload(FACE_STATE, m.face_state)
        if m.face_ldg is not None:
            self.upload(FACE_LDG, m.face_ldg)
        if m.face_wide is not None:
            self.upload(FACE_WIDE, m.face_wide)
        self.bc_ids = []
        for bc in m.bcs:
            self.bc_ids.append(self.add_bc(bc))
        if getattr(m, "halo", None) is not None:
            pp = _i32(m.pre_prolong)
            self._check(self.lib.hexed_b200_set_partition(self.ctx, m.n_cut_car, m.n_cut_def, pp.size, pp.ctypes.data_as(ip)))
            self.send_lists = {peer: self.face_list(s) for peer, s in m.halo.send.items()}
            self.recv_lists = {peer: self.face_list(s) for peer, s in m.halo.recv.items()}
        return self

    def add_bc(self, bc):
        kind = bc["kind"]
        if kind not in (BC_FREESTREAM, BC_COPY, BC_NONPENETRATION, BC_OUTFLOW, BC_PRESSURE_OUTFLOW, BC_NO_SLIP, BC_RIEMANN_INVARIANTS):
            return None  # host-applied boundary condition
        ins, gh, nr = _i32(bc["inside_slot"]), _i32(bc["ghost_slot"]), _i32(bc["normal_slot"])
        params = np.ascontiguousarray(bc["params"], dtype=np.float64) if bc.get("params") is not None else np.zeros(0)
        out = C.c_int(-1)
        self._check(self.lib.hexed_b200_bc_create(self.ctx, kind, ins.size, ins.ctypes.data_as(ip), gh.ctypes.data_as(ip),
                                                  nr.ctypes.data_as(ip), params.ctypes.data_as(dp), params.size, C.byref(out)))
        return out.value

    def upload(self, which, arr, first=0):
        n = arr.shape[0]
        self._check(self.lib.hexed_b200_upload(self.ctx, which, _addr(arr), first, n))

    def download(self, which, arr, first=0):
        self._check(self.lib.hexed_b200_download(self.ctx, which, _addr(arr), first, arr.shape[0]))
        return arr

    def upload_elements(self, elem_data, first_slot=0, n_slots=None, first_elem=0):
        """elem_data: (n, n_slots_in_array, nq) in the reference's slot order, starting at `first_slot`"""
        n, ns, nq = elem_data.shape
        n_slots = ns if n_slots is None else n_slots
        self._check(self.lib.hexed_b200_upload_elem_slots(self.ctx, _addr(elem_data), ns*nq, first_slot, n_slots, first_elem, n))

    def download_elements(self, elem_data, first_slot=0, n_slots=None, first_elem=0):
        n, ns, nq = elem_data.shape
        n_slots = ns if n_slots is None else n_slots
        self._check(self.lib.hexed_b200_download_elem_slots(self.ctx, _addr(elem_data), ns*nq, first_slot, n_slots, first_elem, n))
        return elem_data

    def sync_to_host(self, m=None):
        """conservative sync-out: bring back everything the kernels may have written"""
        m = m or self.mesh
        self.download_elements(m.elem_data)
        self.download(FACE_STATE, m.face_state)
        if m.face_ldg is not None:
            self.download(FACE_LDG, m.face_ldg)
        if m.face_wide is not None:
            self.download(FACE_WIDE, m.face_wide)
        self.download(UNCERT, m.uncert)
        return m

    def face_list(self, slots):
        s = _i32(slots)
        out = C.c_int(-1)
        self._check(self.lib.hexed_b200_face_list_create(self.ctx, s.ctypes.data_as(ip), s.size, C.byref(out)))
        return out.value

    def face_list_download(self, list_id, dst, kind=0):
        self._check(self.lib.hexed_b200_face_list_download(self.ctx, list_id, kind, _addr(dst)))
        return dst

    def face_list_upload(self, list_id, src, kind=0):
        self._check(self.lib.hexed_b200_face_list_upload(self.ctx, list_id, kind, _addr(src)))

    def face_list_gather(self, list_id, device_dst, kind=0):
        """asynchronous on the context's stream; `device_dst` must be device memory"""
        self._check(self.lib.hexed_b200_face_list_gather(self.ctx, list_id, kind, _addr(device_dst)))

    def face_list_scatter(self, list_id, device_src, kind=0):
        self._check(self.lib.hexed_b200_face_list_scatter(self.ctx, list_id, kind, _addr(device_src)))

    def set_partition(self, n_cut_car, n_cut_def, pre_prolong=()):
        """the last n_cut_* rows of the connection tables wait for faces that arrive from outside (halo exchange, or ghost faces a
        host boundary condition writes): compute_euler_begin skips them, compute_euler_finish does them first"""
        pp = _i32(pre_prolong)
        self._check(self.lib.hexed_b200_set_partition(self.ctx, int(n_cut_car), int(n_cut_def), pp.size, pp.ctypes.data_as(ip)))

    def compute_euler_begin(self):
        self._check(self.lib.hexed_b200_compute_euler_begin(self.ctx))

    def compute_euler_finish(self, **kw):
        self._check(self.lib.hexed_b200_compute_euler_finish(self.ctx, self._opts(**kw)))

    def compute_navier_stokes_begin(self, visc, therm_cond, **kw):
        self._check(self.lib.hexed_b200_compute_navier_stokes_begin(self.ctx, self._opts(**kw), visc, therm_cond))

    def compute_navier_stokes_middle(self, flux_bc, visc, therm_cond, **kw):
        self._check(self.lib.hexed_b200_compute_navier_stokes_middle(self.ctx, self._opts(**kw), self._cb(flux_bc), None, visc, therm_cond))

    def compute_navier_stokes_finish(self, visc, therm_cond, **kw):
        self._check(self.lib.hexed_b200_compute_navier_stokes_finish(self.ctx, self._opts(**kw), visc, therm_cond))

    def synchronize(self):
        self._check(self.lib.hexed_b200_synchronize(self.ctx))

    def cuda_stream(self):
        s = C.c_void_p()
        self._check(self.lib.hexed_b200_cuda_stream(self.ctx, C.byref(s)))
        return s.value or 0

    # ---- mirrors of kernels.hpp ----
    @staticmethod
    def _opts(dt=1., i_stage=0, compute_residual=False, use_filter=False):
        return Options(dt, int(i_stage), int(compute_residual), int(use_filter))

    def compute_euler(self, **kw):
        self._check(self.lib.hexed_b200_compute_euler(self.ctx, self._opts(**kw)))

    def max_dt_euler(self, convective_safety, diffusive_safety, local_time, **kw):
        out = C.c_double(0.)
        self._check(self.lib.hexed_b200_max_dt_euler(self.ctx, self._opts(**kw), convective_safety, diffusive_safety, int(local_time), C.byref(out)))
        return out.value

    def compute_write_face(self):
        self._check(self.lib.hexed_b200_compute_write_face(self.ctx))

    def compute_prolong(self, scale=False, offset=False):
        self._check(self.lib.hexed_b200_compute_prolong(self.ctx, int(scale), int(offset)))

    def compute_restrict(self, scale=True, offset=False):
        self._check(self.lib.hexed_b200_compute_restrict(self.ctx, int(scale), int(offset)))

    def face_permutation(self, direction, data, restore=False):
        d = (C.c_int*4)(*direction.as_list())
        assert data.size == (self.n_dim + 2)*self.row_size**(self.n_dim - 1)
        self._check(self.lib.hexed_b200_face_permutation(self.ctx, d, int(restore), data.ctypes.data_as(dp)))
        return data

    def face_permutation_table(self, direction):
        d = (C.c_int*4)(*direction.as_list())
        out = np.zeros(self.row_size**(self.n_dim - 1), np.int32)
        self._check(self.lib.hexed_b200_face_permutation_table(self.ctx, d, out.ctypes.data_as(ip)))
        return out

    def _cb(self, flux_bc):
        if flux_bc is None:
            return None
        cb = CALLBACK(lambda _: flux_bc())
        self._keep = cb
        return C.cast(cb, C.c_void_p)

    def compute_advection(self, advect_length, **kw):
        self._check(self.lib.hexed_b200_compute_advection(self.ctx, self._opts(**kw), advect_length))

    def compute_navier_stokes(self, flux_bc, visc, therm_cond, **kw):
        self._check(self.lib.hexed_b200_compute_navier_stokes(self.ctx, self._opts(**kw), self._cb(flux_bc), None, visc, therm_cond))

    def compute_smooth_av(self, flux_bc, diff_time, chebyshev_step, **kw):
        self._check(self.lib.hexed_b200_compute_smooth_av(self.ctx, self._opts(**kw), self._cb(flux_bc), None, diff_time, chebyshev_step))

    def compute_fix_therm_admis(self, flux_bc, **kw):
        self._check(self.lib.hexed_b200_compute_fix_therm_admis(self.ctx, self._opts(**kw), self._cb(flux_bc), None))

    def max_dt_navier_stokes(self, convective_safety, diffusive_safety, local_time, visc, therm_cond, **kw):
        out = C.c_double(0.)
        self._check(self.lib.hexed_b200_max_dt_navier_stokes(self.ctx, self._opts(**kw), convective_safety, diffusive_safety, int(local_time), visc, therm_cond, C.byref(out)))
        return out.value

    def max_dt_advection(self, convective_safety, diffusive_safety, local_time, advect_length, **kw):
        out = C.c_double(0.)
        self._check(self.lib.hexed_b200_max_dt_advection(self.ctx, self._opts(**kw), convective_safety, diffusive_safety, int(local_time), advect_length, C.byref(out)))
        return out.value

    def max_dt_smooth_av(self, convective_safety, diffusive_safety, local_time, **kw):
        out = C.c_double(0.)
        self._check(self.lib.hexed_b200_max_dt_smooth_av(self.ctx, self._opts(**kw), convective_safety, diffusive_safety, int(local_time), C.byref(out)))
        return out.value

    def max_dt_fix_therm_admis(self, convective_safety, diffusive_safety, local_time, **kw):
        out = C.c_double(0.)
        self._check(self.lib.hexed_b200_max_dt_fix_therm_admis(self.ctx, self._opts(**kw), convective_safety, diffusive_safety, int(local_time), C.byref(out)))
        return out.value

    def compute_prolong_advection(self):
        self._check(self.lib.hexed_b200_compute_prolong_advection(self.ctx))

    def compute_write_face_advection(self):
        self._check(self.lib.hexed_b200_compute_write_face_advection(self.ctx))

    def compute_write_face_smooth_av(self):
        self._check(self.lib.hexed_b200_compute_write_face_smooth_av(self.ctx))

    def stabilizing_art_visc(self, char_speed):
        self._check(self.lib.hexed_b200_stabilizing_art_visc(self.ctx, char_speed))

    def pde_kernel(self, pde, which, deformed, visc=None, therm_cond=None, p0=1., p1=1., **kw):
        self._check(self.lib.hexed_b200_pde_kernel(self.ctx, pde, which, int(deformed), self._opts(**kw), visc or inviscid(),
                                                   therm_cond or inviscid(), p0, p1))

    def apply_flux_bcs(self):
        self._check(self.lib.hexed_b200_apply_flux_bcs(self.ctx))

    # ---- pointwise loops of the artificial-viscosity pipelines (reference src/Solver.cpp:457-581, 1021-1038) ----
    def av_scale_velocity(self, restore=False):
        self._check(self.lib.hexed_b200_av_scale_velocity(self.ctx, int(restore)))

    def av_project_forcing(self, node_weights, orthogonal):
        w = np.ascontiguousarray(node_weights, dtype=np.float64); o = np.ascontiguousarray(orthogonal, dtype=np.float64)
        self._check(self.lib.hexed_b200_av_project_forcing(self.ctx, w.ctypes.data_as(dp), o.ctypes.data_as(dp)))

    def av_finish(self, mult, us_max, n_real, node_weights):
        w = np.ascontiguousarray(node_weights, dtype=np.float64)
        out = C.c_double(0.)
        self._check(self.lib.hexed_b200_av_finish(self.ctx, float(mult), float(us_max), int(n_real), w.ctypes.data_as(dp),
                                                  C.cast(C.byref(out), dp)))
        return out.value

    def interp_vertices(self, target, vertex_values, interp):
        v = np.ascontiguousarray(vertex_values, dtype=np.float64); i = np.ascontiguousarray(interp, dtype=np.float64)
        self._check(self.lib.hexed_b200_interp_vertices(self.ctx, int(target), v.ctypes.data_as(dp), i.ctypes.data_as(dp)))

    def apply_aux_bcs(self, mode):
        """BC_MODE_ADVECTION / BC_MODE_COPY_STATE / BC_MODE_NEGATE_FLUX: boundary loops of the AV and admissibility pipelines"""
        self._check(self.lib.hexed_b200_apply_aux_bcs(self.ctx, int(mode)))

    def av_swap(self):
        self._check(self.lib.hexed_b200_av_swap(self.ctx))

    def av_elwise_ramp(self, scale):
        """Solver::update_art_visc_elwise, the ramp of src/Solver.cpp:590-601 on the element uncertainties"""
        self._check(self.lib.hexed_b200_av_elwise_ramp(self.ctx, scale))

    def av_elwise_forcing(self, restore):
        self._check(self.lib.hexed_b200_av_elwise_forcing(self.ctx, int(restore)))

    def av_elwise_vertices(self, interp):
        a = np.ascontiguousarray(interp, dtype=np.float64)
        self._check(self.lib.hexed_b200_av_elwise_vertices(self.ctx, a.ctypes.data_as(dp)))

    def set_jacobian(self, vertex_pos, node_adj=None):
        """element loop of Solver::calc_jacobian (reference src/Solver.cpp:281-286, src/Deformed_element.cpp:60-136): vertex_pos
        (n_def, 2^nd, nd), node_adj (n_def, 2*nd, nfq) or None"""
        v = np.ascontiguousarray(vertex_pos, dtype=np.float64)
        a = None if node_adj is None else np.ascontiguousarray(node_adj, dtype=np.float64)
        self._check(self.lib.hexed_b200_set_jacobian(self.ctx, v.ctypes.data_as(dp), None if a is None else a.ctypes.data_as(dp)))

    def vertex_topology(self, elem_vertex, n_vertex, matchers=None):
        """per-epoch vertex connectivity for share_vertex_data: elem_vertex (n_elem, 2^nd) vertex ids, matchers (n_match, 8) rows
        {i_dim, is_positive, stretch0, stretch1, fine elements (4, -1 padded)} (reference src/Hanging_vertex_matcher.cpp)"""
        ev = _i32(elem_vertex)
        mt = _i32(matchers if matchers is not None else np.zeros((0, 8)))
        self._check(self.lib.hexed_b200_vertex_topology(self.ctx, ev.ctypes.data_as(ip), int(n_vertex), mt.ctypes.data_as(ip), mt.shape[0] if mt.ndim == 2 else 0))

    def share_vertex_data(self, which, op_max):
        """Solver::share_vertex_data (reference src/Solver.cpp:35-54) on VERTEX_TSS or VERTEX_SCRATCH; op_max False = min"""
        self._check(self.lib.hexed_b200_share_vertex_data(self.ctx, int(which), int(bool(op_max))))

    def fix_admis_spread(self, interp):
        i = np.ascontiguousarray(interp, dtype=np.float64)
        self._check(self.lib.hexed_b200_fix_admis_spread(self.ctx, i.ctypes.data_as(dp)))

    def calc_shared_normals(self):
        """connection passes of Solver::calc_jacobian (reference src/Solver.cpp:287-369)"""
        self._check(self.lib.hexed_b200_calc_shared_normals(self.ctx))

    def update_euler(self, safety, n_steps, use_graph=True, n_cheby=1):
        """n_steps of Solver::update's inviscid flow loop (max_dt + 2 stages with device ghost fills, Chebyshev factors cycling through
        n_cheby) without a host round trip per step; returns (last dt, flow time advanced)"""
        dt, t = C.c_double(0.), C.c_double(0.)
        self._check(self.lib.hexed_b200_update_euler(self.ctx, float(safety), int(n_cheby), int(n_steps), int(bool(use_graph)),
                                                     C.cast(C.byref(dt), dp), C.cast(C.byref(t), dp)))
        return dt.value, t.value

    def update_navier_stokes(self, safety, visc, therm_cond, n_steps, use_graph=True, n_cheby=1):
        """the viscous counterpart of update_euler (stage 0 compute_navier_stokes with device flux boundary conditions, stage 1 compute_euler)"""
        dt, t = C.c_double(0.), C.c_double(0.)
        self._check(self.lib.hexed_b200_update_navier_stokes(self.ctx, float(safety), visc, therm_cond, int(n_cheby), int(n_steps),
                                                             int(bool(use_graph)), C.cast(C.byref(dt), dp), C.cast(C.byref(t), dp)))
        return dt.value, t.value

    def is_admissible(self):
        """Solver::is_admissible (reference src/Solver.cpp:921-958); raises RuntimeError("state is not finite") like the reference's
        HEXED_ASSERT (src/thermo.cpp:14)"""
        out = C.c_int(-1)
        self._check(self.lib.hexed_b200_is_admissible(self.ctx, C.byref(out)))
        return bool(out.value)

    def record(self):
        """Element::record of every element as left by the last is_admissible (1 = inadmissible)"""
        n = self.mesh.n_elem
        out = np.zeros(n, np.int32)
        self._check(self.lib.hexed_b200_download_record(self.ctx, out.ctypes.data_as(ip), 0, n))
        return out

    def neighbor_euler(self, deformed):
        self._check(self.lib.hexed_b200_neighbor_euler(self.ctx, int(deformed)))

    def local_euler(self, deformed, **kw):
        self._check(self.lib.hexed_b200_local_euler(self.ctx, int(deformed), self._opts(**kw)))

    def apply_state_bcs(self):
        self._check(self.lib.hexed_b200_apply_state_bcs(self.ctx))

    def bc_set_params(self, bc_id, params):
        """new parameter block for a registered boundary condition (e.g. a time-dependent freestream state)"""
        p = np.ascontiguousarray(params, dtype=np.float64)
        self._check(self.lib.hexed_b200_bc_set_params(self.ctx, int(bc_id), p.ctypes.data_as(dp), p.size))

    # ---- profiling side-contract ----
    def set_timing(self, enabled):
        self._check(self.lib.hexed_b200_set_timing(self.ctx, int(enabled)))

    def set_option(self, option, value):
        self._check(self.lib.hexed_b200_set_option(self.ctx, option, int(value)))

    def kernel_stats(self):
        buf = (KernelStat*16)()
        n = C.c_int(0)
        self._check(self.lib.hexed_b200_kernel_stats(self.ctx, buf, 16, C.byref(n)))
        return [dict(name=buf[i].name.decode(), deformed=buf[i].deformed, work_units=buf[i].work_units,
                     launches=buf[i].launches, device_seconds=buf[i].device_seconds) for i in range(n.value)]

    def reset_stats(self):
        self._check(self.lib.hexed_b200_reset_stats(self.ctx))

    def launch_count(self):
        return self.lib.hexed_b200_launch_count(self.ctx)
