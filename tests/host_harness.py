"""ctypes driver of hexed_b200/host/harness.cpp: builds a reference-shaped pointer-graph mesh from a FlatMesh and calls the
C++ adapter's `hexed::` entry points (hexed_b200/host/adapter.cpp) on it. Test infrastructure only."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST_DIR = os.path.join(ROOT, "hexed_b200", "host")
GPU_LIB = os.path.join(ROOT, "hexed_b200", "libhexed_b200_host.so")
EMU_LIB = os.path.join(ROOT, "tests", "emu", "libhexed_b200_host_emu.so")

dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
FN = dict(compute_euler=0, compute_advection=1, compute_navier_stokes=2, compute_smooth_av=3, compute_fix_therm_admis=4,
          max_dt_euler=5, max_dt_navier_stokes=6, max_dt_advection=7, max_dt_smooth_av=8, max_dt_fix_therm_admis=9,
          compute_prolong=10, compute_restrict=11, compute_prolong_advection=12, compute_write_face=13,
          compute_write_face_advection=14, compute_write_face_smooth_av=15, stabilizing_art_visc=16)
# hexed_b200::Data_group (hexed_b200/host/adapter.hpp)
STATE, TSS, ART_VISC, ADVECTION, RES_CACHE, FACES, FACES_WIDE, GEOMETRY, UNCERT = 1, 2, 4, 8, 16, 32, 64, 128, 256
ALL_ELEM = STATE | TSS | ART_VISC | ADVECTION | RES_CACHE
SYNC_EVERY_CALL, RESIDENT = 0, 1
FLUX_CB = C.CFUNCTYPE(None)


def build(emu):
    if emu:
        import fcntl
        with open(os.path.join(ROOT, "tests", "emu", ".build.lock"), "w") as lock:  # (pytest-xdist workers share the build tree)
            fcntl.flock(lock, fcntl.LOCK_EX)
            subprocess.run(["make", "-C", os.path.join(ROOT, "hexed_b200", "csrc"), "-j8", "emu"], check=True, stdout=subprocess.DEVNULL)
            subprocess.run(["make", "-C", HOST_DIR, "emu"], check=True, stdout=subprocess.DEVNULL)
        return EMU_LIB
    if not os.path.exists(GPU_LIB):
        raise RuntimeError("%s not built: run __graft_entry__.build()" % GPU_LIB)
    return GPU_LIB


def _d(a):
    return None if a is None else a.ctypes.data_as(dp)


def _i(a):
    return a.ctypes.data_as(ip)


def inviscid():
    return [0., 0., 0., 0.]


def constant(v):
    return [1., v, 0., 0.]


def sutherland(ref_val, ref_temp, offset):
    return [2., ref_val, ref_temp, offset]


def normal_present(m):
    """which normal slots hold a real normal: referenced by a deformed connection, or different from the unit fallback"""
    nd, nfq = m.n_dim, m.nfq
    present = np.zeros(m.n_normal_slot, np.int32)
    present[m.def_con[:, 6]] = 1
    for s in range(2*nd*m.n_def):
        unit = np.zeros((nd, nfq)); unit[(s % (2*nd))//2] = 1.
        if not np.array_equal(m.normals[s], unit):
            present[s] = 1
    present[2*nd*m.n_def:] = 1
    return present


class HostHarness:
    def __init__(self, lib_path, mesh, basis, seed=1):
        self.lib = C.CDLL(lib_path)
        L = self.lib
        L.hbh_create.restype = C.c_void_p
        L.hbh_create.argtypes = [C.c_int, C.c_int, dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, ip, C.c_int, ip, C.c_int, ip, C.c_int, ip, C.c_uint]
        L.hbh_destroy.argtypes = [C.c_void_p]
        L.hbh_error.restype = C.c_char_p; L.hbh_error.argtypes = [C.c_void_p]
        L.hbh_set_flux_bc.argtypes = [C.c_void_p, C.c_void_p]
        L.hbh_put.argtypes = [C.c_void_p] + [dp]*9
        L.hbh_fetch.argtypes = [C.c_void_p] + [dp]*5
        L.hbh_call.argtypes = [C.c_void_p, C.c_int, dp, dp]
        L.hbh_control.argtypes = [C.c_void_p, C.c_int, C.c_uint]
        L.hbh_flatten.argtypes = [C.c_void_p, ip, ip, ip, ip, ip]
        L.hbh_work_units.argtypes = [C.c_void_p, C.POINTER(C.c_longlong)]
        L.hbh_add_device_bc.argtypes = [C.c_void_p, C.c_int, ip, C.c_int, dp, C.c_int]
        L.hbh_apply_bcs.argtypes = [C.c_void_p, C.c_int]
        L.hbh_is_admissible.argtypes = [C.c_void_p, ip, ip]
        L.hbh_av_glue.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_int, dp, C.c_int, dp]
        L.hbh_face_permutation.argtypes = [C.c_int, C.c_int, ip, C.c_int, dp, C.c_char_p, C.c_int]
        L.hbh_host_state_bcs.argtypes = [C.c_void_p, C.c_int, dp, C.c_int, ip, C.c_int, C.c_int]
        L.hbh_host_flux_bcs.argtypes = [C.c_void_p, ip, C.c_int, C.c_int]
        L.hbh_set_devices.argtypes = [C.c_void_p, ip, C.c_int]
        L.hbh_set_element_coordinates.argtypes = [C.c_void_p, ip]
        L.hbh_element_owners.argtypes = [C.c_void_p, ip]
        L.hbh_transport_description.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        m = mesh
        self.m = m
        packed = np.ascontiguousarray(basis.packed())
        car, dfc, ref = (np.ascontiguousarray(a, dtype=np.int32) for a in (m.car_con, m.def_con, m.ref_face))
        pres = normal_present(m) if m.n_normal_slot else np.zeros(1, np.int32)
        self.h = L.hbh_create(m.n_dim, m.row_size, _d(packed), packed.size, m.n_car, m.n_def, m.n_face_slot, m.n_normal_slot,
                              _i(car), car.shape[0], _i(dfc), dfc.shape[0], _i(ref), ref.shape[0], _i(pres), seed)
        self._cb = None
        self.put(m, geometry=True)

    def close(self):
        if self.h:
            self.lib.hbh_destroy(self.h)
            self.h = None

    def _check(self, rc):
        if rc:
            raise RuntimeError(self.lib.hbh_error(self.h).decode())

    def put(self, m, geometry=False, wide=False):
        """host objects <- FlatMesh"""
        g = geometry
        self.lib.hbh_put(self.h, _d(m.elem_data), _d(m.nom_size) if g else None, _d(m.vertex_tss) if g else None,
                         _d(m.ref_normals) if g and m.n_def else None, _d(m.det) if g and m.n_def else None,
                         _d(m.normals) if g and m.n_normal_slot else None,
                         None if wide else _d(m.face_state), None if wide else _d(m.face_ldg), _d(m.face_wide) if wide else None)

    def put_uncert(self, uncert):
        u = np.ascontiguousarray(uncert, dtype=np.float64)
        self.lib.hbh_put_uncert.argtypes = [C.c_void_p, dp]
        self.lib.hbh_put_uncert(self.h, _d(u))

    def fetch(self, m, wide=False):
        """FlatMesh <- host objects"""
        self.lib.hbh_fetch(self.h, _d(m.elem_data), None if wide else _d(m.face_state), None if wide else _d(m.face_ldg),
                           _d(m.face_wide) if wide else None, _d(m.uncert))
        return m

    def call(self, name, *extra, dt=1., i_stage=0, compute_residual=False, use_filter=False):
        a = np.array([dt, i_stage, compute_residual, use_filter] + [float(x) for x in extra] + [0.]*4, dtype=np.float64)
        ret = C.c_double(0.)
        self._check(self.lib.hbh_call(self.h, FN[name], _d(a), C.byref(ret)))
        return ret.value

    def set_flux_bc(self, fn):
        self._cb = FLUX_CB(fn) if fn is not None else None
        self.lib.hbh_set_flux_bc(self.h, C.cast(self._cb, C.c_void_p) if self._cb else None)

    def control(self, what, arg=0):
        self._check(self.lib.hbh_control(self.h, what, arg))

    def set_sync_mode(self, mode): self.control(0, mode)
    def to_host(self, groups): self.control(1, groups)
    def to_device(self, groups): self.control(2, groups)
    def boundary_faces_to_host(self): self.control(3)
    def ghost_faces_to_device(self): self.control(4)
    def inside_state_faces_to_host(self): self.control(8, 1 | 1 << 2)   # (inside, state_half): what Solver::apply_state_bcs reads
    def ghost_state_faces_to_device(self): self.control(9, 2 | 1 << 2)  # (ghost, state_half): what it writes
    def invalidate(self): self.control(5)
    def release(self): self.control(6)

    def host_state_bcs(self, kind, params, def_con_index, n_threads=0):
        """the host loop of Solver::apply_state_bcs over the listed boundary connections (Freestream = 0, Nonpenetration = 2)"""
        p = np.ascontiguousarray(params if params is not None else np.zeros(1), dtype=np.float64)
        idx = np.ascontiguousarray(def_con_index, dtype=np.int32)
        self._check(self.lib.hbh_host_state_bcs(self.h, kind, _d(p), p.size, _i(idx), idx.size, n_threads))

    def host_flux_bcs(self, def_con_index, n_threads=0):
        idx = np.ascontiguousarray(def_con_index, dtype=np.int32)
        self._check(self.lib.hbh_host_flux_bcs(self.h, _i(idx), idx.size, n_threads))

    def inside_ldg_faces_to_host(self): self.control(8, 1 | 2 << 2)
    def ghost_ldg_faces_to_device(self): self.control(9, 2 | 2 << 2)
    def synchronize(self): self.control(7)

    def set_devices(self, devices):
        d = np.ascontiguousarray(devices, dtype=np.int32)
        self._check(self.lib.hbh_set_devices(self.h, _i(d), d.size))

    def set_element_coordinates(self, index):
        c = np.zeros((self.m.n_elem, 3), np.int32)
        c[:, :index.shape[1]] = index
        self._check(self.lib.hbh_set_element_coordinates(self.h, _i(c)))

    def element_owners(self):
        out = np.zeros(max(self.m.n_elem, 1), np.int32)
        self._check(self.lib.hbh_element_owners(self.h, _i(out)))
        return out[:self.m.n_elem]

    def transport_description(self):
        buf = C.create_string_buffer(512)
        self._check(self.lib.hbh_transport_description(self.h, buf, 512))
        return buf.value.decode()

    def add_device_bcs(self, mesh):
        """register every boundary condition of `mesh` on the device through hexed_b200::add_device_bc"""
        for bc in mesh.bcs:
            idx = np.ascontiguousarray(bc["con_index"], dtype=np.int32)
            params = np.ascontiguousarray(bc["params"], dtype=np.float64) if bc.get("params") is not None else np.zeros(1)
            n_params = params.size if bc.get("params") is not None else 0
            self._check(self.lib.hbh_add_device_bc(self.h, bc["kind"], _i(idx), idx.size, _d(params), n_params))

    def set_device_bc_params(self, bc_id, params):
        p = np.ascontiguousarray(params, dtype=np.float64)
        self.lib.hbh_set_device_bc_params.argtypes = [C.c_void_p, C.c_int, dp, C.c_int]
        self._check(self.lib.hbh_set_device_bc_params(self.h, int(bc_id), _d(p), p.size))

    def is_admissible(self):
        ok = np.zeros(1, np.int32); rec = np.zeros(max(self.m.n_elem, 1), np.int32)
        self._check(self.lib.hbh_is_admissible(self.h, _i(ok), _i(rec)))
        return bool(ok[0]), rec[:self.m.n_elem]

    def is_admissible_flag(self):
        """hexed_b200::is_admissible(mesh) without the per-element records (what Solver::update needs after a stage that went well)"""
        ok = np.zeros(1, np.int32)
        self._check(self.lib.hbh_is_admissible(self.h, _i(ok), None))
        return bool(ok[0])

    def av_glue(self, what, a=0., b=0., n=0, values=None):
        v = np.ascontiguousarray(values if values is not None else np.zeros(1), dtype=np.float64)
        out = np.zeros(1)
        self._check(self.lib.hbh_av_glue(self.h, what, float(a), float(b), int(n), _d(v), v.size if values is not None else 0, _d(out)))
        return float(out[0])

    def apply_state_bcs(self): self._check(self.lib.hbh_apply_bcs(self.h, 0))
    def apply_flux_bcs(self): self._check(self.lib.hbh_apply_bcs(self.h, 1))

    def flatten(self):
        m = self.m
        counts = np.zeros(6, np.int32)
        car = np.full_like(np.ascontiguousarray(m.car_con, dtype=np.int32), -7)
        dfc = np.full_like(np.ascontiguousarray(m.def_con, dtype=np.int32), -7)
        ref = np.full_like(np.ascontiguousarray(m.ref_face, dtype=np.int32), -7)
        bnd = np.full(max(dfc.shape[0], 1), -7, np.int32)
        self._check(self.lib.hbh_flatten(self.h, _i(counts), _i(car), _i(dfc), _i(ref), _i(bnd)))
        return counts, car, dfc, ref, bnd[:counts[4]]

    def work_units(self):
        out = (C.c_longlong*10)()
        self.lib.hbh_work_units(self.h, out)
        return list(out)

    def face_permutation(self, n_dim, row_size, direction, data, restore=False):
        d = np.array(direction, np.int32)
        err = C.create_string_buffer(256)
        rc = self.lib.hbh_face_permutation(n_dim, row_size, _i(d), int(restore), _d(data), err, 256)
        if rc:
            raise RuntimeError(err.value.decode())
        return data
