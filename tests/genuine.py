"""ctypes driver of oracle/ref_genuine_mesh.cpp: a mesh of the reference's GENUINE storage classes, stood up identically inside
  oracle/_ref/libhexed_ref.so                   (hexed::compute_* = the reference's own kernels) and
  oracle/_ref/libhexed_adapter_genuine[_emu].so (hexed::compute_* = hexed_b200/host/adapter.cpp compiled against the genuine headers).
TEST INFRASTRUCTURE."""
import ctypes as C
import os

import numpy as np

from pyoracle import ho_options, ho_transport, inviscid, HERE as ORACLE_DIR

REF = os.path.join(ORACLE_DIR, "_ref", "libhexed_ref.so")
ADAPTER = os.path.join(ORACLE_DIR, "_ref", "libhexed_adapter_genuine.so")
ADAPTER_EMU = os.path.join(ORACLE_DIR, "_ref", "libhexed_adapter_genuine_emu.so")
dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int)


def _d(a):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(dp)


class GenuineLib:
    def __init__(self, path):
        self.lib = L = C.CDLL(path)
        L.hr_gm_create.argtypes = [C.c_int, C.c_int, C.c_int, ip, C.c_double]; L.hr_gm_create.restype = C.c_void_p
        L.hr_gm_destroy.argtypes = [C.c_void_p]; L.hr_gm_destroy.restype = None
        L.hr_gm_error.argtypes = [C.c_void_p]; L.hr_gm_error.restype = C.c_char_p
        L.hr_gm_counts.argtypes = [C.c_void_p, ip]; L.hr_gm_counts.restype = None
        L.hr_gm_blob_size.argtypes = [C.c_void_p]; L.hr_gm_blob_size.restype = C.c_long
        for f in ("hr_gm_export", "hr_gm_import", "hr_gm_positions", "hr_gm_bc_freestream"):
            getattr(L, f).argtypes = [C.c_void_p, dp]; getattr(L, f).restype = None
        for f in ("hr_gm_get_slots", "hr_gm_set_slots"):
            getattr(L, f).argtypes = [C.c_void_p, C.c_int, C.c_int, dp]; getattr(L, f).restype = None
        L.hr_gm_bc_copy.argtypes = [C.c_void_p]; L.hr_gm_bc_copy.restype = None
        L.hr_gm_calc_jacobian.argtypes = [C.c_void_p]
        L.hr_gm_compute_write_face.argtypes = [C.c_void_p]
        L.hr_gm_compute_euler.argtypes = [C.c_void_p, ho_options]
        L.hr_gm_compute_navier_stokes.argtypes = [C.c_void_p, ho_options, ho_transport, ho_transport]
        L.hr_gm_max_dt.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_int, ho_transport, ho_transport, dp]
        L.hr_gm_stabilizing_art_visc.argtypes = [C.c_void_p, C.c_double]
        L.hr_gm_work_units.argtypes = [C.c_void_p, ip]; L.hr_gm_work_units.restype = None
        self.has_adapter = bool(L.hr_gm_adapter_present())
        if self.has_adapter:
            L.hr_gm_set_sync_mode.argtypes = [C.c_int]; L.hr_gm_set_sync_mode.restype = None
            for f in ("hr_gm_to_host", "hr_gm_to_device", "hr_gm_boundary_faces_to_host", "hr_gm_ghost_faces_to_device"):
                getattr(L, f).argtypes = [C.c_void_p]
            L.hr_gm_flatten_counts.argtypes = [C.c_void_p, ip]
            L.hr_gm_release.restype = None


class GenuineMesh:
    def __init__(self, glib, n_dim, row_size, n, cell_kind, warp=0.1):
        self.g, self.L = glib, glib.lib
        self.nd, self.rs, self.nq, self.nv = n_dim, row_size, row_size**n_dim, n_dim + 2
        kind = np.ascontiguousarray(cell_kind, dtype=np.int32).ravel()
        assert kind.size == n**n_dim
        self.h = self.L.hr_gm_create(n_dim, row_size, n, kind.ctypes.data_as(ip), warp)
        assert self.h, "hr_gm_create failed"
        c = (C.c_int*6)()
        self.L.hr_gm_counts(self.h, c)
        self.n_car, self.n_def, self.n_car_con, self.n_def_con, self.n_ref, self.n_bc = list(c)
        self.n_elem = self.n_car + self.n_def

    def close(self):
        if self.h:
            self.L.hr_gm_destroy(self.h); self.h = None

    def _check(self, rc):
        if rc == 1:
            raise RuntimeError("demand for invalid kernel")
        if rc:
            raise RuntimeError(self.L.hr_gm_error(self.h).decode())

    def export(self):
        b = np.zeros(self.L.hr_gm_blob_size(self.h))
        self.L.hr_gm_export(self.h, _d(b))
        return b

    def load(self, blob):
        assert blob.size == self.L.hr_gm_blob_size(self.h)
        self.L.hr_gm_import(self.h, _d(np.ascontiguousarray(blob)))

    def positions(self):
        p = np.zeros((self.n_elem, self.nd, self.nq))
        self.L.hr_gm_positions(self.h, _d(p))
        return p

    def slots(self, first, n):
        a = np.zeros((self.n_elem, n, self.nq))
        self.L.hr_gm_get_slots(self.h, first, n, _d(a))
        return a

    def set_slots(self, first, a):
        a = np.ascontiguousarray(a, dtype=np.float64)
        assert a.shape[0] == self.n_elem and a.shape[2] == self.nq
        self.L.hr_gm_set_slots(self.h, first, a.shape[1], _d(a))

    def state(self):
        return self.slots(0, self.nv)

    def calc_jacobian(self):
        self._check(self.L.hr_gm_calc_jacobian(self.h))

    def bc_copy(self):
        self.L.hr_gm_bc_copy(self.h)

    def bc_freestream(self, fs):
        self.L.hr_gm_bc_freestream(self.h, _d(np.ascontiguousarray(fs, dtype=np.float64)))

    def compute_write_face(self):
        self._check(self.L.hr_gm_compute_write_face(self.h))

    def compute_euler(self, dt, i_stage, compute_residual=False, use_filter=False):
        self._check(self.L.hr_gm_compute_euler(self.h, ho_options(dt, i_stage, int(compute_residual), int(use_filter))))

    def compute_navier_stokes(self, visc, cond, dt, i_stage=0, compute_residual=False, use_filter=False):
        self._check(self.L.hr_gm_compute_navier_stokes(self.h, ho_options(dt, i_stage, int(compute_residual), int(use_filter)), visc, cond))

    def max_dt(self, pde, sc, sd, local_time, visc=None, cond=None):
        out = C.c_double(0.)
        self._check(self.L.hr_gm_max_dt(self.h, pde, sc, sd, int(local_time), visc or inviscid(), cond or inviscid(), C.byref(out)))
        return out.value

    def stabilizing_art_visc(self, char_speed):
        self._check(self.L.hr_gm_stabilizing_art_visc(self.h, char_speed))

    def work_units(self):
        c = (C.c_int*5)()
        self.L.hr_gm_work_units(self.h, c)
        return list(c)

    # adapter build only
    def to_host(self):
        self._check(self.L.hr_gm_to_host(self.h))

    def to_device(self):
        self._check(self.L.hr_gm_to_device(self.h))

    def boundary_faces_to_host(self):
        self._check(self.L.hr_gm_boundary_faces_to_host(self.h))

    def ghost_faces_to_device(self):
        self._check(self.L.hr_gm_ghost_faces_to_device(self.h))

    def flatten_counts(self):
        c = (C.c_int*8)()
        self._check(self.L.hr_gm_flatten_counts(self.h, c))
        return list(c)
