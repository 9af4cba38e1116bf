"""The drop-in boundary without a GPU: libhexed_b200.so loads and exports every function include/hexed_b200.h declares (and
the ctypes mirror binds exactly that set), compute calls fail loudly without a device, and the C++ adapter library defines
the reference's own entry points (include/kernels.hpp:22-42, include/stabilizing_art_visc.hpp:13)."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

import hexed_b200 as hb
from hexed_b200.kernels import LIB_PATH, SIGNATURES, load_library, Device

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "hexed_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return set(re.findall(r"\b(hexed_b200_[a-z0-9_]+)\s*\(", text)) - {"hexed_b200_callback"}


def test_header_symbols_exported_and_bound():
    names = declared_functions()
    assert len(names) >= 50
    lib = load_library()
    for n in names:
        getattr(lib, n)  # AttributeError if the library does not export it
    assert names == set(SIGNATURES), (names ^ set(SIGNATURES))


def test_no_device_fails_loudly():
    """there is no CPU fallback: without a CUDA device, creating a context is an error (with a GPU present it succeeds)"""
    lib = load_library()
    n = ctypes.c_int(-1)
    lib.hexed_b200_device_count(ctypes.byref(n))
    if n.value > 0:
        pytest.skip("a CUDA device is visible")
    with pytest.raises(RuntimeError, match="no CUDA device"):
        Device(3, 6, hb.gauss_legendre(6))


def test_invalid_kernel_demand():
    """kernel_factory's error for unsupported (n_dim, row_size), include/kernel_factory.hpp:114-116"""
    for nd, rs in ((0, 6), (4, 6), (3, 1), (3, 9)):
        with pytest.raises(RuntimeError, match="demand for invalid kernel"):
            Device(nd, rs, None)


def test_permutation_indices_need_no_device():
    lib = load_library()
    out = np.zeros(36, np.int32)
    d = (ctypes.c_int*4)(0, 2, 1, 1)
    assert lib.hexed_b200_face_permutation_indices(3, 6, d, out.ctypes.data_as(ctypes.POINTER(ctypes.c_int))) == 0
    assert sorted(out.tolist()) == list(range(36)) and not np.array_equal(out, np.arange(36))


def test_adapter_defines_reference_entry_points():
    path = os.path.join(ROOT, "hexed_b200", "libhexed_b200_host.so")
    assert os.path.exists(path), "run __graft_entry__.build()"
    syms = subprocess.run(["nm", "-DC", "--defined-only", path], check=True, capture_output=True, text=True).stdout
    for fn in ("compute_euler", "compute_advection", "compute_navier_stokes", "compute_smooth_av", "compute_fix_therm_admis",
               "max_dt_euler", "max_dt_navier_stokes", "max_dt_advection", "max_dt_smooth_av", "max_dt_fix_therm_admis",
               "compute_prolong", "compute_restrict", "compute_prolong_advection", "face_permutation", "compute_write_face",
               "compute_write_face_advection", "compute_write_face_smooth_av", "stabilizing_art_visc"):
        assert re.search(r" T hexed::%s\(" % fn, syms), fn


def test_new_entry_points_are_declared_everywhere():
    """the SURVEY section 8 f entry points exist in the header (and, by test_header_symbols_exported_and_bound, in the library and the
    ctypes table)"""
    names = declared_functions()
    for name in ("hexed_b200_is_admissible", "hexed_b200_download_record", "hexed_b200_set_jacobian", "hexed_b200_calc_shared_normals",
                 "hexed_b200_vertex_topology", "hexed_b200_share_vertex_data", "hexed_b200_fix_admis_spread", "hexed_b200_av_scale_velocity",
                 "hexed_b200_av_project_forcing", "hexed_b200_av_finish", "hexed_b200_interp_vertices", "hexed_b200_av_swap",
                 "hexed_b200_apply_aux_bcs", "hexed_b200_av_elwise_ramp", "hexed_b200_av_elwise_forcing", "hexed_b200_av_elwise_vertices"):
        assert name in names and name in SIGNATURES
