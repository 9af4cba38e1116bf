"""shared helpers for the parity tests: run the same call sequence on the oracle and on a Device, compare."""
import numpy as np

import hexed_b200 as hb
from hexed_b200 import mesh as M
from hexed_b200.kernels import Device
from pyoracle import EULER
from hexed_b200.cases import density_wave, freestream_state  # noqa: F401

STATE_TOL = 1e-11   # north-star: state after N stages within relative L2 <= 1e-11 of the reference CPU kernels
MAX_DT_TOL = 1e-13  # north-star: max_dt within 1e-13 relative


def rel_l2(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b))/max(np.linalg.norm(b), 1e-300))


def run_euler_pair(oracle, lib_path, mesh, basis, n_steps=2, local_time=False, use_filter=False, safety=0.7):
    """advance `mesh` n_steps (each: max_dt + 2 stages with ghost-state BCs) on both implementations.
    Returns (device result mesh, oracle result mesh, list of (dt_device, dt_oracle))."""
    ref = mesh.copy()
    dev = Device(mesh.n_dim, mesh.row_size, basis, lib_path=lib_path).load_mesh(mesh)
    dts = []
    for _ in range(n_steps):
        dt_o = oracle.max_dt(EULER, basis, ref, safety, safety, local_time)
        dt_d = dev.max_dt_euler(safety, safety, local_time)
        dts.append((dt_d, dt_o))
        for stage in (0, 1):
            oracle.apply_state_bcs(ref)
            oracle.compute_euler(basis, ref, dt=dt_o, i_stage=stage, use_filter=use_filter)
            dev.apply_state_bcs()
            dev.compute_euler(dt=dt_o, i_stage=stage, use_filter=use_filter)
    out = mesh.copy()
    dev.sync_to_host(out)
    launches = dev.launch_count()
    dev.close()
    return out, ref, dts, launches


def assert_euler_parity(out, ref, dts):
    for dt_d, dt_o in dts:
        assert abs(dt_d - dt_o) <= MAX_DT_TOL*abs(dt_o), (dt_d, dt_o)
    assert rel_l2(out.state(), ref.state()) <= STATE_TOL
    assert rel_l2(out.cache(), ref.cache()) <= 1e-10  # cancellation-prone residual difference, looser by design
    assert rel_l2(out.face_state, ref.face_state) <= STATE_TOL
    assert rel_l2(out.tss(), ref.tss()) <= MAX_DT_TOL
