"""Launch-bound meshes (the 2-D sample cases of BASELINE.json, e.g. samples/vortex: 16 x 16 quads): microseconds per step of Solver::update's
inviscid loop (max_dt + 2 x (characteristic ghost fill + compute_euler)) three ways on the device -- call by call through the
reference-shaped entry points (one synchronisation per step for dt), hexed_b200_update_euler with the time step kept on the device, and the
same with the step replayed from a CUDA graph -- next to the CPU oracle on the host cores. One JSON line per mesh."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import hexed_b200 as hb  # noqa: E402
from hexed_b200 import mesh as M  # noqa: E402
from hexed_b200.kernels import Device  # noqa: E402
from pyoracle import Oracle, EULER  # noqa: E402
import test_config_vortex as V  # noqa: E402


def main():
    rs = 6
    basis = hb.gauss_legendre(rs)
    oracle = Oracle(lib="liboracle_fast.so")
    for n in (16, 64, 256):
        m = M.box_mesh(2, rs, n, basis, deformed=False, bc_kind=M.BC_RIEMANN_INVARIANTS, bc_params=V.FS)
        m.state()[:] = V.vortex(np.asarray(m.qpoint_pos) - 0.5, 0.)
        oracle.compute_write_face(basis, m)
        ref = m.copy()
        dev = Device(2, rs, basis).load_mesh(m)
        n_steps = 2000 if n <= 64 else 300

        def call_by_call(k):
            for _ in range(k):
                dt = dev.max_dt_euler(0.01, 0.01, False)
                for stage in (0, 1):
                    dev.apply_state_bcs(); dev.compute_euler(dt=dt, i_stage=stage)
        res = {"workload": "samples/vortex class: %d x %d Cartesian quads, row size 6, Euler, characteristic BCs" % (n, n), "elements": n*n, "steps": n_steps}
        for name, fn in (("call_by_call", call_by_call), ("device_dt", lambda k: dev.update_euler(0.01, k, False)), ("device_dt_graph", lambda k: dev.update_euler(0.01, k, True))):
            fn(20); torch.cuda.synchronize()
            t = time.perf_counter(); fn(n_steps); dev.synchronize(); el = time.perf_counter() - t
            res[name + "_us_per_step"] = el/n_steps*1e6
        cpu_steps = max(20, n_steps//10)
        t = time.perf_counter()
        for _ in range(cpu_steps):
            dt = oracle.max_dt(EULER, basis, ref, 0.01, 0.01, False)
            for stage in (0, 1):
                oracle.apply_state_bcs(ref); oracle.compute_euler(basis, ref, dt=dt, i_stage=stage)
        res["cpu_oracle_us_per_step"] = (time.perf_counter() - t)/cpu_steps*1e6
        res["cpu_threads"] = oracle.num_threads()
        res["dof_stage_per_s_graph"] = n*n*4*rs*rs*2/(res["device_dt_graph_us_per_step"]*1e-6)
        print(json.dumps(res))
        dev.close()
    # the cylinder class: 2-D viscous, deformed quads, wall boundary conditions, Sutherland air
    from hexed_b200 import kernels as K
    from util import density_wave
    visc, cond = K.sutherland(1.716e-5, 273., 111.), K.sutherland(.0241, 273., 194.)
    for n in (16, 64):
        m = M.box_mesh(2, rs, n, basis, deformed=True, bc_kind=M.BC_NONPENETRATION, with_ldg=True)
        density_wave(m, basis)
        oracle.compute_write_face(basis, m)
        dev = Device(2, rs, basis).load_mesh(m)
        n_steps = 1000

        def call_by_call_ns(k):
            for _ in range(k):
                dt = dev.max_dt_navier_stokes(0.1, 0.1, False, visc, cond)
                dev.apply_state_bcs(); dev.compute_navier_stokes(dev.apply_flux_bcs, visc, cond, dt=dt, i_stage=0)
                dev.apply_state_bcs(); dev.compute_euler(dt=dt, i_stage=1)
        res = {"workload": "samples/cylinder class: %d x %d deformed quads, row size 6, Navier-Stokes, wall BCs" % (n, n), "elements": n*n, "steps": n_steps}
        for name, fn in (("call_by_call", call_by_call_ns), ("device_dt", lambda k: dev.update_navier_stokes(0.1, visc, cond, k, False)),
                         ("device_dt_graph", lambda k: dev.update_navier_stokes(0.1, visc, cond, k, True))):
            fn(20); torch.cuda.synchronize()
            t = time.perf_counter(); fn(n_steps); dev.synchronize(); el = time.perf_counter() - t
            res[name + "_us_per_step"] = el/n_steps*1e6
        print(json.dumps(res))
        dev.close()


if __name__ == "__main__":
    main()
