"""BASELINE config C1 (reference samples/vortex/run.hil): 2-D isentropic vortex, inviscid Euler, 16 x 16 Cartesian quads at row size 6,
global time step with max_safety 0.01, `characteristic` (Riemann_invariants) boundary conditions, sea-level freestream at 100 m/s,
initial condition `vortex` of include/Case.hil:69-82. The synthetic box spans [0, 1]^2 instead of [-1, 1]^2, so lengths and times are
halved (the Euler equations are scale invariant): vortex radius 0.05, centre (0.5, 0.5).

The whole time loop of Solver::update runs on the device (max_dt, Riemann-invariant ghost fill, two stages) next to the same loop on
the oracle; checked: parity of the state after many steps and the L^2 density error against the advected analytic vortex, which is what
the sample's run.hil prints."""
import numpy as np
import pytest

import hexed_b200 as hb
from hexed_b200 import mesh as M
from hexed_b200.kernels import Device
from pyoracle import EULER
from util import rel_l2

HEAT_RAT = 1.4
RHO, P, SPEED = 1.225, 101325., 100.  # sea-level ISA (altitude = 0.), freestream_speed = 100.
FS = np.array([RHO*SPEED, 0., RHO, P/(HEAT_RAT - 1.) + .5*RHO*SPEED**2])
RADIUS, NONDIM_VELOC = 0.05, 0.3  # vortex_argmax_radius (halved with the domain), vortex_nondim_veloc


def vortex(pos, time):
    """include/Case.hil:69-82; pos (..., 2, nq) relative to the vortex centre at time 0 -> state (..., 4, nq)"""
    e_int = FS[3]/FS[2] - .5*(FS[0]**2 + FS[1]**2)/FS[2]**2
    sound = np.sqrt(HEAT_RAT*(HEAT_RAT - 1.)*e_int)
    x = pos[..., 0, :] - FS[0]/FS[2]*time
    y = pos[..., 1, :] - FS[1]/FS[2]*time
    gauss = NONDIM_VELOC*np.exp((1. - (x*x + y*y)/RADIUS**2)/2.)
    scalar = 1. - (HEAT_RAT - 1.)/2.*gauss**2
    mass = FS[2]*scalar**(1./(HEAT_RAT - 1.))
    m0 = (FS[0]/FS[2] - y/RADIUS*gauss*sound)*mass
    m1 = (FS[1]/FS[2] + x/RADIUS*gauss*sound)*mass
    energy = e_int*scalar*mass + .5*(m0*m0 + m1*m1)/mass
    return np.stack([m0, m1, mass, energy], axis=-2)


def run_vortex(oracle, lib, n_steps, n=16, rs=6, safety=0.01):
    basis = hb.gauss_legendre(rs)
    m = M.box_mesh(2, rs, n, basis, deformed=False, bc_kind=M.BC_RIEMANN_INVARIANTS, bc_params=FS)
    pos = np.asarray(m.qpoint_pos) - 0.5
    m.state()[:] = vortex(pos, 0.)
    oracle.compute_write_face(basis, m)
    ref = m.copy()
    dev = Device(2, rs, basis, lib_path=lib).load_mesh(m)
    t_dev = t_ref = 0.
    worst_dt = 0.
    for _ in range(n_steps):
        dt_o = oracle.max_dt(EULER, basis, ref, safety, safety, False)
        dt_d = dev.max_dt_euler(safety, safety, False)
        worst_dt = max(worst_dt, abs(dt_d/dt_o - 1.))
        for stage in (0, 1):
            oracle.apply_state_bcs(ref); oracle.compute_euler(basis, ref, dt=dt_o, i_stage=stage)
            dev.apply_state_bcs(); dev.compute_euler(dt=dt_d, i_stage=stage)   # each side marches with its OWN time step
        t_ref += dt_o; t_dev += dt_d
    assert dev.is_admissible()
    out = m.copy()
    dev.sync_to_host(out)
    dev.close()
    w = np.asarray(basis.weight)
    wq = (w[:, None]*w[None, :]).reshape(-1)*(1./n)**2

    def l2_density_error(mesh, t):  # run.hil: integrand_field errsq = (density - $vortex_mass)^2; integral^0.5
        exact = vortex(pos, t)[:, 2]
        return float(np.sqrt((((mesh.state()[:, 2] - exact)**2)*wq[None, :]).sum()))
    amplitude = float(np.sqrt((((vortex(pos, 0.)[:, 2] - RHO)**2)*wq[None, :]).sum()))
    return dict(out=out, ref=ref, t_dev=t_dev, t_ref=t_ref, worst_dt=worst_dt, err_dev=l2_density_error(out, t_dev),
                err_ref=l2_density_error(ref, t_ref), amplitude=amplitude)


def check(r):
    assert r["worst_dt"] <= 1e-13                                   # north-star: max_dt within 1e-13 relative, every step
    assert abs(r["t_dev"]/r["t_ref"] - 1.) <= 1e-13
    assert rel_l2(r["out"].state(), r["ref"].state()) <= 1e-11      # state after N stages
    assert np.isfinite(r["err_dev"]) and r["err_dev"] <= 2e-3*r["amplitude"]  # the vortex is advected, not smeared
    assert abs(r["err_dev"] - r["err_ref"]) <= 1e-9*r["amplitude"]


def test_vortex_emulated(oracle, emu_lib):
    check(run_vortex(oracle, emu_lib, n_steps=12))


@pytest.mark.gpu
def test_vortex_config_on_b200(oracle, gpu_lib):
    """2 000 steps (4 000 stages) of the sample's time loop, entirely on the device"""
    r = run_vortex(oracle, gpu_lib, n_steps=2000)
    check(r)
    assert r["t_dev"] > 1e-5
