"""BASELINE configs C2 / C3 as parity cases on synthetic meshes of the same class:

C3 (samples/cylinder): 2-D viscous Navier-Stokes, Sutherland viscosity / conductivity (include/Case.hil:83-90), isothermal No_slip wall
   on one side and `characteristic` (Riemann_invariants) far field on the others, deformed quads, local time stepping -- every
   ghost fill and flux condition on the device, several steps of Solver::update's loop.
C2 (samples/naca0012): quads with hanging nodes, Euler stages followed by the artificial-viscosity smoothness update
   (update_art_visc_smoothness) and a viscous stage that uses the new coefficient -- the shock-capturing cycle of that case.
"""
import numpy as np
import pytest

import hexed_b200 as hb
import pyoracle
from hexed_b200 import mesh as M
from hexed_b200.kernels import Device, BC_MODE_ADVECTION, BC_MODE_COPY_STATE, BC_MODE_NEGATE_FLUX
from hexed_b200 import kernels as K
from pyoracle import EULER, NAVIER_STOKES, ADVECTION, SMOOTH_AV
from util import rel_l2, density_wave, freestream_state, assert_pde_parity, STATE_TOL


def split_bcs_by_side(mesh, kinds):
    """replace the single boundary condition of a box mesh by one per (dimension, sign) side; kinds[(d, sign)] = (kind, params)"""
    src = mesh.bcs[0]
    nd = mesh.n_dim
    side = src["inside_slot"] % (2*nd)
    bcs = []
    for (d, sign), (kind, params) in kinds.items():
        sel = np.nonzero(side == 2*d + sign)[0]
        bcs.append(dict(kind=kind, params=params, **{k: np.ascontiguousarray(src[k][sel]) for k in ("inside_slot", "ghost_slot", "normal_slot", "con_index")}))
    assert sum(b["inside_slot"].size for b in bcs) == src["inside_slot"].size
    mesh.bcs = bcs
    return mesh


def run_cylinder_class(oracle, lib, n, rs, n_steps):
    nd = 2
    basis = hb.gauss_legendre(rs)
    fs = freestream_state(nd, mach=0.5)
    m = M.box_mesh(nd, rs, n, basis, deformed=True, bc_kind=M.BC_COPY, with_ldg=True)
    wall = (M.BC_NO_SLIP, M.no_slip_params(M.THERMAL_ENERGY, 2.2e5))  # isothermal wall: prescribed specific energy
    far = (M.BC_RIEMANN_INVARIANTS, fs)
    split_bcs_by_side(m, {(0, 0): far, (0, 1): far, (1, 0): wall, (1, 1): far})
    density_wave(m, basis, mach=0.5)
    oracle.compute_write_face(basis, m)
    ref = m.copy()
    dev = Device(nd, rs, basis, lib_path=lib).load_mesh(m)
    visc_o, cond_o = pyoracle.sutherland(1.716e-5, 273., 111.), pyoracle.sutherland(.0241, 273., 194.)
    visc_d, cond_d = K.sutherland(1.716e-5, 273., 111.), K.sutherland(.0241, 273., 194.)
    dts = []
    for _ in range(n_steps):
        dt_o = oracle.max_dt(NAVIER_STOKES, basis, ref, 0.3, 0.3, True, visc_o, cond_o)
        dt_d = dev.max_dt_navier_stokes(0.3, 0.3, True, visc_d, cond_d)
        dts.append((dt_d, dt_o))
        oracle.apply_state_bcs(ref); dev.apply_state_bcs()
        oracle.compute_navier_stokes(basis, ref, lambda: oracle.apply_flux_bcs(ref), visc_o, cond_o, dt=dt_o, i_stage=0)
        dev.compute_navier_stokes(dev.apply_flux_bcs, visc_d, cond_d, dt=dt_o, i_stage=0)
        oracle.apply_state_bcs(ref); dev.apply_state_bcs()
        oracle.compute_euler(basis, ref, dt=dt_o, i_stage=1)
        dev.compute_euler(dt=dt_o, i_stage=1)
    ok = dev.is_admissible()
    out = m.copy()
    dev.sync_to_host(out)
    dev.close()
    assert ok == oracle.is_admissible(ref)[0]
    assert_pde_parity(out, ref, dts)
    assert rel_l2(out.state(), m.state()) > 1e-6  # the flow really moved


def test_cylinder_class_emulated(oracle, emu_lib):
    run_cylinder_class(oracle, emu_lib, n=4, rs=3, n_steps=2)


@pytest.mark.gpu
def test_cylinder_class_on_b200(oracle, gpu_lib):
    run_cylinder_class(oracle, gpu_lib, n=24, rs=6, n_steps=10)


def run_naca_class(oracle, lib, n, rs, n_cycles):
    nd = 2
    basis = hb.gauss_legendre(rs)
    refine = np.zeros((n,)*nd, bool)
    refine[n//3:2*n//3, n//3:2*n//3] = True   # a refined patch: hanging-node faces all around it
    m = M.refined_box_mesh(nd, rs, n, basis, refine, bc_kind=M.BC_NONPENETRATION, with_ldg=True)
    assert m.ref_face.shape[0] > 0
    m.face_wide = np.zeros((m.n_face_slot, (nd + rs)*m.nfq))
    density_wave(m, basis)
    # a steep (under-resolved) front in velocity, the thing the smoothness indicator exists to find: on a smooth flow the
    # artificial viscosity it returns is round-off sized and there would be nothing to compare
    x = np.asarray(m.qpoint_pos)
    front = 1. + 0.3*np.tanh((x[:, 0] + 0.5*x[:, 1] - 0.7)/0.01)
    m.state()[:, :nd] *= front[:, None, :]   # momentum only: the indicator advects along the velocity, so the velocity must jump
    m.elem_data[:, nd + 9:nd + 9 + rs] = 1.
    oracle.compute_write_face(basis, m); oracle.compute_prolong(basis, m)
    ref = m.copy()
    dev = Device(nd, rs, basis, lib_path=lib).load_mesh(m)
    w = np.asarray(basis.weight); orth = np.asarray(basis.orthogonal).reshape(rs, rs)[rs - 1]
    advect_length, n_real = 0.5/n, 3
    diff_time = 0.5*advect_length**2/n_real
    mult, us_max = 2.*advect_length, advect_length*0.5*np.sqrt(2*2.5e5/1.2)
    inviscid_o, inviscid_d = pyoracle.inviscid(), K.inviscid()
    dts = []
    for _ in range(n_cycles):
        # two Euler stages
        dt_o = oracle.max_dt(EULER, basis, ref, 0.5, 0.5, False); dt_d = dev.max_dt_euler(0.5, 0.5, False)
        dts.append((dt_d, dt_o))
        for stage in (0, 1):
            oracle.apply_state_bcs(ref); oracle.compute_euler(basis, ref, dt=dt_o, i_stage=stage)
            dev.apply_state_bcs(); dev.compute_euler(dt=dt_o, i_stage=stage)
        # update_art_visc_smoothness (reference src/Solver.cpp:457-581): row_size advection iterations, n_real + 1 smoothing sweeps
        pyoracle.av_scale_velocity(ref); dev.av_scale_velocity()
        oracle.compute_write_face(basis, ref); oracle.compute_prolong(basis, ref); dev.compute_write_face(); dev.compute_prolong()
        oracle.max_dt(ADVECTION, basis, ref, 0.5, 1., True, advect_length=advect_length); dev.max_dt_advection(0.5, 1., True, advect_length)
        for _it in range(rs):  # every iteration raises the polynomial degree of the advection states in the node variable by one: the
            # projection on the Legendre polynomial of degree row_size - 1 is round-off until row_size - 1 of them have run
            oracle.compute_write_face(basis, ref, pde=ADVECTION); oracle.compute_prolong(basis, ref, pde=ADVECTION)
            dev.compute_write_face_advection(); dev.compute_prolong_advection()
            for i in (0, 1):
                pyoracle.apply_aux_bcs(ref, BC_MODE_ADVECTION); oracle.compute_advection(basis, ref, advect_length, dt=1., i_stage=i)
                dev.apply_aux_bcs(BC_MODE_ADVECTION); dev.compute_advection(advect_length, dt=1., i_stage=i)
        pyoracle.av_project_forcing(ref, w, orth); dev.av_project_forcing(w, orth)
        oracle.max_dt(SMOOTH_AV, basis, ref, 1., 0.4, True); dev.max_dt_smooth_av(1., 0.4, True)
        oracle.compute_write_face(basis, ref, pde=SMOOTH_AV); oracle.compute_prolong(basis, ref)
        dev.compute_write_face_smooth_av(); dev.compute_prolong()
        for _sweep in range(n_real + 1):  # each sweep carries the forcing one real time step further down the chain of n_real
            pyoracle.apply_aux_bcs(ref, BC_MODE_COPY_STATE); dev.apply_aux_bcs(BC_MODE_COPY_STATE)
            oracle.compute_smooth_av(basis, ref, lambda: pyoracle.apply_aux_bcs(ref, BC_MODE_NEGATE_FLUX), diff_time, 1., dt=1., i_stage=0)
            dev.compute_smooth_av(lambda: dev.apply_aux_bcs(BC_MODE_NEGATE_FLUX), diff_time, 1., dt=1., i_stage=0)
        want = pyoracle.av_finish(ref, mult, us_max, n_real, w); got = dev.av_finish(mult, us_max, n_real, w)
        assert abs(got - want) <= 1e-10*want + 1e-14*us_max  # on a smooth flow the residual is round-off sized
        oracle.compute_write_face(basis, ref); oracle.compute_prolong(basis, ref); dev.compute_write_face(); dev.compute_prolong()
        # a stage with the artificial viscosity switched on (use_ldg() becomes true, src/Solver.cpp:117-120)
        dt_o = oracle.max_dt(NAVIER_STOKES, basis, ref, 0.3, 0.3, False, inviscid_o, inviscid_o)
        dt_d = dev.max_dt_navier_stokes(0.3, 0.3, False, inviscid_d, inviscid_d)
        dts.append((dt_d, dt_o))
        oracle.apply_state_bcs(ref); dev.apply_state_bcs()
        oracle.compute_navier_stokes(basis, ref, lambda: oracle.apply_flux_bcs(ref), inviscid_o, inviscid_o, dt=dt_o, i_stage=0)
        dev.compute_navier_stokes(dev.apply_flux_bcs, inviscid_d, inviscid_d, dt=dt_o, i_stage=0)
    out = m.copy()
    dev.sync_to_host(out)
    dev.close()
    for dt_d, dt_o in dts:
        assert abs(dt_d - dt_o) <= 1e-13*abs(dt_o)
    assert rel_l2(out.state(), ref.state()) <= STATE_TOL
    assert rel_l2(out.face_state, ref.face_state) <= STATE_TOL
    # the LDG faces hold the artificial-viscosity flux, proportional to the coefficient discussed below: held to the flux scale
    assert np.abs(out.face_ldg - ref.face_ldg).max() <= STATE_TOL*np.abs(ref.face_state).max()
    assert rel_l2(out.elem_data[:, nd + 9:nd + 9 + rs], ref.elem_data[:, nd + 9:nd + 9 + rs]) <= STATE_TOL  # advection states
    # The smoothness indicator is the SQUARE of the projection of the advection states (all ~1) on the top Legendre mode: forcing =
    # proj^2 * 2E/rho with |proj| ~ 1e-6 here, so round-off eps*|adv| in the projection is a relative 1e-9 in the forcing and in the
    # AV coefficient that is smoothed from it -- between two CPU builds of the oracle just as between oracle and device. Parity is
    # therefore asked of what the arithmetic actually produces to working precision, the projection: sqrt(forcing) ~ |proj|*sqrt(2E/rho).
    spec = float(np.sqrt(2*ref.state()[:, nd + 1]/ref.state()[:, nd]).max())
    for lo, hi, scale in ((nd + 5, nd + 9, 1.), (nd + 3, nd + 4, mult)):
        a, b = np.sqrt(np.abs(out.elem_data[:, lo:hi])/scale), np.sqrt(np.abs(ref.elem_data[:, lo:hi])/scale)
        assert np.isfinite(b).all() and np.abs(a - b).max() <= 1e-12*spec, (lo, np.abs(a - b).max(), spec)
    assert out.elem_data[:, nd + 3].max() > 1e-12*us_max and ref.elem_data[:, nd + 5].max() > 0.


def test_naca_class_emulated(oracle, emu_lib):
    run_naca_class(oracle, emu_lib, n=3, rs=3, n_cycles=1)


@pytest.mark.gpu
def test_naca_class_on_b200(oracle, gpu_lib):
    run_naca_class(oracle, gpu_lib, n=12, rs=6, n_cycles=3)


def test_navier_stokes_max_dt_screen_error_budget():
    """the single-precision screen of the Navier-Stokes max_dt (PdeNs::screen_dt, hexed_b200/csrc/pde.cuh) skips the FP64 evaluation of
    points whose estimate is more than 1e-4 above the running minimum; that is only sound if the estimate is good to well under half
    of that wherever it is trusted (internal energy >= 3 % of the total). numpy float32 restatement of the estimate against float64 on
    random states: Mach 0 to 30, densities over 3 decades, Sutherland air."""
    rng = np.random.default_rng(7)
    n = 400000
    f32 = np.float32
    rho = 10.**rng.uniform(-2, 1, n); sound = 10.**rng.uniform(1.5, 3., n); mach = rng.uniform(0., 30., n)*rng.integers(0, 2, n)
    mom = rho*sound*mach; en = rho*sound**2/(1.4*0.4) + 0.5*mom**2/rho
    av = 10.**rng.uniform(-6, 1, (2, n))
    h = 10.**rng.uniform(-4, 0, n)
    inv_c, inv_d = 1/0.3, 1/0.05
    gm1_over_r = 0.4/287.05287
    def transport(ref_val, ref_temp, offset, sqrt_temp, temp, dt):
        r = sqrt_temp*dt(1/np.sqrt(ref_temp))
        return dt(ref_val*(ref_temp + offset))*(r*r*r)/(temp + dt(offset))
    def local_dt(dt):
        rho_, mom_, en_, h_ = rho.astype(dt), mom.astype(dt), en.astype(dt), h.astype(dt)
        inv = dt(1)/rho_
        sq = mom_*mom_
        int_ener = en_ - dt(.5)*sq*inv
        speed = np.sqrt(dt(1.4*0.4)*en_*inv) + np.sqrt(sq)*inv
        temp = int_ener*inv*dt(gm1_over_r)
        sqrt_temp = np.sqrt(temp)
        visc = transport(1.716e-5, 273., 111., sqrt_temp, temp, dt)
        cond = transport(.0241, 273., 194., sqrt_temp, temp, dt)*dt(gm1_over_r)
        diffusivity = av[1].astype(dt) + np.maximum(av[0].astype(dt) + visc*inv, cond*inv)
        inv_h = dt(1)/h_
        return dt(1)/(speed*(dt(inv_c)*inv_h) + diffusivity*(dt(inv_d)*inv_h*inv_h)), int_ener > dt(0.03)*en_
    exact, _ = local_dt(np.float64)
    approx, trusted = local_dt(f32)
    assert trusted.sum() > n//2 and (~trusted).sum() > n//10   # both branches are exercised
    err = np.abs(approx[trusted].astype(np.float64)/exact[trusted] - 1)
    assert err.max() < 2.5e-5, err.max()                        # (1 + err)^2 stays well inside the 1e-4 margin
