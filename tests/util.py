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


def run_euler_pair(oracle, lib_path, mesh, basis, n_steps=2, local_time=False, use_filter=False, safety=0.7, options=()):
    """advance `mesh` n_steps (each: max_dt + 2 stages with ghost-state BCs) on both implementations.
    Returns (device result mesh, oracle result mesh, list of (dt_device, dt_oracle))."""
    ref = mesh.copy()
    dev = Device(mesh.n_dim, mesh.row_size, basis, lib_path=lib_path).load_mesh(mesh)
    for opt, val in options:
        dev.set_option(opt, val)
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


# ---------------------------------------------------------------------------------------------------------------
# the other PDEs (reference include/pde.hpp): Navier-Stokes, advection, smooth artificial viscosity, fix therm admis
# ---------------------------------------------------------------------------------------------------------------
from pyoracle import NAVIER_STOKES, ADVECTION, SMOOTH_AV, FIX_THERM_ADMIS  # noqa: E402
import pyoracle  # noqa: E402
from hexed_b200 import kernels as K  # noqa: E402


def prepare_pde_state(mesh, rng, pde):
    """random but benign data in the element slots the PDE touches (slot numbers: reference include/pde.hpp:17-21)"""
    nd, rs = mesh.n_dim, mesh.row_size
    d = mesh.elem_data
    ne, nq = mesh.n_elem, mesh.nq
    if pde == NAVIER_STOKES:
        d[:, M.BULK_AV_SLOT(nd)] = 1e-4*rng.random((ne, nq))
        d[:, M.LAPLACIAN_AV_SLOT(nd)] = 1e-4*rng.random((ne, nq))
    elif pde == ADVECTION:
        d[:, :nd] = rng.standard_normal((ne, nd, nq))                 # advection velocity lives in the momentum slots
        d[:, M.ADVECTION_SLOT(nd):M.ADVECTION_SLOT(nd) + rs] = 1. + 0.1*rng.standard_normal((ne, rs, nq))
        mesh.face_wide[:] = rng.standard_normal(mesh.face_wide.shape)
    elif pde == SMOOTH_AV:
        d[:, M.FORCING_SLOT(nd):M.FORCING_SLOT(nd) + 4] = rng.standard_normal((ne, 4, nq))


def run_pde_pair(oracle, lib_path, mesh, basis, pde, n_steps=1, local_time=False, use_filter=False, safety=0.3, compute_residual=False, options=()):
    """one `max_dt_*` + stage sequence per step on both implementations, mirroring how Solver drives each PDE"""
    ref = mesh.copy()
    dev = Device(mesh.n_dim, mesh.row_size, basis, lib_path=lib_path).load_mesh(mesh)
    for opt, val in options:
        dev.set_option(opt, val)
    visc_o, cond_o = pyoracle.sutherland(1.7e-5, 273., 110.), pyoracle.constant(2.5e-2)
    visc_d, cond_d = K.sutherland(1.7e-5, 273., 110.), K.constant_transport(2.5e-2)
    dts = []
    for _ in range(n_steps):
        if pde == NAVIER_STOKES:
            dt_o = oracle.max_dt(pde, basis, ref, safety, safety, local_time, visc_o, cond_o)
            dt_d = dev.max_dt_navier_stokes(safety, safety, local_time, visc_d, cond_d)
            oracle.apply_state_bcs(ref); dev.apply_state_bcs()
            oracle.compute_navier_stokes(basis, ref, lambda: oracle.apply_flux_bcs(ref), visc_o, cond_o, dt=dt_o, i_stage=0,
                                         use_filter=use_filter, compute_residual=compute_residual)
            dev.compute_navier_stokes(dev.apply_flux_bcs, visc_d, cond_d, dt=dt_o, i_stage=0, use_filter=use_filter, compute_residual=compute_residual)
            if not compute_residual:
                oracle.apply_state_bcs(ref); dev.apply_state_bcs()
                oracle.compute_euler(basis, ref, dt=dt_o, i_stage=1, use_filter=use_filter)
                dev.compute_euler(dt=dt_o, i_stage=1, use_filter=use_filter)
        elif pde == ADVECTION:
            dt_o = oracle.max_dt(pde, basis, ref, safety, safety, local_time, advect_length=0.7)
            dt_d = dev.max_dt_advection(safety, safety, local_time, 0.7)
            for stage in (0, 1):
                oracle.compute_advection(basis, ref, 0.7, dt=dt_o, i_stage=stage, use_filter=use_filter)
                dev.compute_advection(0.7, dt=dt_o, i_stage=stage, use_filter=use_filter)
        elif pde == SMOOTH_AV:
            dt_o = oracle.max_dt(pde, basis, ref, safety, safety, local_time)
            dt_d = dev.max_dt_smooth_av(safety, safety, local_time)
            oracle.compute_smooth_av(basis, ref, None, 0.4, 1.3, dt=dt_o, i_stage=0, use_filter=use_filter)
            dev.compute_smooth_av(None, 0.4, 1.3, dt=dt_o, i_stage=0, use_filter=use_filter)
        else:
            dt_o = oracle.max_dt(pde, basis, ref, safety, safety, local_time)
            dt_d = dev.max_dt_fix_therm_admis(safety, safety, local_time)
            oracle.compute_fix_therm_admis(basis, ref, None, dt=dt_o, i_stage=0, use_filter=use_filter, compute_residual=compute_residual)
            dev.compute_fix_therm_admis(None, dt=dt_o, i_stage=0, use_filter=use_filter, compute_residual=compute_residual)
        dts.append((dt_d, dt_o))
    out = mesh.copy()
    dev.sync_to_host(out)
    dev.close()
    return out, ref, dts


def assert_pde_parity(out, ref, dts, tol=STATE_TOL):
    for dt_d, dt_o in dts:
        assert abs(dt_d - dt_o) <= MAX_DT_TOL*abs(dt_o), (dt_d, dt_o)
    nd, rs = out.n_dim, out.row_size
    c = M.cache_slot(nd, rs)
    # one comparison per slot group so that a small-magnitude group cannot hide behind the flow state
    groups = {"state": (0, nd + 2), "av": (nd + 3, nd + 5), "forcing": (nd + 5, nd + 9), "advection": (nd + 9, nd + 9 + rs)}
    for name, (lo, hi) in groups.items():
        a, b = out.elem_data[:, lo:hi], ref.elem_data[:, lo:hi]
        assert np.isfinite(b).all(), name
        if np.linalg.norm(b) == 0:
            assert np.array_equal(a, b), name
        else:
            assert rel_l2(a, b) <= tol, name
    a, b = out.elem_data[:, c:], ref.elem_data[:, c:]
    assert (np.array_equal(a, b) if np.linalg.norm(b) == 0 else rel_l2(a, b) <= 10*tol), "residual cache"  # cancellation-prone difference
    assert rel_l2(out.tss(), ref.tss()) <= MAX_DT_TOL
    for name in ("face_state", "face_ldg", "face_wide"):
        a, b = getattr(out, name), getattr(ref, name)
        if a is not None:
            assert rel_l2(a, b) <= tol, name


def check_cfl_cache(oracle, lib, rs, n):
    """the stage-1 Local kernel leaves per-element CFL ratios behind for the next max_dt_euler (HEXED_B200_OPT_CFL_CACHE); any
    other write to the state must invalidate them, and the cached and the full reduction must agree to round-off"""
    basis = hb.gauss_legendre(rs)
    m = M.box_mesh(3, rs, n, basis, deformed=True, bc_kind=M.BC_FREESTREAM, bc_params=freestream_state(3))
    density_wave(m, basis)
    oracle.compute_write_face(basis, m)
    ref = m.copy()
    dev = Device(3, rs, basis, lib_path=lib).load_mesh(m)
    dev.set_option(1, 1)                                     # HEXED_B200_OPT_CFL_CACHE (off by default)
    dt = dev.max_dt_euler(0.7, 0.7, False)
    for stage in (0, 1):
        oracle.apply_state_bcs(ref); oracle.compute_euler(basis, ref, dt=dt, i_stage=stage)
        dev.apply_state_bcs(); dev.compute_euler(dt=dt, i_stage=stage)
    n0 = dev.launch_count()
    dt_cached = dev.max_dt_euler(0.7, 0.7, False)           # from the single-precision screen the Local kernel left behind
    assert dev.launch_count() - n0 == 2                      # screen minimum + exact re-evaluation of the near-minimum elements
    dt_o = oracle.max_dt(EULER, basis, ref, 0.7, 0.7, False)
    assert abs(dt_cached/dt_o - 1) <= 1e-13
    dev.set_option(1, 0)                                     # HEXED_B200_OPT_CFL_CACHE off: the full kernel
    assert abs(dev.max_dt_euler(0.7, 0.7, False)/dt_o - 1) <= 1e-13
    dev.set_option(1, 1)
    # a host write to the state: the cache must not be used
    st = ref.state().copy(); st[:, 3 + 1] *= 4.                # four times the energy -> twice the sound speed
    ref.state()[:] = st
    dev.upload_elements(np.ascontiguousarray(st), 0, 5)
    dt_new = dev.max_dt_euler(0.7, 0.7, False)
    dt_o = oracle.max_dt(EULER, basis, ref, 0.7, 0.7, False)
    assert abs(dt_new/dt_o - 1) <= 1e-13 and dt_new < 0.8*dt_cached
    # local time stepping never uses it and leaves tss != 1; the next global call must restore tss = 1
    dev.max_dt_euler(0.7, 0.7, True)
    for stage in (0, 1):
        dev.apply_state_bcs(); dev.compute_euler(dt=1e-9, i_stage=stage)
    dev.max_dt_euler(0.7, 0.7, False)
    out = m.copy(); dev.sync_to_host(out)
    assert np.array_equal(out.tss(), np.ones_like(out.tss()))
    dev.close()


def check_tss_write_skipped(oracle, lib, nd=2, rs=4, pde="euler"):
    """global time stepping writes time_step_scale = 1 (reference Spatial.hpp:823-825); the device skips those 8 bytes per point when the
    array is known to hold 1 already. Every way the array can change (local time stepping, a host upload of its slot) must bring the
    write back."""
    basis = hb.gauss_legendre(rs)
    m = M.box_mesh(nd, rs, 5, basis, deformed=True, bc_kind=M.BC_FREESTREAM, bc_params=freestream_state(nd))
    density_wave(m, basis)
    oracle.compute_write_face(basis, m)
    dev = Device(nd, rs, basis, lib_path=lib).load_mesh(m)
    nv = nd + 2
    if pde == "euler":
        max_dt = lambda local: dev.max_dt_euler(0.7, 0.7, local)
        want = oracle.max_dt(EULER, basis, m, 0.7, 0.7, False)
    else:
        max_dt = lambda local: dev.max_dt_navier_stokes(0.7, 0.7, local, K.constant_transport(1e-3), K.constant_transport(2e-3))
        want = oracle.max_dt(NAVIER_STOKES, basis, m, 0.7, 0.7, False, pyoracle.constant(1e-3), pyoracle.constant(2e-3))
    ones = np.ones_like(m.tss())
    def tss():
        out = m.copy(); dev.sync_to_host(out); return out.tss()
    for rep in range(3):                                     # first call writes, the others skip
        assert abs(max_dt(False)/want - 1) <= MAX_DT_TOL
        assert np.array_equal(tss(), ones)
    junk = np.full((m.n_elem, 1) + m.tss().shape[1:], 7.)
    dev.upload_elements(np.ascontiguousarray(junk), nv, 1)   # a host write to the time-step-scale slot
    assert np.array_equal(tss(), junk[:, 0])
    assert abs(max_dt(False)/want - 1) <= MAX_DT_TOL
    assert np.array_equal(tss(), ones)
    max_dt(True)                                             # local time stepping
    assert not np.array_equal(tss(), ones)
    assert abs(max_dt(False)/want - 1) <= MAX_DT_TOL
    assert np.array_equal(tss(), ones)
    dev.close()


def check_max_dt_running_screen(oracle, lib, nd, rs, n):
    """the global-time-step reduction evaluates only the points that pass a running single-precision screen with the FP64 formula;
    the result must be THE double the unscreened arithmetic gives (the local-time-step kernel writes it per point), whatever the flow:
    a smooth wave, a uniform flow (every point ties: all evaluated), a fluid at rest, a few extreme cells"""
    basis = hb.gauss_legendre(rs)
    m = M.box_mesh(nd, rs, n, basis, deformed=True, bc_kind=M.BC_FREESTREAM, bc_params=freestream_state(nd))
    density_wave(m, basis)
    dev = Device(nd, rs, basis, lib_path=lib).load_mesh(m)
    nv = nd + 2
    rng = np.random.default_rng(nd*100 + rs)
    wave = m.state().copy()
    uniform = np.broadcast_to(freestream_state(nd)[None, :, None], wave.shape).copy()
    rest = uniform.copy(); rest[:, :nd] = 0.
    spiky = wave.copy()
    for e in rng.integers(0, m.n_elem, 5):
        spiky[e, nd + 1, rng.integers(0, m.nq)] *= rng.uniform(2., 50.)   # hot spots: the minimum sits in one of them
    huge = wave.copy(); huge[:, nd + 1] *= 1e36; huge[:, :nd] *= 1e18           # energy*density overflows a float: exact path everywhere
    for name, st in (("wave", wave), ("uniform", uniform), ("rest", rest), ("spiky", spiky), ("huge", huge)):
        m.state()[:] = st
        dev.upload_elements(np.ascontiguousarray(st), 0, nv)
        dt = dev.max_dt_euler(0.7, 0.7, False)
        dev.max_dt_euler(0.7, 0.7, True)
        tss = np.empty((m.n_elem, 1, m.nq)); dev.download_elements(tss, nv, 1)
        assert dt == tss.min(), (name, dt, tss.min())
        assert abs(dt/oracle.max_dt(EULER, basis, m, 0.7, 0.7, False) - 1) <= MAX_DT_TOL, name
    dev.close()


def check_max_dt_running_screen_random(oracle, lib, nd, rs, n, seeds):
    """the same bit-identity on random admissible states with a wide dynamic range (densities over 3 decades, Mach 0 to ~30, cells at
    rest, ties between elements): whatever passes the single-precision screen, the minimum is the unscreened double"""
    basis = hb.gauss_legendre(rs)
    m = M.box_mesh(nd, rs, n, basis, deformed=True, bc_kind=M.BC_FREESTREAM, bc_params=freestream_state(nd))
    dev = Device(nd, rs, basis, lib_path=lib).load_mesh(m)
    nv = nd + 2
    shape = m.state().shape
    for seed in seeds:
        rng = np.random.default_rng(seed)
        st = np.empty(shape)
        rho = 10.**rng.uniform(-2, 1, (shape[0], shape[2]))
        sound = 10.**rng.uniform(1.5, 3., (shape[0], shape[2]))
        mach = rng.uniform(0., 30., (shape[0], 1))*rng.integers(0, 2, (shape[0], 1))      # half of the elements at rest
        direction = rng.normal(size=(shape[0], nd, shape[2])); direction /= np.linalg.norm(direction, axis=1, keepdims=True)
        st[:, :nd] = (rho*sound*mach)[:, None, :]*direction
        st[:, nd] = rho
        st[:, nd + 1] = rho*sound**2/(1.4*0.4) + 0.5*np.sum(st[:, :nd]**2, axis=1)/rho
        if seed % 3 == 0:                                                                   # exact ties between two elements
            st[1] = st[0]
        m.state()[:] = st
        dev.upload_elements(np.ascontiguousarray(st), 0, nv)
        dt = dev.max_dt_euler(0.7, 0.7, False)
        dev.max_dt_euler(0.7, 0.7, True)
        tss = np.empty((m.n_elem, 1, m.nq)); dev.download_elements(tss, nv, 1)
        assert dt == tss.min(), (seed, dt, tss.min())
        assert abs(dt/oracle.max_dt(EULER, basis, m, 0.7, 0.7, False) - 1) <= MAX_DT_TOL, seed
    dev.close()


def check_max_dt_running_screen_ns(oracle, lib, nd, rs, n, seeds):
    """the Navier-Stokes global time step (g_max_dt_screen_kernel) against the unscreened per-point values of local time stepping:
    bit-identical on random admissible states (densities over 3 decades, Mach 0 to ~30 so that the forced-exact branch above Mach ~7
    is taken, artificial-viscosity coefficients that make the diffusive term dominate in places, tiny cells), and against the oracle"""
    basis = hb.gauss_legendre(rs)
    m = M.box_mesh(nd, rs, n, basis, deformed=True, bc_kind=M.BC_FREESTREAM, bc_params=freestream_state(nd))
    dev = Device(nd, rs, basis, lib_path=lib).load_mesh(m)
    nv = nd + 2
    shape = m.state().shape
    ne, nq = shape[0], shape[2]
    models = ((K.sutherland(1.716e-5, 273., 111.), K.sutherland(.0241, 273., 194.), pyoracle.sutherland(1.716e-5, 273., 111.), pyoracle.sutherland(.0241, 273., 194.)),
              (K.constant_transport(3e-2), K.constant_transport(50.), pyoracle.constant(3e-2), pyoracle.constant(50.)))
    for seed in seeds:
        rng = np.random.default_rng(1000 + seed)
        visc_d, cond_d, visc_o, cond_o = models[seed % 2]
        st = np.empty(shape)
        rho = 10.**rng.uniform(-2, 1, (ne, nq))
        sound = 10.**rng.uniform(1.5, 3., (ne, nq))
        mach = rng.uniform(0., 30., (ne, 1))*rng.integers(0, 2, (ne, 1))
        direction = rng.normal(size=(ne, nd, nq)); direction /= np.linalg.norm(direction, axis=1, keepdims=True)
        st[:, :nd] = (rho*sound*mach)[:, None, :]*direction
        st[:, nd] = rho
        st[:, nd + 1] = rho*sound**2/(1.4*0.4) + 0.5*np.sum(st[:, :nd]**2, axis=1)/rho
        if seed % 3 == 0:
            st[1] = st[0]
        av = 10.**rng.uniform(-6, 1 if seed % 4 == 0 else -3, (ne, 2, nq))*rng.choice([-1., 1.], (ne, 2, nq))
        m.state()[:] = st
        m.elem_data[:, M.BULK_AV_SLOT(nd)] = av[:, 0]
        m.elem_data[:, M.LAPLACIAN_AV_SLOT(nd)] = av[:, 1]
        dev.upload_elements(np.ascontiguousarray(st), 0, nv)
        dev.upload_elements(np.ascontiguousarray(av), M.BULK_AV_SLOT(nd), 2)
        dt = dev.max_dt_navier_stokes(0.7, 0.6, False, visc_d, cond_d)
        dev.max_dt_navier_stokes(0.7, 0.6, True, visc_d, cond_d)
        tss = np.empty((ne, 1, nq)); dev.download_elements(tss, nv, 1)
        assert dt == tss.min(), (seed, dt, tss.min())
        assert abs(dt/oracle.max_dt(NAVIER_STOKES, basis, m, 0.7, 0.6, False, visc_o, cond_o) - 1) <= MAX_DT_TOL, seed
    dev.close()


def mixed_bcs(mesh, rng):
    """replace the soup mesh's single boundary condition by one of every device-side kind over disjoint subsets of its faces"""
    src = mesh.bcs[0]
    n = src["ghost_slot"].size
    nd = mesh.n_dim
    kinds = [(M.BC_COPY, None), (M.BC_NONPENETRATION, None), (M.BC_FREESTREAM, freestream_state(nd)), (M.BC_OUTFLOW, None),
             (M.BC_PRESSURE_OUTFLOW, np.array([0.9e5])), (M.BC_NO_SLIP, M.no_slip_params(M.THERMAL_ENERGY, 2.1e5)),
             (M.BC_NO_SLIP, M.no_slip_params(M.THERMAL_HEAT_FLUX, 30.)), (M.BC_NO_SLIP, M.no_slip_params(M.THERMAL_EQUILIBRIUM, .8, 12., 280.)),
             (M.BC_RIEMANN_INVARIANTS, freestream_state(nd))]
    owner = np.arange(n) % len(kinds)
    bcs = []
    for k, (kind, params) in enumerate(kinds):
        sel = np.nonzero(owner == k)[0]
        bcs.append(dict(kind=kind, params=params, **{key: np.ascontiguousarray(src[key][sel]) for key in ("inside_slot", "ghost_slot", "normal_slot", "con_index")}))
    mesh.bcs = bcs
    return mesh


def check_riemann_bc(oracle, lib, nd, rs, seed=11):
    """Riemann_invariants::apply_state / apply_flux (src/Boundary_condition.cpp:97-182) on every boundary face of a soup mesh with
    inside states from Mach 0 to 3 in both directions through the face (sub/supersonic in/outflow), arbitrary normals, and a few
    zero-pressure points where the eigenvector matrix is singular and the pivoted QR gives the basic least-squares solution"""
    rng = np.random.default_rng(seed)
    basis = hb.gauss_legendre(rs)
    m = M.soup_mesh(nd, rs, rng, n_car=12, n_def=30, n_ref=2, with_ldg=True)
    M.random_flow_state(m, rng)
    src = m.bcs[0]
    fs = freestream_state(nd)
    m.bcs = [dict(src, kind=M.BC_RIEMANN_INVARIANTS, params=fs)]
    ins = src["inside_slot"]
    n, nfq, nv = ins.size, m.nfq, nd + 2
    mass = rng.uniform(.4, 2., (n, nfq))
    pres = rng.uniform(2e4, 2e5, (n, nfq))
    sound = np.sqrt(1.4*pres/mass)
    veloc = rng.normal(0., 1., (n, nd, nfq))
    veloc *= (rng.uniform(0., 3., (n, 1, nfq))*sound[:, None])/np.linalg.norm(veloc, axis=1, keepdims=True)
    f = np.zeros((n, nv, nfq))
    f[:, :nd] = mass[:, None]*veloc
    f[:, nd] = mass
    f[:, nd + 1] = pres/.4 + .5*mass*(veloc**2).sum(1)
    f[0, nd + 1, 0] = .5*mass[0, 0]*(veloc[0, :, 0]**2).sum()      # zero pressure
    f[1 % n, nd + 1, 0] = .4*mass[1 % n, 0]*(veloc[1 % n, :, 0]**2).sum()  # negative pressure
    m.face_state[ins] = f.reshape(n, -1)
    # viscous-flux-like magnitudes per variable (momentum : mass : energy flux ~ 1 : 1/|u| : |u|): an unscaled random vector would be
    # dominated by its component along the energy eigenvectors and the decomposition would cancel 1e5-fold (6e-9 between an FMA and a
    # non-FMA CPU build of the oracle itself)
    flux_scale = np.array([50.]*nd + [50./300., 50.*300.])
    m.face_ldg[ins] = (rng.normal(0., 1., (n, nv, nfq))*flux_scale[None, :, None]).reshape(n, -1)
    ref = m.copy()
    dev = Device(nd, rs, basis, lib_path=lib).load_mesh(m)
    oracle.apply_state_bcs(ref); dev.apply_state_bcs()
    oracle.apply_flux_bcs(ref); dev.apply_flux_bcs()
    out = m.copy()
    dev.sync_to_host(out)
    dev.close()
    gh = src["ghost_slot"]
    assert np.isfinite(ref.face_state[gh]).all() and np.isfinite(ref.face_ldg[gh]).all()
    # the decomposition cancels O(|eigenvector entries|) terms, so compare point by point against the state magnitude
    scale = np.abs(ref.face_state[ins]).reshape(n, nv, nfq).max(1, keepdims=True) + np.abs(fs).max()
    err = np.abs(out.face_state[gh] - ref.face_state[gh]).reshape(n, nv, nfq)/scale
    fscale = flux_scale[None, :, None]
    ferr = np.abs(out.face_ldg[gh] - ref.face_ldg[gh]).reshape(n, nv, nfq)/fscale
    # at the two non-positive-pressure points the eigenvector matrix is exactly rank deficient: whether the last pivot (pure
    # round-off, ~eps*norm) falls under ColPivHouseholderQR's threshold (eps*norm)^2*(rows - k)/rows depends on the last bit, i.e. on
    # FMA contraction -- the reference's own answer there changes with its compiler flags. Those points must stay finite; every
    # other point is held to the tolerance.
    regular = np.ones((n, 1, nfq), bool)
    regular[0, 0, 0] = regular[1 % n, 0, 0] = False
    assert np.isfinite(out.face_state[gh]).all() and np.isfinite(out.face_ldg[gh]).all()
    # the eigenvector matrix has condition number 1e5..1e6 here (Mach up to 3), so a backward-stable solve is good to cond*eps ~ 1e-10
    # at the worst point: hold the north-star's relative L2 <= 1e-11 over the regular points and cap the pointwise error at 1e-9
    for name, e_, d_, r_ in (("state", err, out.face_state[gh] - ref.face_state[gh], ref.face_state[gh]),
                             ("flux", ferr, out.face_ldg[gh] - ref.face_ldg[gh], ref.face_ldg[gh])):
        worst = np.unravel_index(np.argmax(e_*regular), e_.shape)
        assert (e_*regular).max() <= 1e-9, (name, e_[worst], worst)
        mask = np.broadcast_to(regular, e_.shape)
        assert np.linalg.norm(d_.reshape(e_.shape)[mask]) <= STATE_TOL*np.linalg.norm(r_.reshape(e_.shape)[mask]), name
    # the mix of regimes is real: some points fully inside, some fully freestream, some in between
    g = ref.face_state[gh].reshape(n, nv, nfq)
    same_in = np.isclose(g, f, rtol=1e-9).all(1)
    same_fs = np.isclose(g, fs[None, :, None], rtol=1e-9).all(1)
    assert same_in.any() and same_fs.any() and (~same_in & ~same_fs).any()


def check_admissibility(oracle, lib, nd, rs, seed=3):
    """Solver::is_admissible on the device against the oracle's restatement: admissible state; non-positive mass inside an element;
    non-positive energy on an element face only; on a fine mortar face only (flag without a record); non-finite value -> error"""
    import pytest
    rng = np.random.default_rng(seed)
    basis = hb.gauss_legendre(rs)
    m = M.soup_mesh(nd, rs, rng, n_car=9, n_def=14, n_ref=3 if nd > 1 else 0)
    M.random_flow_state(m, rng)
    oracle.compute_write_face(basis, m)
    oracle.compute_prolong(basis, m)
    nv, nq, nfq = nd + 2, m.nq, m.nfq

    def both(mesh):
        dev = Device(nd, rs, basis, lib_path=lib).load_mesh(mesh)
        try:
            got = dev.is_admissible()
            rec = dev.record()
        finally:
            dev.close()
        want, want_rec = oracle.is_admissible(mesh)
        assert got == want and np.array_equal(rec, want_rec)
        return got, rec

    ok, rec = both(m)
    assert ok and not rec.any()
    a = m.copy()
    a.state()[5, nd, nq//2] = -1e-3                       # mass <= 0 at one interior point
    a.state()[7, nd + 1, 0] = 0.                          # energy == 0 counts as inadmissible (strict >)
    ok, rec = both(a)
    assert not ok and rec.sum() == 2 and rec[5] and rec[7]
    b = m.copy()
    b.face_state[2*nd*11 + 1].reshape(nv, nfq)[nd + 1, nfq - 1] = -5.   # only on face 1 of element 11
    ok, rec = both(b)
    assert not ok and rec.sum() == 1 and rec[11]
    if m.ref_face.shape[0]:
        c = m.copy()
        c.face_state[m.ref_face[1, 1]].reshape(nv, nfq)[nd, 0] = -1.   # first fine mortar face of refined face 1
        ok, rec = both(c)
        assert not ok and not rec.any()
    d = m.copy()
    d.state()[3, 0, 1] = np.nan                           # momentum: not part of the sign test, but must be finite
    dev = Device(nd, rs, basis, lib_path=lib).load_mesh(d)
    with pytest.raises(RuntimeError, match="state is not finite"):
        dev.is_admissible()
    dev.close()
    with pytest.raises(RuntimeError, match="state is not finite"):
        oracle.is_admissible(d)


def check_set_jacobian(lib, nd, rs, seed=21):
    """hexed_b200_set_jacobian against the loop-by-loop numpy restatement of Deformed_element::set_jacobian (pinned by the reference's
    test_Deformed_element values): perturbed vertices, random face-warping node adjustments, Cartesian elements' unit face Jacobian"""
    import pyoracle
    from hexed_b200.kernels import REF_NORMALS, JAC_DET, VERTEX_TSS, FACE_STATE
    rng = np.random.default_rng(seed)
    basis = hb.gauss_legendre(rs)
    n_car, n_def = 3, 7
    m = M.FlatMesh(nd, rs, n_car, n_def)
    nfq, n_vert = m.nfq, 2**nd
    m.nom_size = rng.choice([.125, .25, .5], n_car + n_def)
    corners = np.array([[(i >> (nd - 1 - d)) & 1 for d in range(nd)] for i in range(n_vert)], dtype=float)
    vert = np.stack([(corners + rng.integers(0, 4, nd) + rng.uniform(-.12, .12, (n_vert, nd)))*m.nom_size[n_car + e] for e in range(n_def)])
    adj = rng.uniform(-.05, .05, (n_def, 2*nd, nfq))
    adj[0] = 0.
    m.face_state[:] = 7.  # everything but the first nd*nfq doubles of each element face must survive
    dev = Device(nd, rs, basis, lib_path=lib).load_mesh(m)
    dev.set_jacobian(vert, adj)
    refn = dev.download(REF_NORMALS, np.zeros_like(m.ref_normals))
    det = dev.download(JAC_DET, np.zeros_like(m.det))
    vtss = dev.download(VERTEX_TSS, np.zeros_like(m.vertex_tss))
    faces = dev.download(FACE_STATE, np.zeros_like(m.face_state))
    dev.close()
    for e in range(n_def):
        g = pyoracle.set_jacobian(vert[e], adj[e].reshape(-1), m.nom_size[n_car + e], basis)
        scale = np.abs(g["ref_normals"]).max()
        assert np.abs(refn[e] - g["ref_normals"]).max() <= 1e-13*scale
        assert np.abs(det[e] - g["det"]).max() <= 1e-13*np.abs(g["det"]).max()
        # extrapolation to faces / vertices sums row_size^(1..3) terms of alternating sign (boundary coefficients up to ~5 at row size 8)
        assert np.abs(vtss[n_car + e] - g["vertex_tss"]).max() <= 2e-12*np.abs(g["vertex_tss"]).max()
        f = faces[(n_car + e)*2*nd:(n_car + e + 1)*2*nd].reshape(2*nd, nd + 2, nfq)
        assert np.abs(f[:, :nd] - g["face_normals"]).max() <= 2e-12*scale
        assert np.all(f[:, nd:] == 7.)
    fc = faces[:n_car*2*nd].reshape(n_car, 2*nd, nd + 2, nfq)
    for f in range(2*nd):
        for j in range(nd):
            assert np.all(fc[:, f, j] == (1. if f//2 == j else 0.))
    assert np.all(fc[:, :, nd:] == 7.) and np.all(vtss[:n_car] == 1.)


def check_av_glue(lib, nd, rs, seed=8):
    """the pointwise loops of Solver::update_art_visc_smoothness / fix_admissibility on the device against their
    numpy restatements, chained in the order the Solver runs them"""
    import pyoracle
    rng = np.random.default_rng(seed)
    basis = hb.gauss_legendre(rs)
    m = M.soup_mesh(nd, rs, rng, n_car=6, n_def=9, n_ref=0)
    M.random_flow_state(m, rng)
    m.elem_data[:, nd + 3:nd + 5] = rng.uniform(0., 2e-3, (m.n_elem, 2, m.nq))          # AV coefficients
    m.elem_data[:, nd + 5:nd + 9] = rng.uniform(0., 1e-3, (m.n_elem, 4, m.nq))          # forcing
    m.elem_data[:, nd + 9:nd + 9 + rs] = rng.normal(1., .3, (m.n_elem, rs, m.nq))       # advection state
    m.nom_size = rng.choice([.25, .5, 1.], m.n_elem)
    w = np.asarray(basis.weight); orth = np.asarray(basis.orthogonal).reshape(rs, rs)[rs - 1]  # Basis::orthogonal(row_size - 1)
    vert = rng.uniform(0., 1., (m.n_elem, 2**nd))
    interp = np.stack([1. - np.asarray(basis.node), np.asarray(basis.node)], axis=1)
    ref = m.copy()
    dev = Device(nd, rs, basis, lib_path=lib).load_mesh(m)
    # update_art_visc_smoothness: normalise, (advection kernels), project, (diffusion kernels), finish
    dev.av_scale_velocity(); pyoracle.av_scale_velocity(ref)
    dev.av_project_forcing(w, orth); pyoracle.av_project_forcing(ref, w, orth)
    got = dev.av_finish(0.7, 3e-3, 3, w); want = pyoracle.av_finish(ref, 0.7, 3e-3, 3, w)
    assert abs(got - want) <= 1e-12*abs(want), (got, want)
    # fix_admissibility: vertex factors -> laplacian AV coefficient, swapped with the bulk coefficient
    dev.interp_vertices(1, vert, interp); pyoracle.interp_vertices(ref, 1, vert, interp)
    dev.av_swap(); pyoracle.av_swap(ref)
    out = m.copy()
    dev.sync_to_host(out)
    dev.close()
    for name, lo, hi in (("state", 0, nd + 2), ("av", nd + 3, nd + 5), ("forcing", nd + 5, nd + 9)):
        assert rel_l2(out.elem_data[:, lo:hi], ref.elem_data[:, lo:hi]) <= 1e-14, name
    assert rel_l2(out.state(), m.state()) <= 1e-14  # the velocity normalisation is undone exactly to rounding


def check_aux_bcs(lib, nd, rs, seed=13):
    """the boundary loops of the AV / admissibility pipelines (advection ghosts on the wide faces, diffusion state copy, negated
    LDG flux) for every registered boundary-condition kind"""
    import pyoracle
    from hexed_b200.kernels import BC_MODE_ADVECTION, BC_MODE_COPY_STATE, BC_MODE_NEGATE_FLUX
    rng = np.random.default_rng(seed)
    basis = hb.gauss_legendre(rs)
    m = M.soup_mesh(nd, rs, rng, n_car=10, n_def=24, n_ref=2, with_ldg=True, with_wide=True)
    M.random_flow_state(m, rng)
    mixed_bcs(m, rng)
    m.face_state[:] = rng.normal(0., 1., m.face_state.shape)
    m.face_ldg[:] = rng.normal(0., 1., m.face_ldg.shape)
    m.face_wide[:] = rng.normal(0., 1., m.face_wide.shape)
    ref = m.copy()
    dev = Device(nd, rs, basis, lib_path=lib).load_mesh(m)
    for mode in (BC_MODE_ADVECTION, BC_MODE_COPY_STATE, BC_MODE_NEGATE_FLUX):
        dev.apply_aux_bcs(mode)
        pyoracle.apply_aux_bcs(ref, mode)
    out = m.copy()
    dev.sync_to_host(out)
    dev.close()
    assert np.array_equal(out.face_state, ref.face_state) and np.array_equal(out.face_ldg, ref.face_ldg)
    assert rel_l2(out.face_wide, ref.face_wide) <= 1e-15
    assert not np.array_equal(out.face_wide, m.face_wide)


def check_av_pipeline(oracle, lib, nd, rs, seed=17, advect_iters=2, diff_iters=1, n_cheby=2):
    """Solver::update_art_visc_smoothness (reference src/Solver.cpp:457-581, with diffuse_art_visc :428-455) end to end with nothing on
    the host: velocity normalisation, advection pseudo-iterations with their ghost fill, Legendre projection into the forcing, smoothing
    iterations with their state / flux ghosts, AV coefficient + residual + velocity restore, final write_face / prolong. Device calls
    against the oracle's kernels and the numpy restatements of the glue, same order, same parameters."""
    import pyoracle
    from pyoracle import ADVECTION, SMOOTH_AV
    from hexed_b200.kernels import BC_MODE_ADVECTION, BC_MODE_COPY_STATE, BC_MODE_NEGATE_FLUX
    rng = np.random.default_rng(seed)
    basis = hb.gauss_legendre(rs)
    m = M.soup_mesh(nd, rs, rng, n_car=8, n_def=16, n_ref=2, with_ldg=True, with_wide=True)
    M.random_flow_state(m, rng)
    mixed_bcs(m, rng)
    m.elem_data[:, nd + 9:nd + 9 + rs] = 1.   # advection state starts at 1 (the ghosts' "2 - inside" keeps it there)
    m.elem_data[:, nd + 3] = rng.uniform(0., 1e-3, (m.n_elem, m.nq))
    advect_length, adv_safety, diff_safety = 0.3, 0.5, 0.4
    n_real = 3
    diff_time = 0.5*advect_length*advect_length/n_real
    mult, us_max = 0.8*advect_length, advect_length*0.2*np.sqrt(2*2.5e5/1.2)
    w = np.asarray(basis.weight); orth = np.asarray(basis.orthogonal).reshape(rs, rs)[rs - 1]

    def cheby(n, i):  # math::chebyshev_step, reference src/math.cpp:104-107
        return 1/(1 - np.cos((n - i - .5)*np.pi/n))
    ref = m.copy()
    dev = Device(nd, rs, basis, lib_path=lib).load_mesh(m)
    # -- oracle
    pyoracle.av_scale_velocity(ref)
    oracle.compute_write_face(basis, ref); oracle.compute_prolong(basis, ref)
    oracle.max_dt(ADVECTION, basis, ref, adv_safety, 1., True, advect_length=advect_length)
    for _ in range(advect_iters):
        oracle.compute_write_face(basis, ref, pde=ADVECTION); oracle.compute_prolong(basis, ref, pde=ADVECTION)
        for i in (0, 1):
            pyoracle.apply_aux_bcs(ref, BC_MODE_ADVECTION)
            oracle.compute_advection(basis, ref, advect_length, dt=1., i_stage=i)
    pyoracle.av_project_forcing(ref, w, orth)
    oracle.max_dt(SMOOTH_AV, basis, ref, 1., diff_safety, True)
    oracle.compute_write_face(basis, ref, pde=SMOOTH_AV); oracle.compute_prolong(basis, ref)
    for _ in range(diff_iters):
        for ic in range(n_cheby):
            sc = cheby(n_cheby, ic)
            pyoracle.apply_aux_bcs(ref, BC_MODE_COPY_STATE)
            oracle.compute_smooth_av(basis, ref, lambda: pyoracle.apply_aux_bcs(ref, BC_MODE_NEGATE_FLUX), diff_time, sc, dt=sc, i_stage=0)
    want = pyoracle.av_finish(ref, mult, us_max, n_real, w)
    oracle.compute_write_face(basis, ref); oracle.compute_prolong(basis, ref)
    # -- device
    dev.av_scale_velocity()
    dev.compute_write_face(); dev.compute_prolong()
    dev.max_dt_advection(adv_safety, 1., True, advect_length)
    for _ in range(advect_iters):
        dev.compute_write_face_advection(); dev.compute_prolong_advection()
        for i in (0, 1):
            dev.apply_aux_bcs(BC_MODE_ADVECTION)
            dev.compute_advection(advect_length, dt=1., i_stage=i)
    dev.av_project_forcing(w, orth)
    dev.max_dt_smooth_av(1., diff_safety, True)
    dev.compute_write_face_smooth_av(); dev.compute_prolong()
    for _ in range(diff_iters):
        for ic in range(n_cheby):
            sc = cheby(n_cheby, ic)
            dev.apply_aux_bcs(BC_MODE_COPY_STATE)
            dev.compute_smooth_av(lambda: dev.apply_aux_bcs(BC_MODE_NEGATE_FLUX), diff_time, sc, dt=sc, i_stage=0)
    got = dev.av_finish(mult, us_max, n_real, w)
    dev.compute_write_face(); dev.compute_prolong()
    out = m.copy()
    dev.sync_to_host(out)
    dev.close()
    assert np.isfinite(want) and want > 0 and abs(got - want) <= 1e-10*want, (got, want)
    assert_pde_parity(out, ref, [])
    assert rel_l2(out.state(), m.state()) <= 1e-13          # the flow state comes back as it was found
    assert not np.array_equal(out.elem_data[:, nd + 3], m.elem_data[:, nd + 3])


def check_vertex_sharing(oracle, lib, nd, rs, seed=23):
    """Solver::share_vertex_data (min on the vertex time-step scale as at the end of calc_jacobian, max on the scratch array) with
    hanging-vertex matchers of every stretch, and the spreading step of fix_admissibility built on it, against the numpy restatement
    (whose matcher is the golden-vector-pinned hexed_b200.tables.hanging_vertex_match)"""
    import pyoracle
    from hexed_b200.kernels import VERTEX_TSS, VERTEX_SCRATCH
    rng = np.random.default_rng(seed)
    basis = hb.gauss_legendre(rs)
    m = M.soup_mesh(nd, rs, rng, n_car=14, n_def=16, n_ref=0)
    M.random_flow_state(m, rng)
    oracle.compute_write_face(basis, m)
    ne, n_vert = m.n_elem, 2**nd
    n_vertex = ne*n_vert//3 + 5
    elem_vertex = rng.integers(0, n_vertex, (ne, n_vert)).astype(np.int32)
    for e in range(ne):  # an element's vertices are distinct mesh vertices
        elem_vertex[e] = rng.choice(n_vertex, n_vert, replace=False)
    matchers = []
    if nd > 1:
        free = list(rng.permutation(ne))
        for stretch in ([(0, 0), (1, 0), (0, 1)] if nd == 3 else [(0, 0)]):
            for i_dim in range(nd):
                n_fine = 2**(nd - 1)//((1 + stretch[0])*(1 + stretch[1]))
                fine = [int(free.pop()) for _ in range(n_fine)] + [-1]*(4 - n_fine)
                matchers.append([i_dim, int(rng.integers(0, 2)), stretch[0], stretch[1]] + fine)
    matchers = np.array(matchers, np.int32).reshape(-1, 8)
    m.vertex_tss = rng.uniform(.1, 1., (ne, n_vert))
    m.state()[3, nd, 0] = -1.; m.state()[9, nd + 1, 1] = -2.   # two inadmissible elements
    m.elem_data[:, nd + 3:nd + 5] = rng.uniform(0., 1e-3, (ne, 2, m.nq))
    ref = m.copy()
    dev = Device(nd, rs, basis, lib_path=lib).load_mesh(m)
    dev.vertex_topology(elem_vertex, n_vertex, matchers)
    dev.share_vertex_data(VERTEX_TSS, False)
    got_tss = dev.download(VERTEX_TSS, np.zeros_like(m.vertex_tss))
    want_tss = pyoracle.share_vertex_data(ref.vertex_tss.copy(), elem_vertex, n_vertex, matchers, nd, False)
    close = lambda a, b: np.allclose(a, b, rtol=4e-16, atol=0.)  # noqa: E731  (the matcher's midpoints: contraction order, 1 ulp)
    assert close(got_tss, want_tss)
    scratch = rng.uniform(0., 1., (ne, n_vert))
    dev.upload(VERTEX_SCRATCH, scratch)
    dev.share_vertex_data(VERTEX_SCRATCH, True)
    assert close(dev.download(VERTEX_SCRATCH, np.zeros_like(scratch)), pyoracle.share_vertex_data(scratch.copy(), elem_vertex, n_vertex, matchers, nd, True))
    ok = dev.is_admissible()
    want_ok, record = oracle.is_admissible(ref)
    assert not ok and not want_ok and record.sum() == 2
    interp = np.stack([1. - np.asarray(basis.node), np.asarray(basis.node)], axis=1)
    dev.fix_admis_spread(interp)
    v = pyoracle.fix_admis_spread(ref, record, elem_vertex, n_vertex, matchers, interp)
    assert close(dev.download(VERTEX_SCRATCH, np.zeros_like(scratch)), v)
    out = m.copy()
    dev.sync_to_host(out)
    dev.close()
    assert rel_l2(out.elem_data[:, nd + 3:nd + 5], ref.elem_data[:, nd + 3:nd + 5]) <= 1e-15
    assert out.elem_data[:, nd + 3].max() > 0.5   # the flagged elements and their neighbours carry a factor ~1 in what is now the bulk slot


def check_av_elwise(lib, nd, rs, seed=29):
    """Solver::update_art_visc_elwise (reference src/Solver.cpp:584-633) after set_uncertainty: the scalar ramp on Element::uncertainty
    (values below, inside and above the ramp window, zero and denormal-small included), the two point loops of the PDE-based branch, and
    the vertex-based branch (share_vertex_data(max) with hanging-vertex matchers + multilinear interpolation), against the numpy restatement"""
    import pyoracle
    from hexed_b200.kernels import UNCERT, VERTEX_SCRATCH
    rng = np.random.default_rng(seed)
    basis = hb.gauss_legendre(rs)
    m = M.soup_mesh(nd, rs, rng, n_car=14, n_def=16, n_ref=0)
    ne, n_vert = m.n_elem, 2**nd
    center = -4 - 4.25*np.log10(rs - 1)
    m.uncert = 10**(0.5*rng.uniform(center - 1.5, center + 1.5, ne))   # 2*log10(u) spread across the window [center - .5, center + .5]
    m.uncert[:4] = [0., 1e-300, 10**(0.5*(center - 0.5)), 10**(0.5*(center + 0.5))]
    m.elem_data[:, nd + 3:nd + 9] = rng.uniform(0., 1e-3, (ne, 6, m.nq))
    scale = 0.013/(rs - 1)*(1.7 + 2.3)
    n_vertex = ne*n_vert//3 + 5
    elem_vertex = np.stack([rng.choice(n_vertex, n_vert, replace=False) for _ in range(ne)]).astype(np.int32)
    matchers = []
    if nd > 1:
        free = list(rng.permutation(ne))
        for stretch in ([(0, 0), (1, 0), (0, 1)] if nd == 3 else [(0, 0)]):
            n_fine = 2**(nd - 1)//((1 + stretch[0])*(1 + stretch[1]))
            matchers.append([int(rng.integers(0, nd)), int(rng.integers(0, 2)), stretch[0], stretch[1]] + [int(free.pop()) for _ in range(n_fine)] + [-1]*(4 - n_fine))
    matchers = np.array(matchers, np.int32).reshape(-1, 8)
    interp = np.stack([1. - np.asarray(basis.node), np.asarray(basis.node)], axis=1)
    ref = m.copy()
    dev = Device(nd, rs, basis, lib_path=lib).load_mesh(m)
    dev.upload(UNCERT, m.uncert)
    dev.av_elwise_ramp(scale); pyoracle.av_elwise_ramp(ref, scale)
    got = dev.download(UNCERT, np.zeros(ne))
    assert np.all(got[:3] == 0.) and got[3] == scale and ((got > 0) & (got < scale)).sum() > 3   # all three branches of the ramp taken
    assert np.allclose(got, ref.uncert, rtol=1e-13, atol=1e-16*scale)   # (log / sin of the device library against libm's)
    ref.uncert[:] = got   # from here on the two must agree exactly
    dev.av_elwise_forcing(False); pyoracle.av_elwise_forcing(ref, False)
    out = m.copy(); dev.sync_to_host(out)
    assert np.array_equal(out.elem_data[:, nd + 3:nd + 9], ref.elem_data[:, nd + 3:nd + 9])
    junk = rng.uniform(0., 1e-3, (ne, m.nq))
    ref.elem_data[:, nd + 6] = junk; out.elem_data[:, nd + 6] = junk   # what diffuse_art_visc would have left in forcing[1]
    dev.upload_elements(out.elem_data)
    dev.av_elwise_forcing(True); pyoracle.av_elwise_forcing(ref, True)
    dev.sync_to_host(out)
    assert np.array_equal(out.elem_data[:, nd + 3:nd + 9], ref.elem_data[:, nd + 3:nd + 9])
    dev.vertex_topology(elem_vertex, n_vertex, matchers)
    dev.av_elwise_vertices(interp)
    v = pyoracle.av_elwise_vertices(ref, elem_vertex, n_vertex, matchers, interp)
    assert np.allclose(dev.download(VERTEX_SCRATCH, np.zeros((ne, n_vert))), v, rtol=4e-16, atol=0.)
    dev.sync_to_host(out)
    dev.close()
    assert rel_l2(out.elem_data[:, nd + 4], ref.elem_data[:, nd + 4]) <= 1e-15
    assert np.array_equal(out.elem_data[:, nd + 3], ref.elem_data[:, nd + 3])   # bulk coefficient untouched


def check_set_art_visc_admis(oracle, lib, nd, rs, seed=37):
    """Solver::set_art_visc_admis (reference src/Solver.cpp:636-658) with nothing on the host: stabilizing_art_visc leaves the desired viscosity
    in Element::uncertainty, share_vertex_data(vector_max) makes it C0 at the vertices (hanging-vertex matchers included), hypercube_matvec
    interpolates it to laplacian_av_coef = hexed_b200_stabilizing_art_visc + hexed_b200_vertex_topology + hexed_b200_av_elwise_vertices"""
    import pyoracle
    from hexed_b200.kernels import VERTEX_SCRATCH
    rng = np.random.default_rng(seed)
    basis = hb.gauss_legendre(rs)
    m = M.soup_mesh(nd, rs, rng, n_car=10, n_def=12, n_ref=0, with_ldg=False)
    M.random_flow_state(m, rng)
    m.state()[:, nd] *= 1 + 0.3*rng.random(m.state()[:, nd].shape)
    m.elem_data[:, nd + 3:nd + 5] = rng.uniform(0., 1e-3, (m.n_elem, 2, m.nq))
    ne, n_vert = m.n_elem, 2**nd
    n_vertex = ne*n_vert//3 + 5
    elem_vertex = np.stack([rng.choice(n_vertex, n_vert, replace=False) for _ in range(ne)]).astype(np.int32)
    matchers = np.array([[nd - 1, 1, 0, 0] + [2, 7, 11, 13][:2**(nd - 1)] + [-1]*(4 - 2**(nd - 1))], np.int32) if nd > 1 else np.zeros((0, 8), np.int32)
    interp = np.stack([1. - np.asarray(basis.node), np.asarray(basis.node)], axis=1)
    ref = m.copy()
    dev = Device(nd, rs, basis, lib_path=lib).load_mesh(m)
    oracle.stabilizing_art_visc(basis, ref, 340.)
    v = pyoracle.av_elwise_vertices(ref, elem_vertex, n_vertex, matchers, interp)
    dev.stabilizing_art_visc(340.)
    dev.vertex_topology(elem_vertex, n_vertex, matchers)
    dev.av_elwise_vertices(interp)
    got_v = dev.download(VERTEX_SCRATCH, np.zeros((ne, n_vert)))
    out = m.copy()
    dev.sync_to_host(out)
    dev.close()
    assert np.abs(out.uncert - ref.uncert).max() <= 1e-11*np.abs(ref.uncert).max()
    assert np.abs(got_v - v).max() <= 1e-11*np.abs(v).max()
    assert rel_l2(out.elem_data[:, nd + 4], ref.elem_data[:, nd + 4]) <= 1e-11
    assert np.array_equal(out.elem_data[:, nd + 3], ref.elem_data[:, nd + 3])


def check_shared_normals_soup(oracle, lib, nd, rs, seed=31):
    """the connection passes of Solver::calc_jacobian on a soup mesh (every direction, hanging faces, boundary ghosts) with arbitrary
    element-face normals: bit-identical to the numpy restatement (all factors are 0.5 and +-1)"""
    import pyoracle
    from hexed_b200.kernels import NORMALS
    rng = np.random.default_rng(seed)
    basis = hb.gauss_legendre(rs)
    m = M.soup_mesh(nd, rs, rng, n_car=8, n_def=20, n_ref=4 if nd > 1 else 0)
    m.face_state[:] = rng.normal(0., 1., m.face_state.shape)
    m.normals[:] = -9.
    ref = m.copy()
    dev = Device(nd, rs, basis, lib_path=lib).load_mesh(m)
    dev.calc_shared_normals()
    got_n = dev.download(NORMALS, np.zeros_like(m.normals))
    out = m.copy()
    dev.sync_to_host(out)
    dev.close()
    pyoracle.calc_shared_normals(oracle, basis, ref)
    written = ref.normals != -9.
    assert written.any()
    # bit-identical except where the prolonged mortar normals enter (they go through the interpolation matrices: FMA contraction)
    assert np.allclose(got_n[written], ref.normals[written], rtol=1e-13, atol=1e-14)
    assert rel_l2(out.face_state, ref.face_state) <= 1e-14


def check_calc_jacobian_box(lib, nd, rs, n):
    """calc_jacobian end to end on the device (set_jacobian + connection passes + vertex minimum of the time-step scale) from nothing
    but the vertex positions of the warped box reproduces the metric terms the mesh generator computed on the host"""
    import torch
    from hexed_b200.kernels import NORMALS, REF_NORMALS, JAC_DET, VERTEX_TSS
    basis = hb.gauss_legendre(rs)
    m = M.box_mesh(nd, rs, n, basis, deformed=True, bc_kind=M.BC_COPY)
    h = 1./n
    grid = torch.stack(torch.meshgrid(*[torch.arange(n + 1, dtype=torch.float64)*h]*nd, indexing="ij"), -1)
    vpos = M.default_warp(grid, 0.1*h).reshape(-1, nd).numpy()
    idx = np.stack(np.meshgrid(*[np.arange(n)]*nd, indexing="ij"), -1).reshape(-1, nd)
    corner = np.array([[(i >> (nd - 1 - d)) & 1 for d in range(nd)] for i in range(2**nd)])
    vstr = np.array([(n + 1)**(nd - 1 - d) for d in range(nd)])
    vid = ((idx[:, None, :] + corner[None, :, :])*vstr).sum(-1).astype(np.int32)
    want = dict(normals=np.asarray(m.normals).copy(), refn=np.asarray(m.ref_normals).copy(), det=np.asarray(m.det).copy(), vtss=m.vertex_tss.copy())
    m.normals = np.full_like(want["normals"], -9.); m.ref_normals = np.zeros_like(want["refn"]); m.det = np.zeros_like(want["det"])
    m.vertex_tss = np.zeros_like(want["vtss"])
    dev = Device(nd, rs, basis, lib_path=lib).load_mesh(m)
    dev.set_jacobian(vpos[vid])
    dev.calc_shared_normals()
    dev.vertex_topology(vid, (n + 1)**nd)
    dev.share_vertex_data(VERTEX_TSS, False)
    got = dict(normals=dev.download(NORMALS, np.zeros_like(want["normals"])), refn=dev.download(REF_NORMALS, np.zeros_like(want["refn"])),
               det=dev.download(JAC_DET, np.zeros_like(want["det"])), vtss=dev.download(VERTEX_TSS, np.zeros_like(want["vtss"])))
    dev.close()
    for key in want:
        assert rel_l2(got[key], want[key]) <= 1e-13, key
        assert np.abs(got[key] - want[key]).max() <= 1e-12*np.abs(want[key]).max(), key


def check_fused_admissibility(oracle, lib, nd, rs, n):
    """HEXED_B200_OPT_FUSED_ADMIS: the pipelined Local kernels leave the admissibility bits of what they write, is_admissible right
    after compute_euler reduces those instead of scanning the state. Same answer and record as the full scan and as the oracle, on an
    admissible step, on a step that drives two elements inadmissible, and after something else has touched the faces (fallback)."""
    from hexed_b200.kernels import OPT_FUSED_ADMIS, FACE_STATE
    basis = hb.gauss_legendre(rs)
    fs = freestream_state(nd)
    m = M.box_mesh(nd, rs, n, basis, deformed=True, bc_kind=M.BC_FREESTREAM, bc_params=fs)
    density_wave(m, basis)
    oracle.compute_write_face(basis, m)
    ref = m.copy()
    dev = Device(nd, rs, basis, lib_path=lib).load_mesh(m)
    dev.set_option(OPT_FUSED_ADMIS, 1)

    def stage(i_stage, dt):
        oracle.apply_state_bcs(ref); oracle.compute_euler(basis, ref, dt=dt, i_stage=i_stage)
        dev.apply_state_bcs(); dev.compute_euler(dt=dt, i_stage=i_stage)

    def both_ways():
        launches0 = dev.launch_count()
        fused = dev.is_admissible(); rec_fused = dev.record()
        n_fused = dev.launch_count() - launches0
        dev.set_option(OPT_FUSED_ADMIS, 0)       # clears the flags: full scan of the same device state
        full = dev.is_admissible(); rec_full = dev.record()
        dev.set_option(OPT_FUSED_ADMIS, 1)
        want, rec_want = oracle.is_admissible(ref)
        assert fused == full == want and np.array_equal(rec_fused, rec_full) and np.array_equal(rec_full, rec_want)
        return fused, rec_fused, n_fused
    dt = oracle.max_dt(EULER, basis, ref, 0.5, 0.5, False); dev.max_dt_euler(0.5, 0.5, False)
    stage(0, dt)
    ok, rec, _ = both_ways()
    assert ok and not rec.any()
    stage(1, dt)
    ok, rec, _ = both_ways()
    assert ok
    # a time step far beyond the stability limit overshoots: finite, but mass / energy go non-positive in many elements
    snapshot_dev, snapshot_ref = m.copy(), ref.copy()
    dev.sync_to_host(snapshot_dev)
    stage(0, 3e3*dt)
    ok, rec, _ = both_ways()
    assert not ok and 0 < rec.sum()
    for a, b in ((snapshot_dev, m), (snapshot_ref, ref)):   # back to the admissible state on both sides
        b.elem_data[:] = a.elem_data; b.face_state[:] = a.face_state
    dev.upload_elements(m.elem_data); dev.upload(FACE_STATE, m.face_state)
    assert dev.is_admissible() and oracle.is_admissible(ref)[0]
    # fallback: the flags must not survive a write to the element faces by anybody else
    stage(1, dt)
    f = dev.download(FACE_STATE, np.zeros_like(m.face_state))
    f[5].reshape(nd + 2, -1)[nd, 0] = -1.
    ref.face_state[5].reshape(nd + 2, -1)[nd, 0] = -1.
    dev.upload(FACE_STATE, f)
    got = dev.is_admissible()
    assert got == oracle.is_admissible(ref)[0] and not got
    assert dev.record()[5//(2*nd)] == 1
    dev.close()


def _cheby(n, i):  # math::chebyshev_step, reference src/math.cpp:104-107
    return 1/(1 - np.cos((n - i - .5)*np.pi/n))


def check_update_euler(oracle, lib, nd, rs, n, n_steps, use_graph, deformed=False, bc="riemann", refined=False, n_cheby=1):
    """hexed_b200_update_euler (time step kept on the device, optional CUDA graph) is bit-identical to the same steps made call by call
    through the reference-shaped entry points, and both track the oracle"""
    basis = hb.gauss_legendre(rs)
    fs = freestream_state(nd)
    kind = M.BC_RIEMANN_INVARIANTS if bc == "riemann" else M.BC_FREESTREAM
    if refined:  # hanging-node faces: Restrict / Prolong inside every stage of the captured step
        refine = np.zeros((n,)*nd, bool)
        refine[(slice(n//3, max(n//3 + 1, 2*n//3)),)*nd] = True
        m = M.refined_box_mesh(nd, rs, n, basis, refine, bc_kind=M.BC_NONPENETRATION)
        assert m.ref_face.shape[0] > 0
    else:
        m = M.box_mesh(nd, rs, n, basis, deformed=deformed, bc_kind=kind, bc_params=fs)
    density_wave(m, basis)
    oracle.compute_write_face(basis, m)
    if refined:
        oracle.compute_prolong(basis, m)
    ref = m.copy()
    a = Device(nd, rs, basis, lib_path=lib).load_mesh(m)
    b = Device(nd, rs, basis, lib_path=lib).load_mesh(m)
    t_a = 0.
    max_cheby = _cheby(n_cheby, n_cheby - 1)
    for i_step in range(n_steps):  # Solver::update, reference src/Solver.cpp:842-851
        f = _cheby(n_cheby, i_step % n_cheby)
        dt = a.max_dt_euler(0.4/max_cheby, 0.4, False)*f
        dt_o = oracle.max_dt(EULER, basis, ref, 0.4/max_cheby, 0.4, False)*f
        for stage in (0, 1):
            a.apply_state_bcs(); a.compute_euler(dt=dt, i_stage=stage)
            oracle.apply_state_bcs(ref); oracle.compute_euler(basis, ref, dt=dt_o, i_stage=stage)
        t_a += dt
    dt_b, t_b = b.update_euler(0.4, n_steps, use_graph, n_cheby=n_cheby)
    out_a, out_b = m.copy(), m.copy()
    a.sync_to_host(out_a); b.sync_to_host(out_b)
    a.close(); b.close()
    assert dt_b == dt and t_b == t_a
    assert np.array_equal(out_a.elem_data, out_b.elem_data) and np.array_equal(out_a.face_state, out_b.face_state)
    assert rel_l2(out_b.state(), ref.state()) <= STATE_TOL


def check_update_navier_stokes(oracle, lib, nd, rs, n, n_steps, use_graph, deformed=True, n_cheby=1):
    """hexed_b200_update_navier_stokes (time step on the device, optional CUDA graph) is bit-identical to the same viscous steps made
    call by call, and both track the oracle"""
    import pyoracle
    from hexed_b200 import kernels as K
    basis = hb.gauss_legendre(rs)
    m = M.box_mesh(nd, rs, n, basis, deformed=deformed, bc_kind=M.BC_NONPENETRATION, with_ldg=True)
    density_wave(m, basis)
    oracle.compute_write_face(basis, m)
    ref = m.copy()
    visc_o, cond_o = pyoracle.sutherland(1.716e-5, 273., 111.), pyoracle.sutherland(.0241, 273., 194.)
    visc_d, cond_d = K.sutherland(1.716e-5, 273., 111.), K.sutherland(.0241, 273., 194.)
    a = Device(nd, rs, basis, lib_path=lib).load_mesh(m)
    b = Device(nd, rs, basis, lib_path=lib).load_mesh(m)
    t_a = 0.
    max_cheby = _cheby(n_cheby, n_cheby - 1)
    for i_step in range(n_steps):
        f = _cheby(n_cheby, i_step % n_cheby)
        dt = a.max_dt_navier_stokes(0.3/max_cheby, 0.3, False, visc_d, cond_d)*f
        dt_o = oracle.max_dt(NAVIER_STOKES, basis, ref, 0.3/max_cheby, 0.3, False, visc_o, cond_o)*f
        a.apply_state_bcs(); a.compute_navier_stokes(a.apply_flux_bcs, visc_d, cond_d, dt=dt, i_stage=0)
        a.apply_state_bcs(); a.compute_euler(dt=dt, i_stage=1)
        oracle.apply_state_bcs(ref); oracle.compute_navier_stokes(basis, ref, lambda: oracle.apply_flux_bcs(ref), visc_o, cond_o, dt=dt_o, i_stage=0)
        oracle.apply_state_bcs(ref); oracle.compute_euler(basis, ref, dt=dt_o, i_stage=1)
        t_a += dt
    dt_b, t_b = b.update_navier_stokes(0.3, visc_d, cond_d, n_steps, use_graph, n_cheby=n_cheby)
    out_a, out_b = m.copy(), m.copy()
    a.sync_to_host(out_a); b.sync_to_host(out_b)
    a.close(); b.close()
    assert dt_b == dt and t_b == t_a
    assert np.array_equal(out_a.elem_data, out_b.elem_data) and np.array_equal(out_a.face_state, out_b.face_state)
    assert np.array_equal(out_a.face_ldg, out_b.face_ldg)
    assert rel_l2(out_b.state(), ref.state()) <= STATE_TOL
