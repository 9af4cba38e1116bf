"""The C++ adapter (hexed_b200/host/adapter.cpp): `hexed::compute_euler(Kernel_mesh, Kernel_options)` & co. with the reference's
signatures (include/kernels.hpp:22-42), driven on a reference-shaped pointer-graph mesh (hexed_b200/host/harness.cpp).

CPU part (no marker): the device-free flattening reproduces the integer tables the mesh was built from bit for bit, the
library exports every symbol, and the whole adapter -> C ABI -> kernel-source chain agrees with the oracle when the C ABI is the
host-thread emulation build of the same .cu files. GPU part (-m gpu): the same calls through libhexed_b200.so on the B200.
The boundary conditions are applied ON THE HOST between calls (oracle code on the fetched FlatMesh), the way Solver does."""
import ctypes

import numpy as np
import pytest

import hexed_b200 as hb
from hexed_b200.cases import density_wave, freestream_state
from hexed_b200 import mesh as M
from hexed_b200.tables import Connection_direction
import host_harness as H
from pyoracle import EULER, NAVIER_STOKES, ADVECTION, SMOOTH_AV, FIX_THERM_ADMIS
import pyoracle
from util import rel_l2, assert_euler_parity, assert_pde_parity, prepare_pde_state, density_wave, freestream_state


@pytest.fixture(scope="module")
def host_emu(emu_lib):
    return H.build(emu=True)


@pytest.fixture(scope="module")
def host_gpu(gpu_lib):
    return H.build(emu=False)


def soup(nd, rs, seed, **kw):
    rng = np.random.default_rng(seed)
    m = M.soup_mesh(nd, rs, rng, **kw)
    M.random_flow_state(m, rng)
    return m, rng


# ------------------------------------------------------------------------------------------------ flattening (integer-exact)
@pytest.mark.parametrize("nd,rs", [(1, 2), (2, 3), (2, 6), (3, 2), (3, 6)])
def test_flatten_reproduces_tables(host_emu, nd, rs):
    """pointer graph -> slot tables must give back exactly the tables the pointer graph was made from: connection order,
    face slots of both sides, direction codes, normal slots, refined-face fine ordering and stretch flags, boundary detection"""
    m, _ = soup(nd, rs, 5 + nd, n_car=8, n_def=14, n_ref=6)
    for seed in (1, 2):  # two different heap layouts
        h = H.HostHarness(host_emu, m, hb.gauss_legendre(rs), seed=seed)
        counts, car, dfc, ref, bnd = h.flatten()
        assert counts[0] == m.n_car and counts[1] == m.n_def
        assert counts[2] == m.n_face_slot and counts[3] == m.n_normal_slot
        assert np.array_equal(car, m.car_con)
        assert np.array_equal(dfc, m.def_con)
        assert np.array_equal(ref, m.ref_face)
        assert np.array_equal(bnd, m.bcs[0]["con_index"])
        assert counts[5] == int((H.normal_present(m) == 0).sum())  # faces that fall back to the unit normal
        h.close()


def test_flatten_box_with_cartesian_boundaries(host_emu):
    """a Cartesian box: boundary connections of Cartesian elements sit in def_cons (src/Accessible_mesh.cpp:136-147)"""
    basis = hb.gauss_legendre(3)
    m = M.box_mesh(2, 3, 3, basis, deformed=False, bc_kind=M.BC_COPY)
    h = H.HostHarness(host_emu, m, basis)
    counts, car, dfc, ref, bnd = h.flatten()
    assert np.array_equal(car, m.car_con) and np.array_equal(dfc, m.def_con)
    assert sorted(bnd.tolist()) == sorted(np.concatenate([bc["con_index"] for bc in m.bcs]).tolist())
    h.close()


# ------------------------------------------------------------------------------------------------ drivers shared by emu and GPU
def host_bcs(oracle, h, work, resident, flux=False):
    """what Solver::apply_state_bcs / apply_flux_bcs do: read the inside faces of the host objects, write the ghost faces"""
    lean = host_bcs.lean and resident and not flux
    if lean:
        h.inside_state_faces_to_host()  # asynchronous route: collected from the download the previous stage driver started
    elif resident:
        h.boundary_faces_to_host()
    h.fetch(work)
    (oracle.apply_flux_bcs if flux else oracle.apply_state_bcs)(work)
    h.put(work)
    if lean:
        h.ghost_state_faces_to_device()  # lands in the face storage inside the next stage driver, after its interior Neighbor kernels
    elif resident:
        h.ghost_faces_to_device()


host_bcs.lean = False


def run_euler(oracle, lib, m, basis, resident, n_steps=2, local_time=False, use_filter=False, safety=0.7, devices=None, coords=None):
    ref, work = m.copy(), m.copy()
    h = H.HostHarness(lib, m, basis, seed=3)
    if devices is not None:
        h.set_devices(devices)
        if coords is not None:
            h.set_element_coordinates(coords)
    h.set_sync_mode(H.RESIDENT if resident else H.SYNC_EVERY_CALL)
    h.invalidate()
    dts = []
    for _ in range(n_steps):
        dt_o = oracle.max_dt(EULER, basis, ref, safety, safety, local_time)
        dt_d = h.call("max_dt_euler", safety, safety, local_time)
        dts.append((dt_d, dt_o))
        for stage in (0, 1):
            oracle.apply_state_bcs(ref)
            oracle.compute_euler(basis, ref, dt=dt_o, i_stage=stage, use_filter=use_filter)
            host_bcs(oracle, h, work, resident)
            h.call("compute_euler", dt=dt_o, i_stage=stage, use_filter=use_filter)
    if resident:
        h.to_host(H.ALL_ELEM | H.FACES | H.UNCERT)
    h.fetch(work)
    units = h.work_units()
    if devices is not None:
        run_euler.last_owners = h.element_owners().copy()
        run_euler.last_transport = h.transport_description()
    h.set_sync_mode(H.SYNC_EVERY_CALL)
    h.close()
    if devices is not None:
        h2 = H.HostHarness(lib, m, basis, seed=3); h2.set_devices([0]); h2.close()
    return work, ref, dts, units


def check_work_units(m, units, n_steps, diffusive_stage0=0):
    n_cc, n_dc, n_ref = m.car_con.shape[0], m.def_con.shape[0], m.ref_face.shape[0]
    stages = 2*n_steps
    assert units[0] == n_cc*stages and units[4] == n_dc*stages          # "neighbor"
    assert units[1] == m.n_car*stages and units[5] == m.n_def*stages    # "local"
    assert units[3] == m.n_car*n_steps and units[7] == m.n_def*n_steps  # "compute time step"
    assert units[8] == 2*n_ref*stages                                   # restrict + prolong through sw_pr
    assert units[9] > 0


def run_pde(oracle, lib, m, basis, pde, resident, devices=None):
    ref, work = m.copy(), m.copy()
    wide = pde == ADVECTION
    h = H.HostHarness(lib, m, basis, seed=4)
    if devices is not None:
        h.set_devices(devices)
    if wide:
        h.put(m, wide=True)
    h.set_sync_mode(H.RESIDENT if resident else H.SYNC_EVERY_CALL)
    h.invalidate()
    visc_o, cond_o = pyoracle.sutherland(1.7e-5, 273., 110.), pyoracle.constant(2.5e-2)
    visc_h, cond_h = H.sutherland(1.7e-5, 273., 110.), H.constant(2.5e-2)
    s = 0.3
    if pde == NAVIER_STOKES:
        dt_o = oracle.max_dt(pde, basis, ref, s, s, False, visc_o, cond_o)
        dt_d = h.call("max_dt_navier_stokes", s, s, False, *visc_h, *cond_h)
        oracle.apply_state_bcs(ref)
        oracle.compute_navier_stokes(basis, ref, lambda: oracle.apply_flux_bcs(ref), visc_o, cond_o, dt=dt_o, i_stage=0)
        host_bcs(oracle, h, work, resident)

        def flux_bc():  # resident mode: the callback owns its face traffic, like Solver::apply_flux_bcs with the two added lines
            host_bcs(oracle, h, work, resident, flux=True)
        h.set_flux_bc(flux_bc)
        h.call("compute_navier_stokes", *visc_h, *cond_h, dt=dt_o, i_stage=0)
        oracle.apply_state_bcs(ref)
        oracle.compute_euler(basis, ref, dt=dt_o, i_stage=1)
        host_bcs(oracle, h, work, resident)
        h.call("compute_euler", dt=dt_o, i_stage=1)
    elif pde == ADVECTION:
        dt_o = oracle.max_dt(pde, basis, ref, s, s, False, advect_length=0.7)
        dt_d = h.call("max_dt_advection", s, s, False, 0.7)
        for stage in (0, 1):
            oracle.compute_advection(basis, ref, 0.7, dt=dt_o, i_stage=stage)
            h.call("compute_advection", 0.7, dt=dt_o, i_stage=stage)
    elif pde == SMOOTH_AV:
        dt_o = oracle.max_dt(pde, basis, ref, s, s, False)
        dt_d = h.call("max_dt_smooth_av", s, s, False)
        oracle.compute_smooth_av(basis, ref, None, 0.4, 1.3, dt=dt_o, i_stage=0)
        h.call("compute_smooth_av", 0.4, 1.3, dt=dt_o, i_stage=0)
    else:
        dt_o = oracle.max_dt(pde, basis, ref, s, s, False)
        dt_d = h.call("max_dt_fix_therm_admis", s, s, False)
        oracle.compute_fix_therm_admis(basis, ref, None, dt=dt_o, i_stage=0)
        h.call("compute_fix_therm_admis", dt=dt_o, i_stage=0)
    if resident:
        h.to_host(H.ALL_ELEM | (H.FACES_WIDE if wide else H.FACES))
    h.fetch(work, wide=wide)
    if wide:  # the narrow views of the same host storage are not what this PDE wrote
        work.face_state, work.face_ldg, ref.face_state, ref.face_ldg = None, None, None, None
    else:
        work.face_wide, ref.face_wide = None, None
    h.set_sync_mode(H.SYNC_EVERY_CALL)
    h.close()
    if devices is not None:
        h2 = H.HostHarness(lib, m, basis, seed=4); h2.set_devices([0]); h2.close()
    return work, ref, [(dt_d, dt_o)]


# ------------------------------------------------------------------------------------------------ CPU: adapter on the emulation build
@pytest.mark.parametrize("resident", [False, True])
@pytest.mark.parametrize("nd,rs", [(2, 3), (3, 2)])
def test_adapter_euler_emu(oracle, host_emu, nd, rs, resident):
    m, _ = soup(nd, rs, 21, with_ldg=True)
    out, ref, dts, units = run_euler(oracle, host_emu, m, hb.gauss_legendre(rs), resident, n_steps=2)
    assert_euler_parity(out, ref, dts)
    check_work_units(m, units, 2)


@pytest.mark.parametrize("devices", [None, [0, 0, 0]])
def test_adapter_async_boundary_traffic_emu(oracle, host_emu, devices):
    """resident mode with only (inside, state) down and (ghost, state) up, prefetched / deferred (adapter.hpp): same answer"""
    m, _ = soup(2, 4, 23, with_ldg=True)
    oracle.compute_prolong(hb.gauss_legendre(4), m)
    host_bcs.lean = True
    try:
        out, ref, dts, _ = run_euler(oracle, host_emu, m, hb.gauss_legendre(4), True, n_steps=3, devices=devices)
    finally:
        host_bcs.lean = False
    assert_euler_parity(out, ref, dts)


def test_adapter_euler_options_emu(oracle, host_emu):
    m, _ = soup(2, 4, 22, with_ldg=True)
    out, ref, dts, _ = run_euler(oracle, host_emu, m, hb.gauss_legendre(4), False, n_steps=1, local_time=True, use_filter=True, safety=0.05)
    assert_euler_parity(out, ref, dts)


@pytest.mark.parametrize("resident", [False, True])
@pytest.mark.parametrize("pde", [NAVIER_STOKES, ADVECTION, SMOOTH_AV, FIX_THERM_ADMIS])
def test_adapter_other_pdes_emu(oracle, host_emu, pde, resident):
    m, rng = soup(2, 3, 30 + pde, with_ldg=True, with_wide=True)
    prepare_pde_state(m, rng, pde)
    out, ref, dts = run_pde(oracle, host_emu, m, hb.gauss_legendre(3), pde, resident)
    assert_pde_parity(out, ref, dts)


# ---- one Kernel_mesh on several devices BEHIND the kernels.hpp boundary (hexed_b200::set_devices). The emulation build
# runs the same partitioning, halo bookkeeping and stage splitting with memcpy where the product posts the NCCL group.
@pytest.mark.parametrize("resident", [False, True])
@pytest.mark.parametrize("nd,rs,n_dev", [(2, 3, 2), (2, 3, 3), (3, 2, 4)])
def test_adapter_multi_device_euler_emu(oracle, host_emu, nd, rs, n_dev, resident):
    m, _ = soup(nd, rs, 61, with_ldg=True)  # hanging faces, every direction; the graph ordering cuts through all of it
    oracle.compute_prolong(hb.gauss_legendre(rs), m)  # the soup's faces are random: make the mortar faces what they are after any stage
    n_steps = 2 if nd == 2 else 1  # (the meaningless 3-D soup geometry goes non-finite in its fourth stage, reference included)
    out, ref, dts, units = run_euler(oracle, host_emu, m, hb.gauss_legendre(rs), resident, n_steps=n_steps, devices=[0]*n_dev)
    assert np.isfinite(ref.state()).all()
    assert_euler_parity(out, ref, dts)
    check_work_units(m, units, n_steps)
    assert len(set(run_euler.last_owners.tolist())) == n_dev


@pytest.mark.parametrize("resident", [False, True])
def test_adapter_multi_device_box_morton_emu(oracle, host_emu, resident):
    basis = hb.gauss_legendre(3)
    m = M.box_mesh(3, 3, 4, basis, deformed=True, bc_kind=M.BC_FREESTREAM, bc_params=freestream_state(3))
    density_wave(m, basis)
    oracle.compute_write_face(basis, m)
    out, ref, dts, _ = run_euler(oracle, host_emu, m, basis, resident, n_steps=2, devices=[0]*8, coords=m.elem_index)
    assert_euler_parity(out, ref, dts)
    from hexed_b200 import partition as P
    assert np.array_equal(run_euler.last_owners, P.split_by_curve(P.morton_keys(m.elem_index), 8))  # 8 octants of 8 elements


@pytest.mark.parametrize("resident", [False, True])
def test_adapter_multi_device_refined_box_emu(oracle, host_emu, resident):
    """the onera_m6 class: Cartesian hanging-node faces cut by the device split"""
    basis = hb.gauss_legendre(3)
    refine = np.zeros((4, 4), bool); refine[1, 1] = refine[2, 2] = refine[3, 0] = True
    m = M.refined_box_mesh(2, 3, 4, basis, refine, bc_kind=M.BC_COPY)
    M.random_flow_state(m, np.random.default_rng(5), mach=0.2)
    oracle.compute_write_face(basis, m); oracle.compute_prolong(basis, m)
    out, ref, dts, _ = run_euler(oracle, host_emu, m, basis, resident, n_steps=2, devices=[0]*3)
    assert_euler_parity(out, ref, dts)


@pytest.mark.parametrize("resident", [False, True])
def test_adapter_multi_device_navier_stokes_emu(oracle, host_emu, resident):
    m, rng = soup(2, 3, 62, with_ldg=True, with_wide=True)
    prepare_pde_state(m, rng, NAVIER_STOKES)
    oracle.compute_prolong(hb.gauss_legendre(3), m)
    out, ref, dts = run_pde(oracle, host_emu, m, hb.gauss_legendre(3), NAVIER_STOKES, resident, devices=[0, 0, 0])
    assert_pde_parity(out, ref, dts)


def test_adapter_multi_device_unsupported_entry_points_say_so(oracle, host_emu):
    m, rng = soup(2, 3, 63, with_ldg=True, with_wide=True)
    h = H.HostHarness(host_emu, m, hb.gauss_legendre(3))
    h.set_devices([0, 0])
    with pytest.raises(RuntimeError, match="not available on more than one device"):
        h.call("compute_advection", 0.7, dt=1e-4, i_stage=0)
    h.set_devices([0])
    h.close()


@pytest.mark.parametrize("resident", [False, True])
def test_adapter_device_bcs_emu(oracle, host_emu, resident):
    check_adapter_device_bcs(oracle, host_emu, resident, 2, 3)


@pytest.mark.gpu
@pytest.mark.parametrize("resident", [False, True])
def test_adapter_device_bcs_gpu(oracle, host_gpu, resident):
    check_adapter_device_bcs(oracle, host_gpu, resident, 3, 4)


def test_adapter_device_bc_params_change_emu(oracle, host_emu):
    """hexed_b200::set_device_bc_params: a freestream state that changes between steps (stream-ordered update of the parameter block)"""
    basis = hb.gauss_legendre(3)
    fs = freestream_state(2)
    m = M.box_mesh(2, 3, 4, basis, deformed=True, bc_kind=M.BC_FREESTREAM, bc_params=fs)
    density_wave(m, basis)
    oracle.compute_write_face(basis, m)
    ref, work = m.copy(), m.copy()
    h = H.HostHarness(host_emu, m, basis, seed=4)
    h.set_sync_mode(H.RESIDENT)
    h.invalidate()
    dts = []
    for step in range(3):
        new_fs = np.asarray(fs)*(1. + 0.01*step)
        ref.bcs[0]["params"] = new_fs
        dt_o = oracle.max_dt(EULER, basis, ref, 0.5, 0.5, False)
        dt_d = h.call("max_dt_euler", 0.5, 0.5, False)
        if step == 0:
            h.add_device_bcs(m)
        h.set_device_bc_params(0, new_fs)
        dts.append((dt_d, dt_o))
        for stage in (0, 1):
            oracle.apply_state_bcs(ref); h.apply_state_bcs()
            oracle.compute_euler(basis, ref, dt=dt_o, i_stage=stage); h.call("compute_euler", dt=dt_o, i_stage=stage)
    h.to_host(H.ALL_ELEM | H.FACES)
    h.fetch(work)
    h.set_sync_mode(H.SYNC_EVERY_CALL)
    h.close()
    assert_euler_parity(work, ref, dts)
    with pytest.raises(RuntimeError):
        h2 = H.HostHarness(host_emu, m, basis, seed=4)
        try:
            h2.call("max_dt_euler", 0.5, 0.5, False)
            h2.set_device_bc_params(7, fs)
        finally:
            h2.close()


@pytest.mark.parametrize("resident", [False, True])
def test_adapter_is_admissible_emu(oracle, host_emu, resident):
    """hexed_b200::is_admissible through the pointer-graph adapter: same answer and Element::record as the oracle's restatement of
    Solver::is_admissible, and the reference's exception text for a non-finite state"""
    nd, rs = 2, 3
    m, rng = soup(nd, rs, 12, n_car=7, n_def=9, n_ref=2)
    basis = hb.gauss_legendre(rs)
    oracle.compute_write_face(basis, m)
    oracle.compute_prolong(basis, m)
    m.state()[4, nd, 2] = -1.
    m.face_state[2*nd*9 + 3].reshape(nd + 2, -1)[nd + 1, 1] = 0.
    want, want_rec = oracle.is_admissible(m)
    assert not want and want_rec.sum() == 2
    h = H.HostHarness(host_emu, m, basis, seed=4)
    h.set_sync_mode(H.RESIDENT if resident else H.SYNC_EVERY_CALL)
    h.invalidate()
    if resident:
        h.to_device(H.ALL_ELEM | H.FACES)
    got, rec = h.is_admissible()
    assert got == want and np.array_equal(rec, want_rec)
    m.state()[4, nd, 2] = np.nan
    h.put(m)
    if resident:
        h.to_device(H.ALL_ELEM | H.FACES)
    with pytest.raises(RuntimeError, match="state is not finite"):
        h.is_admissible()
    h.set_sync_mode(H.SYNC_EVERY_CALL)
    h.close()


def test_adapter_is_admissible_after_stages_emu(oracle, host_emu):
    """resident mode, the way Solver::update would use it: device boundary conditions, compute_euler, is_admissible after every stage.
    The adapter switches the fused bits on at the first check, so the later ones take the 4-bytes-per-element path; same answers as
    the oracle throughout, including after an over-long time step that leaves part of the mesh inadmissible."""
    from util import density_wave, freestream_state
    nd, rs = 2, 4
    basis = hb.gauss_legendre(rs)
    m = M.box_mesh(nd, rs, 5, basis, deformed=True, bc_kind=M.BC_FREESTREAM, bc_params=freestream_state(nd))
    density_wave(m, basis)
    oracle.compute_write_face(basis, m)
    ref = m.copy()
    h = H.HostHarness(host_emu, m, basis, seed=2)
    h.set_sync_mode(H.RESIDENT)
    h.invalidate()
    dt = oracle.max_dt(EULER, basis, ref, 0.5, 0.5, False)
    h.call("max_dt_euler", 0.5, 0.5, False)
    h.add_device_bcs(m)
    seen = []
    for step_dt in (dt, dt, 3e3*dt):
        for stage in (0, 1):
            oracle.apply_state_bcs(ref); oracle.compute_euler(basis, ref, dt=step_dt, i_stage=stage)
            h.apply_state_bcs(); h.call("compute_euler", dt=step_dt, i_stage=stage)
            try:
                want, want_rec = oracle.is_admissible(ref)
            except RuntimeError:
                with pytest.raises(RuntimeError, match="state is not finite"):
                    h.is_admissible()
                seen.append("nonfinite")
                break
            got, rec = h.is_admissible()
            assert got == want and np.array_equal(rec, want_rec)
            seen.append(got)
        if seen and seen[-1] == "nonfinite":
            break
    h.set_sync_mode(H.SYNC_EVERY_CALL)
    h.close()
    assert seen[:4] == [True]*4 and (False in seen or "nonfinite" in seen)


@pytest.mark.parametrize("resident", [False, True])
def test_adapter_av_glue_emu(oracle, host_emu, resident):
    """hexed_b200::av_scale_velocity / av_project_forcing / av_finish / interp_vertices / av_swap through the pointer-graph adapter (the
    basis operators come from `Kernel_mesh::basis`), against the numpy restatements, in both coherence modes"""
    nd, rs = 2, 3
    m, rng = soup(nd, rs, 19, n_car=6, n_def=8, n_ref=0)
    basis = hb.gauss_legendre(rs)
    m.elem_data[:, nd + 3:nd + 5] = rng.uniform(0., 2e-3, (m.n_elem, 2, m.nq))
    m.elem_data[:, nd + 5:nd + 9] = rng.uniform(0., 1e-3, (m.n_elem, 4, m.nq))
    m.elem_data[:, nd + 9:nd + 9 + rs] = rng.normal(1., .3, (m.n_elem, rs, m.nq))
    m.nom_size = rng.choice([.25, .5, 1.], m.n_elem)
    ref, work = m.copy(), m.copy()
    w = np.asarray(basis.weight); orth = np.asarray(basis.orthogonal).reshape(rs, rs)[rs - 1]
    vert = rng.uniform(0., 1., (m.n_elem, 2**nd))
    interp = np.stack([1. - np.asarray(basis.node), np.asarray(basis.node)], axis=1)
    h = H.HostHarness(host_emu, m, basis, seed=6)
    h.set_sync_mode(H.RESIDENT if resident else H.SYNC_EVERY_CALL)
    h.invalidate()
    if resident:
        h.to_device(H.ALL_ELEM | H.FACES)
    h.av_glue(0); pyoracle.av_scale_velocity(ref)
    h.av_glue(1); pyoracle.av_project_forcing(ref, w, orth)
    got = h.av_glue(2, 0.7, 3e-3, 3); want = pyoracle.av_finish(ref, 0.7, 3e-3, 3, w)
    h.av_glue(3, n=1, values=vert); pyoracle.interp_vertices(ref, 1, vert, interp)
    h.av_glue(4); pyoracle.av_swap(ref)
    if resident:
        h.to_host(H.ALL_ELEM)
    h.fetch(work)
    h.set_sync_mode(H.SYNC_EVERY_CALL)
    h.close()
    assert abs(got - want) <= 1e-12*abs(want)
    for lo, hi in ((0, nd + 2), (nd + 3, nd + 5), (nd + 5, nd + 9)):
        assert rel_l2(work.elem_data[:, lo:hi], ref.elem_data[:, lo:hi]) <= 1e-14


@pytest.mark.parametrize("resident", [False, True])
def test_adapter_av_elwise_emu(oracle, host_emu, resident):
    """hexed_b200::av_elwise_ramp / av_elwise_forcing / vertex_topology / av_elwise_vertices: the loops of Solver::update_art_visc_elwise
    (src/Solver.cpp:584-633) on Element::uncertainty of the host objects, against the numpy restatement"""
    nd, rs = 2, 3
    m, rng = soup(nd, rs, 27, n_car=6, n_def=8, n_ref=0)
    basis = hb.gauss_legendre(rs)
    ne, n_vert = m.n_elem, 2**nd
    center = -4 - 4.25*np.log10(rs - 1)
    m.uncert = 10**(0.5*rng.uniform(center - 1.5, center + 1.5, ne))
    m.uncert[0] = 0.
    m.elem_data[:, nd + 3:nd + 9] = rng.uniform(0., 1e-3, (ne, 6, m.nq))
    n_vertex = ne*n_vert//3 + 5
    elem_vertex = np.stack([rng.choice(n_vertex, n_vert, replace=False) for _ in range(ne)]).astype(np.int32)
    matchers = np.array([[1, 0, 0, 0, 3, 5, -1, -1]], np.int32)
    interp = np.stack([1. - np.asarray(basis.node), np.asarray(basis.node)], axis=1)
    ref, work = m.copy(), m.copy()
    h = H.HostHarness(host_emu, m, basis, seed=6)
    h.put_uncert(m.uncert)
    h.set_sync_mode(H.RESIDENT if resident else H.SYNC_EVERY_CALL)
    h.invalidate()
    if resident:
        h.to_device(H.ALL_ELEM | H.FACES)
    scale = 0.02
    h.av_glue(6, scale); pyoracle.av_elwise_ramp(ref, scale)
    h.fetch(work)   # the ramped value is back in Element::uncertainty in either mode
    assert work.uncert[0] == 0. and np.allclose(work.uncert, ref.uncert, rtol=1e-13, atol=1e-18)
    ref.uncert[:] = work.uncert
    h.av_glue(7, n=0); pyoracle.av_elwise_forcing(ref, False)
    h.av_glue(7, n=1); pyoracle.av_elwise_forcing(ref, True)
    h.av_glue(9, values=np.concatenate([[n_vertex], elem_vertex.reshape(-1), matchers.reshape(-1)]).astype(np.float64))
    h.av_glue(8); pyoracle.av_elwise_vertices(ref, elem_vertex, n_vertex, matchers, interp)
    if resident:
        h.to_host(H.ALL_ELEM)
    h.fetch(work)
    h.set_sync_mode(H.SYNC_EVERY_CALL)
    h.close()
    assert rel_l2(work.elem_data[:, nd + 3:nd + 9], ref.elem_data[:, nd + 3:nd + 9]) <= 1e-15


def check_adapter_device_bcs(oracle, host_emu, resident, nd, rs):
    """hexed_b200::add_device_bc / apply_state_bcs / apply_flux_bcs: a viscous step with every device-side boundary condition and no
    host boundary loop at all"""
    from util import mixed_bcs
    m, rng = soup(nd, rs, 45, n_car=10, n_def=24, n_ref=2, with_ldg=True)
    prepare_pde_state(m, rng, NAVIER_STOKES)
    assert m.bcs[0]["ghost_slot"].size >= 8
    mixed_bcs(m, rng)
    basis = hb.gauss_legendre(rs)
    ref, work = m.copy(), m.copy()
    h = H.HostHarness(host_emu, m, basis, seed=9)
    h.set_sync_mode(H.RESIDENT if resident else H.SYNC_EVERY_CALL)
    h.invalidate()
    visc_o, cond_o = pyoracle.sutherland(1.7e-5, 273., 110.), pyoracle.constant(2.5e-2)
    visc_h, cond_h = H.sutherland(1.7e-5, 273., 110.), H.constant(2.5e-2)
    dt = oracle.max_dt(NAVIER_STOKES, basis, ref, 0.3, 0.3, False, visc_o, cond_o)
    dt_d = h.call("max_dt_navier_stokes", 0.3, 0.3, False, *visc_h, *cond_h)  # first call of the epoch: the mirror exists from here on
    h.add_device_bcs(m)
    oracle.apply_state_bcs(ref)
    oracle.compute_navier_stokes(basis, ref, lambda: oracle.apply_flux_bcs(ref), visc_o, cond_o, dt=dt, i_stage=0)
    oracle.apply_state_bcs(ref)
    oracle.compute_euler(basis, ref, dt=dt, i_stage=1)
    h.apply_state_bcs()
    h.set_flux_bc(h.apply_flux_bcs)
    h.call("compute_navier_stokes", *visc_h, *cond_h, dt=dt, i_stage=0)
    h.apply_state_bcs()
    h.call("compute_euler", dt=dt, i_stage=1)
    if resident:
        h.to_host(H.ALL_ELEM | H.FACES)
    h.fetch(work)
    work.face_wide, ref.face_wide = None, None
    h.set_sync_mode(H.SYNC_EVERY_CALL)
    h.close()
    assert_pde_parity(work, ref, [(dt_d, dt)])


def test_adapter_standalone_entry_points_emu(oracle, host_emu):
    """write_face, prolong, restrict, stabilizing_art_visc through the reference signatures"""
    basis = hb.gauss_legendre(4)
    m, rng = soup(2, 4, 41, with_ldg=True)
    m.state()[:, 2] *= 1 + 0.3*rng.random(m.state()[:, 2].shape)
    ref, work = m.copy(), m.copy()
    h = H.HostHarness(host_emu, m, basis)
    h.invalidate()
    oracle.compute_write_face(basis, ref); h.call("compute_write_face")
    oracle.compute_prolong(basis, ref); h.call("compute_prolong", 0, 0)
    oracle.compute_restrict(basis, ref); h.call("compute_restrict", 1, 0)
    oracle.stabilizing_art_visc(basis, ref, 340.); h.call("stabilizing_art_visc", 340.)
    h.fetch(work)
    assert rel_l2(work.face_state, ref.face_state) <= 1e-13
    assert np.abs(work.uncert - ref.uncert).max() <= 1e-11*np.abs(ref.uncert).max()
    h.close()


def test_adapter_mesh_epoch_detection_emu(oracle, host_emu):
    """a different mesh behind the same (n_dim, row_size) must be re-flattened without an explicit invalidate"""
    basis = hb.gauss_legendre(3)
    for seed, n_def in ((51, 10), (52, 7)):
        m, _ = soup(2, 3, seed, n_def=n_def, with_ldg=True)
        ref, work = m.copy(), m.copy()
        h = H.HostHarness(host_emu, m, basis, seed=seed)
        oracle.compute_euler(basis, ref, dt=1e-4, i_stage=0)
        h.call("compute_euler", dt=1e-4, i_stage=0)
        h.fetch(work)
        assert rel_l2(work.state(), ref.state()) <= 1e-11
        h.lib.hbh_destroy.argtypes = [ctypes.c_void_p]
        # deliberately no h.close() (which invalidates): the next mesh must be detected through its fingerprint
        h.h = None


def test_adapter_errors(host_emu):
    """bad (n_dim, row_size) -> std::runtime_error("demand for invalid kernel") like include/kernel_factory.hpp:114-116"""
    m, _ = soup(2, 3, 61)
    h = H.HostHarness(host_emu, m, hb.gauss_legendre(3))
    with pytest.raises(RuntimeError, match="demand for invalid kernel"):
        h.face_permutation(4, 3, [0, 0, 1, 0], np.zeros(64))
    with pytest.raises(RuntimeError, match="demand for invalid kernel"):
        h.face_permutation(2, 9, [0, 0, 1, 0], np.zeros(64))
    h.close()


@pytest.mark.parametrize("nd", [2, 3])
def test_adapter_face_permutation(oracle, host_emu, nd):
    """hexed::face_permutation(n_dim, row_size, dir, data)->match_faces()/restore() against the oracle for every direction"""
    rs = 5
    m, rng = soup(nd, 2, 62)
    h = H.HostHarness(host_emu, m, hb.gauss_legendre(2))
    for d0 in range(nd):
        for d1 in range(nd):
            for s0 in range(2):
                for s1 in range(2):
                    data = rng.standard_normal((nd + 2)*rs**(nd - 1))
                    a, b = data.copy(), data.copy()
                    oracle.face_permutation(nd, rs, nd + 2, Connection_direction([d0, d1], [s0, s1]), a)
                    h.face_permutation(nd, rs, [d0, d1, s0, s1], b)
                    assert np.array_equal(a, b)
                    h.face_permutation(nd, rs, [d0, d1, s0, s1], b, restore=True)
                    assert np.array_equal(b, data)
    h.close()


# ------------------------------------------------------------------------------------------------ GPU: adapter on the product library
@pytest.mark.gpu
@pytest.mark.parametrize("resident", [False, True])
@pytest.mark.parametrize("nd,rs", [(2, 6), (3, 4), (3, 6)])
def test_adapter_euler_gpu(oracle, host_gpu, nd, rs, resident):
    m, _ = soup(nd, rs, 71, n_car=8, n_def=14, n_ref=6, with_ldg=True)
    out, ref, dts, units = run_euler(oracle, host_gpu, m, hb.gauss_legendre(rs), resident, n_steps=2)
    assert_euler_parity(out, ref, dts)
    check_work_units(m, units, 2)


@pytest.mark.gpu
@pytest.mark.parametrize("nd,rs", [(2, 6), (3, 6)])
def test_adapter_async_boundary_traffic_gpu(oracle, host_gpu, nd, rs):
    """prefetched download of the inside faces + deferred upload of the ghost faces (copy stream, pinned memory) on a B200"""
    basis = hb.gauss_legendre(rs)
    m = M.box_mesh(nd, rs, 6, basis, deformed=True, bc_kind=M.BC_NONPENETRATION)
    density_wave(m, basis)
    oracle.compute_write_face(basis, m)
    host_bcs.lean = True
    try:
        out, ref, dts, _ = run_euler(oracle, host_gpu, m, basis, True, n_steps=3)
    finally:
        host_bcs.lean = False
    assert_euler_parity(out, ref, dts)


@pytest.mark.gpu
@pytest.mark.parametrize("resident", [False, True])
def test_adapter_box_gpu(oracle, host_gpu, resident):
    """the headline mesh class at an oracle-friendly size: 3-D deformed box, row size 6, host-applied freestream ghosts"""
    basis = hb.gauss_legendre(6)
    m = M.box_mesh(3, 6, 5, basis, deformed=True, bc_kind=M.BC_FREESTREAM, bc_params=freestream_state(3))
    density_wave(m, basis)
    oracle.compute_write_face(basis, m)
    out, ref, dts, _ = run_euler(oracle, host_gpu, m, basis, resident, n_steps=3)
    assert_euler_parity(out, ref, dts)


@pytest.mark.gpu
@pytest.mark.parametrize("resident", [False, True])
@pytest.mark.parametrize("pde", [NAVIER_STOKES, ADVECTION, SMOOTH_AV, FIX_THERM_ADMIS])
def test_adapter_other_pdes_gpu(oracle, host_gpu, pde, resident):
    m, rng = soup(3, 4, 80 + pde, with_ldg=True, with_wide=True)
    prepare_pde_state(m, rng, pde)
    out, ref, dts = run_pde(oracle, host_gpu, m, hb.gauss_legendre(4), pde, resident)
    assert_pde_parity(out, ref, dts)


# ------------------------------------------------------------------------------------------------ B200s: one Kernel_mesh on several GPUs through the adapter
def _gpu_devices(n):
    import ctypes
    from hexed_b200.kernels import load_library
    count = ctypes.c_int(0)
    load_library().hexed_b200_device_count(ctypes.byref(count))
    if count.value < n:
        pytest.skip("needs %d CUDA devices, this box has %d" % (n, count.value))
    return list(range(n))


@pytest.mark.gpu
@pytest.mark.parametrize("resident", [False, True])
@pytest.mark.parametrize("nd,rs,n_dev", [(2, 6, 2), (3, 4, 2), (3, 6, 2), (3, 6, 4), (3, 6, 8)])
def test_adapter_multi_device_euler_gpu(oracle, host_gpu, nd, rs, n_dev, resident):
    """hexed::compute_euler / max_dt_euler of ONE Kernel_mesh on n_dev B200s: NCCL send/recv of the cut faces + ncclAllReduce(min),
    hanging faces of every stretch split across devices, against the CPU checker on the undivided mesh"""
    devices = _gpu_devices(n_dev)
    m, _ = soup(nd, rs, 71, n_car=16, n_def=40, n_ref=8, with_ldg=True)
    oracle.compute_prolong(hb.gauss_legendre(rs), m)
    out, ref, dts, units = run_euler(oracle, host_gpu, m, hb.gauss_legendre(rs), resident, n_steps=1, devices=devices)
    assert np.isfinite(ref.state()).all()
    assert_euler_parity(out, ref, dts)
    check_work_units(m, units, 1)
    assert len(set(run_euler.last_owners.tolist())) == n_dev
    assert "NCCL" in run_euler.last_transport


@pytest.mark.gpu
@pytest.mark.parametrize("resident", [False, True])
@pytest.mark.parametrize("n_dev", [2, 8])
def test_adapter_multi_device_box_gpu(oracle, host_gpu, n_dev, resident):
    devices = _gpu_devices(n_dev)
    basis = hb.gauss_legendre(6)
    m = M.box_mesh(3, 6, 8, basis, deformed=True, bc_kind=M.BC_FREESTREAM, bc_params=freestream_state(3))
    density_wave(m, basis)
    oracle.compute_write_face(basis, m)
    out, ref, dts, _ = run_euler(oracle, host_gpu, m, basis, resident, n_steps=2, devices=devices, coords=m.elem_index)
    assert_euler_parity(out, ref, dts)


@pytest.mark.gpu
@pytest.mark.parametrize("resident", [False, True])
def test_adapter_multi_device_refined_box_gpu(oracle, host_gpu, resident):
    """C5 class on hardware: Cartesian hanging-node faces cut by the device split"""
    devices = _gpu_devices(2)
    basis = hb.gauss_legendre(6)
    refine = np.zeros((4, 4, 4), bool); refine[1, 1, 1] = refine[2, 2, 1] = refine[3, 0, 2] = True
    m = M.refined_box_mesh(3, 6, 4, basis, refine, bc_kind=M.BC_COPY)
    M.random_flow_state(m, np.random.default_rng(5), mach=0.2)
    oracle.compute_write_face(basis, m); oracle.compute_prolong(basis, m)
    out, ref, dts, _ = run_euler(oracle, host_gpu, m, basis, resident, n_steps=2, devices=devices)
    assert_euler_parity(out, ref, dts)


@pytest.mark.gpu
@pytest.mark.parametrize("resident", [False, True])
@pytest.mark.parametrize("n_dev", [2, 4])
def test_adapter_multi_device_navier_stokes_gpu(oracle, host_gpu, n_dev, resident):
    devices = _gpu_devices(n_dev)
    m, rng = soup(3, 4, 72, n_car=12, n_def=30, n_ref=6, with_ldg=True, with_wide=True)
    prepare_pde_state(m, rng, NAVIER_STOKES)
    oracle.compute_prolong(hb.gauss_legendre(4), m)
    out, ref, dts = run_pde(oracle, host_gpu, m, hb.gauss_legendre(4), NAVIER_STOKES, resident, devices=devices)
    assert_pde_parity(out, ref, dts)
