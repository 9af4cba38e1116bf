"""Parity of the CUDA path (through the C ABI) against the CPU oracle on identical inputs. Run with -m gpu on a B200.
Tolerances are the north-star's: state rel-L2 <= 1e-11, max_dt <= 1e-13 relative."""
import numpy as np
import pytest

import hexed_b200 as hb
from hexed_b200 import mesh as M
from hexed_b200.kernels import Device
from util import (run_euler_pair, assert_euler_parity, density_wave, freestream_state, rel_l2, run_pde_pair, assert_pde_parity,
                  prepare_pde_state, NAVIER_STOKES, ADVECTION, SMOOTH_AV, FIX_THERM_ADMIS)

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("nd,rs", [(1, 2), (1, 6), (2, 2), (2, 3), (2, 6), (2, 8), (3, 2), (3, 3), (3, 4), (3, 5), (3, 6), (3, 7), (3, 8)])
def test_soup_all_orientations(oracle, gpu_lib, nd, rs):
    """every connection direction, hanging faces with every stretch flag, car/def mix, unit-normal fallback, copy BCs"""
    rng = np.random.default_rng(406)
    basis = hb.gauss_legendre(rs)
    m = M.soup_mesh(nd, rs, rng, n_car=8, n_def=14, n_ref=6, with_ldg=False)
    M.random_flow_state(m, rng)
    out, ref, dts, launches = run_euler_pair(oracle, gpu_lib, m, basis, n_steps=2)
    assert launches > 0
    assert_euler_parity(out, ref, dts)


@pytest.mark.parametrize("local_time,use_filter", [(True, False), (False, True), (True, True)])
def test_soup_options(oracle, gpu_lib, local_time, use_filter):
    rng = np.random.default_rng(11)
    basis = hb.gauss_legendre(6)
    m = M.soup_mesh(3, 6, rng, with_ldg=False)
    M.random_flow_state(m, rng)
    out, ref, dts, _ = run_euler_pair(oracle, gpu_lib, m, basis, n_steps=2, local_time=local_time, use_filter=use_filter, safety=0.05)
    assert_euler_parity(out, ref, dts)


@pytest.mark.parametrize("nd,rs,n,deformed", [(2, 6, 16, False), (2, 6, 8, True), (3, 6, 5, False), (3, 6, 5, True), (3, 4, 6, True),
                                              (2, 6, 101, False), (2, 6, 101, True), (2, 4, 150, True), (2, 8, 70, False)])
def test_box(oracle, gpu_lib, nd, rs, n, deformed):
    """BASELINE configs at oracle-friendly sizes: 2-D vortex-like Cartesian box (16x16, row size 6) and the 3-D row-size-6 box"""
    basis = hb.gauss_legendre(rs)
    m = M.box_mesh(nd, rs, n, basis, deformed=deformed, bc_kind=M.BC_FREESTREAM, bc_params=freestream_state(nd))
    density_wave(m, basis)
    oracle.compute_write_face(basis, m)
    out, ref, dts, _ = run_euler_pair(oracle, gpu_lib, m, basis, n_steps=5)
    assert_euler_parity(out, ref, dts)


def test_compute_residual_mode(oracle, gpu_lib):
    from pyoracle import EULER
    basis = hb.gauss_legendre(6)
    m = M.box_mesh(3, 6, 3, basis, deformed=True, bc_kind=M.BC_COPY)
    density_wave(m, basis)
    oracle.compute_write_face(basis, m)
    ref = m.copy()
    dev = Device(3, 6, basis, lib_path=gpu_lib).load_mesh(m)
    oracle.apply_state_bcs(ref); oracle.compute_euler(basis, ref, dt=1., i_stage=0, compute_residual=True)
    dev.apply_state_bcs(); dev.compute_euler(dt=1., i_stage=0, compute_residual=True)
    dev.sync_to_host(m)
    assert rel_l2(m.cache(), ref.cache()) <= 1e-11
    assert np.array_equal(m.state(), ref.state())  # the state itself must be untouched
    with pytest.raises(RuntimeError):
        dev.compute_euler(dt=1., i_stage=1, compute_residual=True)  # reference include/Spatial.hpp:323
    dev.close()


def test_write_face_prolong_restrict_standalone(oracle, gpu_lib):
    rng = np.random.default_rng(3)
    basis = hb.gauss_legendre(5)
    for nd in (2, 3):
        m = M.soup_mesh(nd, 5, rng, n_ref=8, with_ldg=True)
        M.random_flow_state(m, rng)
        ref = m.copy()
        dev = Device(nd, 5, basis, lib_path=gpu_lib).load_mesh(m)
        oracle.compute_write_face(basis, ref); dev.compute_write_face()
        for scale, offset in [(False, False), (True, False), (True, True), (False, True)]:
            oracle.compute_prolong(basis, ref, scale, offset); dev.compute_prolong(scale, offset)
            oracle.compute_restrict(basis, ref, scale, offset); dev.compute_restrict(scale, offset)
        dev.sync_to_host(m)
        assert rel_l2(m.face_state, ref.face_state) <= 1e-13
        assert rel_l2(m.face_ldg, ref.face_ldg) <= 1e-13
        dev.close()


def test_face_permutation_bit_exact(oracle, gpu_lib):
    """the device permutation tables and the in-place oracle permutation must agree exactly, for every direction"""
    from hexed_b200.tables import face_permutation
    rng = np.random.default_rng(5)
    for nd, rs in [(2, 6), (3, 4), (3, 6)]:
        basis = hb.gauss_legendre(rs)
        dev = Device(nd, rs, basis, lib_path=gpu_lib)
        m = M.FlatMesh(nd, rs, 1, 0); dev.load_mesh(m)
        for direction in M.all_directions(nd):
            data = rng.standard_normal((nd + 2)*rs**(nd - 1))
            a, b = data.copy(), data.copy()
            oracle.face_permutation(nd, rs, nd + 2, direction, a)
            dev.face_permutation(direction, b)
            assert np.array_equal(a, b)
            assert np.array_equal(dev.face_permutation_table(direction), face_permutation(nd, rs, direction))
            oracle.face_permutation(nd, rs, nd + 2, direction, a, restore=True)
            dev.face_permutation(direction, b, restore=True)
            assert np.array_equal(a, data) and np.array_equal(b, data)
        dev.close()


def test_invalid_kernel_raises(gpu_lib):
    with pytest.raises(RuntimeError, match="demand for invalid kernel"):
        Device(3, 9, None, lib_path=gpu_lib)
    with pytest.raises(RuntimeError, match="demand for invalid kernel"):
        Device(4, 6, None, lib_path=gpu_lib)


def test_full_size_properties(gpu_lib):
    """size-independent checks at a large size the oracle cannot do in seconds (3-D deformed, row size 6, 48^3 = 110k elements):
    a uniform freestream is preserved by a full step on the warped mesh (metric identities), mass and energy are conserved
    to round-off for a non-trivial state with copy ghosts, and max_dt equals the minimum of the local time-step scale."""
    import torch
    nd, rs, n = 3, 6, 48
    basis = hb.gauss_legendre(rs)
    fs = freestream_state(nd)
    m = M.box_mesh(nd, rs, n, basis, deformed=True, bc_kind=M.BC_FREESTREAM, bc_params=fs, device="cuda", geometry_chunk=16384)
    st = m.state()
    st[:] = fs[None, :, None]
    dev = Device(nd, rs, basis, lib_path=gpu_lib).load_mesh(m)
    dev.set_option(1, 1)  # HEXED_B200_OPT_CFL_CACHE on for the screen check below
    dev.compute_write_face()
    dt = dev.max_dt_euler(0.7, 0.7, False)
    assert dt > 0
    for stage in (0, 1):
        dev.apply_state_bcs(); dev.compute_euler(dt=dt, i_stage=stage)
    out = np.empty_like(st)
    dev.download_elements(out, 0, nd + 2)
    assert rel_l2(out, st) <= 1e-11  # freestream preservation
    # the CFL screen left by the stage-1 Local kernel: in a uniform flow every element ties, so every element is re-evaluated;
    # the result must be the very number the full max_dt kernel computes
    n0 = dev.launch_count()
    dt_screen = dev.max_dt_euler(0.7, 0.7, False)
    assert dev.launch_count() - n0 == 2
    dev.set_option(1, 0); dt_full = dev.max_dt_euler(0.7, 0.7, False); dev.set_option(1, 1)
    assert dt_screen == dt_full
    # max_dt vs local tss: global dt == min over qpoints of the local scale
    dev.upload_elements(np.ascontiguousarray(st), 0, nd + 2)
    dev.max_dt_euler(0.7, 0.7, True)
    tss = np.empty((m.n_elem, 1, m.nq)); dev.download_elements(tss, nd + 2, 1)
    assert abs(tss.min()/dt - 1) <= 1e-13
    dev.close()


@pytest.mark.parametrize("pde", [NAVIER_STOKES, ADVECTION, SMOOTH_AV, FIX_THERM_ADMIS])
@pytest.mark.parametrize("nd,rs", [(1, 4), (2, 2), (2, 6), (3, 3), (3, 6), (3, 8)])
def test_soup_other_pdes(oracle, gpu_lib, pde, nd, rs):
    """compute_navier_stokes / compute_advection / compute_smooth_av / compute_fix_therm_admis and their max_dt_* on the
    structurally complete soup mesh (every orientation, hanging faces, car/def mix); 2 steps, with and without the modal filter"""
    rng = np.random.default_rng(100*pde + 10*nd + rs)
    basis = hb.gauss_legendre(rs)
    m = M.soup_mesh(nd, rs, rng, with_ldg=True, with_wide=True)
    M.random_flow_state(m, rng)
    prepare_pde_state(m, rng, pde)
    out, ref, dts = run_pde_pair(oracle, gpu_lib, m, basis, pde, n_steps=2, use_filter=(rs % 2 == 0), safety=0.1)
    assert_pde_parity(out, ref, dts)


@pytest.mark.parametrize("pde", [NAVIER_STOKES, FIX_THERM_ADMIS])
def test_other_pdes_residual_mode_and_local_time(oracle, gpu_lib, pde):
    rng = np.random.default_rng(77)
    basis = hb.gauss_legendre(6)
    m = M.soup_mesh(3, 6, rng, with_ldg=True)
    M.random_flow_state(m, rng)
    prepare_pde_state(m, rng, pde)
    out, ref, dts = run_pde_pair(oracle, gpu_lib, m, basis, pde, n_steps=1, compute_residual=True, safety=0.1)
    assert_pde_parity(out, ref, dts)
    out, ref, dts = run_pde_pair(oracle, gpu_lib, m, basis, pde, n_steps=1, local_time=True, safety=0.05)
    assert_pde_parity(out, ref, dts)


@pytest.mark.parametrize("nd,n,deformed", [(2, 8, True), (2, 8, False), (3, 4, True), (3, 4, False)])
def test_box_navier_stokes(oracle, gpu_lib, nd, n, deformed):
    """BASELINE config samples/cylinder in miniature: row size 6 viscous flow with Sutherland viscosity on a box, 3 steps"""
    basis = hb.gauss_legendre(6)
    m = M.box_mesh(nd, 6, n, basis, deformed=deformed, bc_kind=M.BC_FREESTREAM, bc_params=freestream_state(nd), with_ldg=True)
    density_wave(m, basis)
    oracle.compute_write_face(basis, m)
    out, ref, dts = run_pde_pair(oracle, gpu_lib, m, basis, NAVIER_STOKES, n_steps=3, safety=0.5)
    assert_pde_parity(out, ref, dts)


@pytest.mark.parametrize("rs", [4, 6])
def test_navier_stokes_3d_line_kernel(oracle, gpu_lib, rs):
    """3-D row size 4 / 6 without the modal filter takes ns_local_line_kernel; the generic point-per-thread kernel (option off)
    must give the same answer to round-off"""
    rng = np.random.default_rng(91)
    basis = hb.gauss_legendre(rs)
    m = M.soup_mesh(3, rs, rng, n_car=8, n_def=14, n_ref=6, with_ldg=True)
    M.random_flow_state(m, rng)
    prepare_pde_state(m, rng, NAVIER_STOKES)
    out, ref, dts = run_pde_pair(oracle, gpu_lib, m, basis, NAVIER_STOKES, n_steps=2, safety=0.1)
    assert_pde_parity(out, ref, dts)


@pytest.mark.parametrize("options", [(), ((1, 1), (2, 1))], ids=["default", "cfl_cache+fused_admis"])
@pytest.mark.parametrize("rs", [6, 4])
def test_large_mixed_soup_every_persistent_cta_loops(oracle, gpu_lib, rs, options):
    """The persistent pipelined Local kernels run 3 CTAs on each of 148 SMs (444 CTAs): with 1 900 Cartesian + 1 900 deformed elements in
    ONE mesh every CTA of both launches goes through its double-buffer wrap-around >= 4 times, the deformed launch starts at
    elem_begin = n_car > 0, 300 hanging-node faces (every stretch flag) sit in between, and the result is compared with the oracle
    element by element -- with the defaults and with the CFL cache and the fused admissibility bits switched on."""
    rng = np.random.default_rng(5)
    basis = hb.gauss_legendre(rs)
    m = M.soup_mesh(3, rs, rng, n_car=1900, n_def=1900, n_ref=300, with_ldg=False)
    M.random_flow_state(m, rng)
    out, ref, dts, _ = run_euler_pair(oracle, gpu_lib, m, basis, n_steps=2, safety=0.05, options=options)
    assert_euler_parity(out, ref, dts)
    per_elem = np.linalg.norm((out.state() - ref.state()).reshape(m.n_elem, -1), axis=1)/np.linalg.norm(ref.state().reshape(m.n_elem, -1), axis=1)
    assert per_elem.max() <= 1e-11, int(per_elem.argmax())  # no single element may hide behind the mesh-wide norm


@pytest.mark.parametrize("deformed", [True, False])
def test_box_3d_every_persistent_cta_loops(oracle, gpu_lib, deformed):
    """14^3 = 2 744 elements of the headline shape (3-D, row size 6): >= 6 iterations per persistent CTA, oracle-compared"""
    basis = hb.gauss_legendre(6)
    m = M.box_mesh(3, 6, 14, basis, deformed=deformed, bc_kind=M.BC_FREESTREAM, bc_params=freestream_state(3))
    density_wave(m, basis)
    oracle.compute_write_face(basis, m)
    out, ref, dts, _ = run_euler_pair(oracle, gpu_lib, m, basis, n_steps=3)
    assert_euler_parity(out, ref, dts)


def test_navier_stokes_3d_padded_layout_is_bit_identical(port_oracle, gpu_lib):
    """HEXED_B200_OPT_NS_LOCAL_LAYOUT 1 (ns_local_pad_kernel, default) against 0 (ns_local_line_kernel): same operations in the same
    order, so bit-identical state, faces and residual cache on 500 + 500 elements with hanging faces"""
    rng = np.random.default_rng(93)
    basis = hb.gauss_legendre(6)
    m = M.soup_mesh(3, 6, rng, n_car=500, n_def=500, n_ref=100, with_ldg=True)
    M.random_flow_state(m, rng)
    prepare_pde_state(m, rng, NAVIER_STOKES)
    a, ref, dts = run_pde_pair(port_oracle, gpu_lib, m, basis, NAVIER_STOKES, n_steps=2, safety=0.1, options=((3, 0),))
    assert_pde_parity(a, ref, dts)
    for layout in (1, 2):
        b, _, _ = run_pde_pair(port_oracle, gpu_lib, m, basis, NAVIER_STOKES, n_steps=2, safety=0.1, options=((3, layout),))
        assert np.array_equal(a.elem_data, b.elem_data) and np.array_equal(a.face_ldg, b.face_ldg) and np.array_equal(a.face_state, b.face_state), layout


def test_navier_stokes_3d_large_mixed_soup(oracle, gpu_lib):
    """ns_local_line_kernel / ns_reconcile_bulk_kernel on 1 000 + 1 000 elements with 200 hanging faces, element by element"""
    rng = np.random.default_rng(92)
    basis = hb.gauss_legendre(6)
    m = M.soup_mesh(3, 6, rng, n_car=1000, n_def=1000, n_ref=200, with_ldg=True)
    M.random_flow_state(m, rng)
    prepare_pde_state(m, rng, NAVIER_STOKES)
    out, ref, dts = run_pde_pair(oracle, gpu_lib, m, basis, NAVIER_STOKES, n_steps=2, safety=0.1)
    assert_pde_parity(out, ref, dts)
    per_elem = np.linalg.norm((out.state() - ref.state()).reshape(m.n_elem, -1), axis=1)/np.linalg.norm(ref.state().reshape(m.n_elem, -1), axis=1)
    assert per_elem.max() <= 1e-11, int(per_elem.argmax())


@pytest.mark.parametrize("rs,n", [(4, 5), (6, 4)])
def test_cfl_cache_follows_the_state(oracle, gpu_lib, rs, n):
    from util import check_cfl_cache
    check_cfl_cache(oracle, gpu_lib, rs, n)


def test_stabilizing_art_visc(oracle, gpu_lib):
    rng = np.random.default_rng(8)
    for nd, rs in [(2, 6), (3, 6), (3, 4)]:
        basis = hb.gauss_legendre(rs)
        m = M.soup_mesh(nd, rs, rng, with_ldg=False)
        M.random_flow_state(m, rng)
        m.state()[:, nd] *= 1 + 0.3*rng.random(m.state()[:, nd].shape)
        ref = m.copy()
        dev = Device(nd, rs, basis, lib_path=gpu_lib).load_mesh(m)
        oracle.stabilizing_art_visc(basis, ref, 340.)
        dev.stabilizing_art_visc(340.)
        dev.sync_to_host(m)
        assert np.abs(m.uncert - ref.uncert).max() <= 1e-11*np.abs(ref.uncert).max()
        dev.close()


@pytest.mark.parametrize("nd,rs", [(2, 6), (3, 6), (3, 4)])
def test_every_device_boundary_condition(oracle, gpu_lib, nd, rs):
    """Freestream, Copy, Nonpenetration, Outflow, Pressure_outflow, No_slip (three thermal kinds) on the device,
    as ghost-state and flux conditions of viscous steps"""
    from util import mixed_bcs
    rng = np.random.default_rng(5)
    basis = hb.gauss_legendre(rs)
    m = M.soup_mesh(nd, rs, rng, n_car=10, n_def=24, n_ref=4, with_ldg=True)
    M.random_flow_state(m, rng)
    prepare_pde_state(m, rng, NAVIER_STOKES)
    assert m.bcs[0]["ghost_slot"].size >= 8
    mixed_bcs(m, rng)
    out, ref, dts = run_pde_pair(oracle, gpu_lib, m, basis, NAVIER_STOKES, n_steps=2, safety=0.1)
    assert_pde_parity(out, ref, dts)


@pytest.mark.gpu
@pytest.mark.parametrize("nd,rs", [(1, 6), (2, 6), (3, 6), (3, 3)])
def test_riemann_invariants_bc(oracle, gpu_lib, nd, rs):
    """Riemann_invariants (the `characteristic` condition of every sample case) state + flux against the oracle's restatement of
    Characteristics / ColPivHouseholderQR, over sub/supersonic in/outflow and singular (zero-pressure) points"""
    from util import check_riemann_bc
    check_riemann_bc(oracle, gpu_lib, nd, rs)


@pytest.mark.parametrize("nd,rs", [(1, 6), (2, 6), (3, 6), (3, 4)])
def test_is_admissible(oracle, gpu_lib, nd, rs):
    """Solver::is_admissible / Element::record on the device"""
    from util import check_admissibility
    check_admissibility(oracle, gpu_lib, nd, rs)


@pytest.mark.parametrize("nd,rs", [(2, 6), (3, 6), (3, 4), (3, 8)])
def test_set_jacobian(gpu_lib, nd, rs):
    """metric terms of deformed elements computed on the device from vertex positions and node adjustments"""
    from util import check_set_jacobian
    check_set_jacobian(gpu_lib, nd, rs)


@pytest.mark.parametrize("nd,rs", [(2, 6), (3, 6), (3, 4)])
def test_av_glue(gpu_lib, nd, rs):
    """pointwise loops of the artificial-viscosity pipelines on the device"""
    from util import check_av_glue
    check_av_glue(gpu_lib, nd, rs)


@pytest.mark.parametrize("nd,rs", [(2, 6), (3, 6), (3, 8)])
def test_aux_bcs(gpu_lib, nd, rs):
    from util import check_aux_bcs
    check_aux_bcs(gpu_lib, nd, rs)


@pytest.mark.parametrize("nd,rs", [(2, 6), (3, 6), (3, 4)])
def test_av_smoothness_pipeline(oracle, gpu_lib, nd, rs):
    """the naca0012-class configuration's shock-capturing update, Solver::update_art_visc_smoothness, with no host loop left"""
    from util import check_av_pipeline
    check_av_pipeline(oracle, gpu_lib, nd, rs)


@pytest.mark.parametrize("nd,rs", [(2, 6), (3, 6)])
def test_vertex_sharing_and_fix_admis_spread(oracle, gpu_lib, nd, rs):
    """share_vertex_data + the spreading step of fix_admissibility on the device"""
    from util import check_vertex_sharing
    check_vertex_sharing(oracle, gpu_lib, nd, rs)


@pytest.mark.parametrize("nd,rs", [(2, 6), (3, 6)])
def test_set_art_visc_admis(oracle, gpu_lib, nd, rs):
    from util import check_set_art_visc_admis
    check_set_art_visc_admis(oracle, gpu_lib, nd, rs)


@pytest.mark.parametrize("nd,rs", [(2, 6), (3, 6), (3, 3)])
def test_av_elwise(gpu_lib, nd, rs):
    from util import check_av_elwise
    check_av_elwise(gpu_lib, nd, rs)


@pytest.mark.parametrize("nd,rs", [(2, 6), (3, 6), (3, 4)])
def test_shared_normals_soup(oracle, gpu_lib, nd, rs):
    """connection passes of Solver::calc_jacobian on the device"""
    from util import check_shared_normals_soup
    check_shared_normals_soup(oracle, gpu_lib, nd, rs)


@pytest.mark.parametrize("nd,rs,n", [(2, 6, 12), (3, 6, 6), (3, 4, 8)])
def test_calc_jacobian_box(gpu_lib, nd, rs, n):
    from util import check_calc_jacobian_box
    check_calc_jacobian_box(gpu_lib, nd, rs, n)


@pytest.mark.parametrize("rs,n", [(4, 40), (6, 37), (8, 20)])
@pytest.mark.parametrize("deformed", [False, True])
def test_navier_stokes_2d_line_kernel(oracle, gpu_lib, deformed, rs, n):
    """the cylinder-class kernel: 2-D batched line-task Navier-Stokes Local (partial last batch: n^2 is not a multiple of the batch)"""
    rng = np.random.default_rng(79)
    basis = hb.gauss_legendre(rs)
    m = M.box_mesh(2, rs, n, basis, deformed=deformed, bc_kind=M.BC_NONPENETRATION, with_ldg=True)
    density_wave(m, basis)
    oracle.compute_write_face(basis, m)
    prepare_pde_state(m, rng, NAVIER_STOKES)
    out, ref, dts = run_pde_pair(oracle, gpu_lib, m, basis, NAVIER_STOKES, n_steps=2, safety=0.1)
    assert_pde_parity(out, ref, dts)
    out, ref, dts = run_pde_pair(oracle, gpu_lib, m, basis, NAVIER_STOKES, n_steps=1, compute_residual=True)
    assert_pde_parity(out, ref, dts)


@pytest.mark.parametrize("nd,rs,n", [(2, 6, 40), (3, 6, 8), (3, 4, 9)])
def test_fused_admissibility(oracle, gpu_lib, nd, rs, n):
    from util import check_fused_admissibility
    check_fused_admissibility(oracle, gpu_lib, nd, rs, n)


@pytest.mark.parametrize("mode,deformed", [(0, True), (2, True), (3, False), (4, True), (4, False)])
@pytest.mark.parametrize("rs,n", [(6, 7), (4, 9)])
def test_box_3d_other_local_kernels(oracle, gpu_lib, rs, n, mode, deformed):
    """HEXED_B200_OPT_PIPELINED_LOCAL = 0 (the general Local kernel) and 2 (the pipelined kernel with its earlier, fully staged
    shared-memory layout) stay available for A/B measurements, 3 = Cartesian elements in the lean layout with four resident CTAs:
    same parity bar as the default kernels"""
    basis = hb.gauss_legendre(rs)
    m = M.box_mesh(3, rs, n, basis, deformed=deformed, bc_kind=M.BC_FREESTREAM, bc_params=freestream_state(3))
    density_wave(m, basis)
    oracle.compute_write_face(basis, m)
    out, ref, dts, _ = run_euler_pair(oracle, gpu_lib, m, basis, n_steps=2, options=((0, mode),))
    assert_euler_parity(out, ref, dts)


@pytest.mark.parametrize("nd,rs,n", [(2, 6, 60), (3, 6, 16), (3, 4, 20)])
def test_max_dt_running_screen_random_states(oracle, gpu_lib, nd, rs, n):
    from util import check_max_dt_running_screen_random
    check_max_dt_running_screen_random(oracle, gpu_lib, nd, rs, n, range(40))


@pytest.mark.parametrize("nd,rs,n", [(2, 6, 60), (3, 6, 16), (3, 4, 20)])
def test_max_dt_running_screen_navier_stokes(oracle, gpu_lib, nd, rs, n):
    from util import check_max_dt_running_screen_ns
    check_max_dt_running_screen_ns(oracle, gpu_lib, nd, rs, n, range(40))


@pytest.mark.parametrize("nd,rs,n", [(2, 6, 150), (3, 6, 40), (3, 3, 30)])
def test_max_dt_running_screen_is_exact(oracle, gpu_lib, nd, rs, n):
    from util import check_max_dt_running_screen
    check_max_dt_running_screen(oracle, gpu_lib, nd, rs, n)


@pytest.mark.gpu
@pytest.mark.parametrize("nd,rs,pde", [(2, 4, "euler"), (3, 3, "euler"), (2, 4, "navier_stokes")])
def test_time_step_scale_write_skipped_only_when_known_one(oracle, gpu_lib, nd, rs, pde):
    from util import check_tss_write_skipped
    check_tss_write_skipped(oracle, gpu_lib, nd, rs, pde)


@pytest.mark.parametrize("use_graph", [False, True])
@pytest.mark.parametrize("nd,rs,n,deformed", [(2, 6, 16, False), (2, 6, 24, True), (3, 6, 6, True), (3, 5, 4, False)])
def test_update_euler_device_time_step(oracle, gpu_lib, nd, rs, n, deformed, use_graph):
    """time step on the device + CUDA graph replay of the step: bit-identical to the call-by-call sequence"""
    from util import check_update_euler
    check_update_euler(oracle, gpu_lib, nd, rs, n, n_steps=25, use_graph=use_graph, deformed=deformed)


@pytest.mark.parametrize("nd,rs,n", [(2, 6, 9), (3, 6, 4)])
def test_update_euler_refined_mesh(oracle, gpu_lib, nd, rs, n):
    """the graph-replayed step on a mesh with hanging-node faces"""
    from util import check_update_euler
    check_update_euler(oracle, gpu_lib, nd, rs, n, n_steps=20, use_graph=True, refined=True)


@pytest.mark.parametrize("use_graph", [False, True])
@pytest.mark.parametrize("nd,rs,n", [(2, 6, 16), (3, 6, 5), (2, 5, 9)])
def test_update_navier_stokes_device_time_step(oracle, gpu_lib, nd, rs, n, use_graph):
    """the viscous loop with the time step on the device + CUDA graph replay: bit-identical to the call-by-call sequence"""
    from util import check_update_navier_stokes
    check_update_navier_stokes(oracle, gpu_lib, nd, rs, n, n_steps=12, use_graph=use_graph)


@pytest.mark.parametrize("n_cheby,n_steps", [(3, 8), (4, 9)])
def test_update_loops_with_chebyshev_steps(oracle, gpu_lib, n_cheby, n_steps):
    """n_cheby_flow > 1 (the shock-capturing cases): Chebyshev factors cycle, one graph per cycle, a partial cycle at the end"""
    from util import check_update_euler, check_update_navier_stokes
    check_update_euler(oracle, gpu_lib, 2, 4, 4, n_steps=n_steps, use_graph=True, deformed=True, n_cheby=n_cheby)
    check_update_navier_stokes(oracle, gpu_lib, 2, 4, 3, n_steps=n_steps, use_graph=True, n_cheby=n_cheby)
