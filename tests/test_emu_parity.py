"""CPU-side check of the CUDA kernels' logic: the product's .cu sources compiled for host threads
(tests/emu/cuda_emu.hpp) against the oracle. No GPU needed; small meshes only. The real parity tests are in
test_gpu_parity.py (-m gpu); this file exists so indexing bugs are caught in the GPU-less container."""
import numpy as np
import pytest

import hexed_b200 as hb
from hexed_b200 import mesh as M
from util import (run_euler_pair, assert_euler_parity, density_wave, freestream_state, run_pde_pair, assert_pde_parity,
                  prepare_pde_state, NAVIER_STOKES, ADVECTION, SMOOTH_AV, FIX_THERM_ADMIS)


@pytest.mark.parametrize("nd,rs", [(1, 3), (2, 2), (2, 5), (3, 2), (3, 3)])
def test_soup_all_orientations(oracle, emu_lib, nd, rs):
    rng = np.random.default_rng(406)
    basis = hb.gauss_legendre(rs)
    m = M.soup_mesh(nd, rs, rng, with_ldg=False)
    M.random_flow_state(m, rng)
    out, ref, dts, _ = run_euler_pair(oracle, emu_lib, m, basis, n_steps=1)
    assert_euler_parity(out, ref, dts)


def test_soup_local_time_and_filter(oracle, emu_lib):
    rng = np.random.default_rng(7)
    basis = hb.gauss_legendre(4)
    m = M.soup_mesh(2, 4, rng, with_ldg=False)
    M.random_flow_state(m, rng)
    out, ref, dts, _ = run_euler_pair(oracle, emu_lib, m, basis, n_steps=1, local_time=True, use_filter=True)
    assert_euler_parity(out, ref, dts)


@pytest.mark.parametrize("deformed", [False, True])
def test_box_2d(oracle, emu_lib, deformed):
    basis = hb.gauss_legendre(4)
    m = M.box_mesh(2, 4, 3, basis, deformed=deformed, bc_kind=M.BC_FREESTREAM, bc_params=freestream_state(2))
    density_wave(m, basis)
    oracle.compute_write_face(basis, m)
    out, ref, dts, _ = run_euler_pair(oracle, emu_lib, m, basis, n_steps=1)
    assert_euler_parity(out, ref, dts)


@pytest.mark.parametrize("pde", [NAVIER_STOKES, ADVECTION, SMOOTH_AV, FIX_THERM_ADMIS])
@pytest.mark.parametrize("nd,rs", [(1, 3), (2, 3), (3, 2)])
def test_soup_other_pdes(oracle, emu_lib, pde, nd, rs):
    rng = np.random.default_rng(100*pde + 10*nd + rs)
    basis = hb.gauss_legendre(rs)
    m = M.soup_mesh(nd, rs, rng, with_ldg=True, with_wide=True)
    M.random_flow_state(m, rng)
    prepare_pde_state(m, rng, pde)
    out, ref, dts = run_pde_pair(oracle, emu_lib, m, basis, pde, n_steps=1, use_filter=(rs == 3))
    assert_pde_parity(out, ref, dts)


def test_stab_art_visc(oracle, emu_lib):
    from hexed_b200.kernels import Device
    rng = np.random.default_rng(8)
    basis = hb.gauss_legendre(4)
    m = M.soup_mesh(2, 4, rng, with_ldg=False)
    M.random_flow_state(m, rng)
    m.state()[:, 2] *= 1 + 0.3*rng.random(m.state()[:, 2].shape)  # rough density so the indicator lands on the ramp for some elements
    ref = m.copy()
    dev = Device(2, 4, basis, lib_path=emu_lib).load_mesh(m)
    oracle.stabilizing_art_visc(basis, ref, 340.)
    dev.stabilizing_art_visc(340.)
    dev.sync_to_host(m)
    assert np.abs(m.uncert - ref.uncert).max() <= 1e-11*np.abs(ref.uncert).max()
    dev.close()


@pytest.mark.parametrize("rs", [4, 6])
@pytest.mark.parametrize("deformed", [False, True])
def test_box_3d_pipelined_local(oracle, emu_lib, deformed, rs):
    """3-D row size 4 / 6 takes the persistent TMA-pipelined Local kernel (local_euler_pipe.cu); the emulated device has one SM with
    two resident CTAs, so every CTA walks several elements and both stage buffers and all barrier phases are exercised. Row size 6
    uses the bank-conflict-avoiding line map (LineMap<6>: sparse dimension-1 lanes, 16-byte dimension-2 accesses)"""
    basis = hb.gauss_legendre(rs)
    m = M.box_mesh(3, rs, 2, basis, deformed=deformed, bc_kind=M.BC_FREESTREAM, bc_params=freestream_state(3))
    density_wave(m, basis)
    oracle.compute_write_face(basis, m)
    out, ref, dts, _ = run_euler_pair(oracle, emu_lib, m, basis, n_steps=2)
    assert_euler_parity(out, ref, dts)


@pytest.mark.parametrize("rs", [4, 6])
@pytest.mark.parametrize("kind", ["soup", "box_car", "box_def"])
def test_navier_stokes_3d_line_kernel(oracle, emu_lib, kind, rs):
    """3-D row size 4 / 6 Navier-Stokes takes the line-task Local kernel (ns_local_line_kernel in generic_part.cu)"""
    rng = np.random.default_rng(77)
    basis = hb.gauss_legendre(rs)
    if kind == "soup":
        m = M.soup_mesh(3, rs, rng, n_car=3, n_def=5, n_ref=2, with_ldg=True)
        M.random_flow_state(m, rng)
    else:
        m = M.box_mesh(3, rs, 2, basis, deformed=kind == "box_def", bc_kind=M.BC_NONPENETRATION, with_ldg=True)
        density_wave(m, basis)
        oracle.compute_write_face(basis, m)
    prepare_pde_state(m, rng, NAVIER_STOKES)
    out, ref, dts = run_pde_pair(oracle, emu_lib, m, basis, NAVIER_STOKES, n_steps=2)
    assert_pde_parity(out, ref, dts)
    out, ref, dts = run_pde_pair(oracle, emu_lib, m, basis, NAVIER_STOKES, n_steps=1, compute_residual=True)
    assert_pde_parity(out, ref, dts)


def test_cfl_cache_follows_the_state(oracle, emu_lib):
    from util import check_cfl_cache
    check_cfl_cache(oracle, emu_lib, 4, 2)


@pytest.mark.parametrize("nd,rs", [(2, 3), (3, 2)])
def test_every_device_boundary_condition(oracle, emu_lib, nd, rs):
    """Freestream, Copy, Nonpenetration, Outflow, Pressure_outflow and No_slip (isothermal / heat flux / thermal equilibrium) as ghost-state
    and flux conditions of a viscous step (src/Boundary_condition.cpp), each on its own subset of the boundary faces"""
    from util import mixed_bcs
    rng = np.random.default_rng(5)
    basis = hb.gauss_legendre(rs)
    m = M.soup_mesh(nd, rs, rng, n_car=10, n_def=24, n_ref=2, with_ldg=True)
    M.random_flow_state(m, rng)
    prepare_pde_state(m, rng, NAVIER_STOKES)
    n_bc_faces = m.bcs[0]["ghost_slot"].size
    mixed_bcs(m, rng)
    assert n_bc_faces >= 8, n_bc_faces
    out, ref, dts = run_pde_pair(oracle, emu_lib, m, basis, NAVIER_STOKES, n_steps=2)
    assert_pde_parity(out, ref, dts)


@pytest.mark.parametrize("nd,rs", [(1, 3), (2, 3), (3, 2)])
def test_riemann_invariants_bc(oracle, emu_lib, nd, rs):
    from util import check_riemann_bc
    check_riemann_bc(oracle, emu_lib, nd, rs)


@pytest.mark.parametrize("rs,n", [(6, 11), (4, 9), (8, 5)])
@pytest.mark.parametrize("deformed", [False, True])
def test_box_2d_pipelined_local(oracle, emu_lib, deformed, rs, n):
    """local_euler_pipe2d.cu: batches of elements per persistent CTA; sizes chosen so that a CTA iterates several times (both
    stage buffers reused) and the last batch is partial"""
    basis = hb.gauss_legendre(rs)
    m = M.box_mesh(2, rs, n, basis, deformed=deformed, bc_kind=M.BC_FREESTREAM, bc_params=freestream_state(2))
    density_wave(m, basis)
    oracle.compute_write_face(basis, m)
    out, ref, dts, _ = run_euler_pair(oracle, emu_lib, m, basis, n_steps=2)
    assert_euler_parity(out, ref, dts)


@pytest.mark.parametrize("nd,rs", [(1, 3), (2, 3), (3, 2)])
def test_is_admissible(oracle, emu_lib, nd, rs):
    """Solver::is_admissible / Element::record on the device"""
    from util import check_admissibility
    check_admissibility(oracle, emu_lib, nd, rs)


@pytest.mark.parametrize("nd,rs", [(2, 3), (3, 2), (3, 3)])
def test_set_jacobian(emu_lib, nd, rs):
    """metric terms of deformed elements computed on the device from vertex positions and node adjustments"""
    from util import check_set_jacobian
    check_set_jacobian(emu_lib, nd, rs)


@pytest.mark.parametrize("nd,rs", [(1, 4), (2, 3), (3, 3)])
def test_av_glue(emu_lib, nd, rs):
    """pointwise loops of the artificial-viscosity pipelines on the device"""
    from util import check_av_glue
    check_av_glue(emu_lib, nd, rs)


@pytest.mark.parametrize("nd,rs", [(2, 3), (3, 2)])
def test_aux_bcs(emu_lib, nd, rs):
    from util import check_aux_bcs
    check_aux_bcs(emu_lib, nd, rs)


@pytest.mark.parametrize("nd,rs", [(2, 3), (3, 2)])
def test_av_smoothness_pipeline(oracle, emu_lib, nd, rs):
    """the naca0012-class configuration's shock-capturing update, Solver::update_art_visc_smoothness, with no host loop left"""
    from util import check_av_pipeline
    check_av_pipeline(oracle, emu_lib, nd, rs)


@pytest.mark.parametrize("nd,rs", [(2, 4), (3, 3)])
def test_set_art_visc_admis(oracle, emu_lib, nd, rs):
    from util import check_set_art_visc_admis
    check_set_art_visc_admis(oracle, emu_lib, nd, rs)


@pytest.mark.parametrize("nd,rs", [(1, 3), (2, 3), (3, 2)])
def test_av_elwise(emu_lib, nd, rs):
    """Solver::update_art_visc_elwise: ramp, forcing loops, vertex branch"""
    from util import check_av_elwise
    check_av_elwise(emu_lib, nd, rs)


@pytest.mark.parametrize("nd,rs", [(1, 3), (2, 3), (3, 2)])
def test_vertex_sharing_and_fix_admis_spread(oracle, emu_lib, nd, rs):
    """share_vertex_data + the spreading step of fix_admissibility on the device"""
    from util import check_vertex_sharing
    check_vertex_sharing(oracle, emu_lib, nd, rs)


@pytest.mark.parametrize("nd,rs", [(1, 3), (2, 3), (3, 2), (3, 3)])
def test_shared_normals_soup(oracle, emu_lib, nd, rs):
    """connection passes of Solver::calc_jacobian on the device"""
    from util import check_shared_normals_soup
    check_shared_normals_soup(oracle, emu_lib, nd, rs)


@pytest.mark.parametrize("nd,rs,n", [(2, 3, 4), (3, 2, 3)])
def test_calc_jacobian_box(emu_lib, nd, rs, n):
    from util import check_calc_jacobian_box
    check_calc_jacobian_box(emu_lib, nd, rs, n)


@pytest.mark.parametrize("kind", ["soup", "box_def"])
def test_navier_stokes_3d_padded_layout_is_bit_identical(port_oracle, emu_lib, kind):
    """ns_local_pad_kernel (row size 6 default: plane-padded shared memory, lines dealt to lanes through tables, 16-byte accesses on the
    contiguous lines) performs the operations of ns_local_line_kernel in the same order: HEXED_B200_OPT_NS_LOCAL_LAYOUT = 0 / 1 must agree
    bit for bit, state, LDG faces and residual cache"""
    rng = np.random.default_rng(79)
    basis = hb.gauss_legendre(6)
    if kind == "soup":
        m = M.soup_mesh(3, 6, rng, n_car=3, n_def=5, n_ref=2, with_ldg=True)
        M.random_flow_state(m, rng)
    else:
        m = M.box_mesh(3, 6, 2, basis, deformed=True, bc_kind=M.BC_NONPENETRATION, with_ldg=True)
        density_wave(m, basis)
        port_oracle.compute_write_face(basis, m)
    prepare_pde_state(m, rng, NAVIER_STOKES)
    a, ref, dts = run_pde_pair(port_oracle, emu_lib, m, basis, NAVIER_STOKES, n_steps=2, options=((3, 0),))
    assert_pde_parity(a, ref, dts)
    for layout in (1, 2):
        b, _, _ = run_pde_pair(port_oracle, emu_lib, m, basis, NAVIER_STOKES, n_steps=2, options=((3, layout),))
        assert np.array_equal(a.elem_data, b.elem_data) and np.array_equal(a.face_ldg, b.face_ldg) and np.array_equal(a.face_state, b.face_state), layout


@pytest.mark.parametrize("rs,n", [(4, 5), (6, 4), (8, 3)])
@pytest.mark.parametrize("kind", ["soup", "box_car", "box_def"])
def test_navier_stokes_2d_line_kernel(oracle, emu_lib, kind, rs, n):
    """2-D row size 4 / 6 / 8 Navier-Stokes takes the batched line-task Local kernel (ns_local_line2d_kernel): several batches per
    launch, the last one partial"""
    rng = np.random.default_rng(78)
    basis = hb.gauss_legendre(rs)
    if kind == "soup":
        m = M.soup_mesh(2, rs, rng, n_car=7, n_def=15, n_ref=2, with_ldg=True)
        M.random_flow_state(m, rng)
    else:
        m = M.box_mesh(2, rs, n, basis, deformed=kind == "box_def", bc_kind=M.BC_NONPENETRATION, with_ldg=True)
        density_wave(m, basis)
        oracle.compute_write_face(basis, m)
    prepare_pde_state(m, rng, NAVIER_STOKES)
    out, ref, dts = run_pde_pair(oracle, emu_lib, m, basis, NAVIER_STOKES, n_steps=2)
    assert_pde_parity(out, ref, dts)
    out, ref, dts = run_pde_pair(oracle, emu_lib, m, basis, NAVIER_STOKES, n_steps=1, compute_residual=True)
    assert_pde_parity(out, ref, dts)


@pytest.mark.parametrize("nd,rs,n", [(2, 4, 5), (3, 4, 2)])
def test_fused_admissibility(oracle, emu_lib, nd, rs, n):
    from util import check_fused_admissibility
    check_fused_admissibility(oracle, emu_lib, nd, rs, n)


@pytest.mark.parametrize("mode,deformed", [(0, True), (2, True), (3, False), (4, True), (4, False)])
@pytest.mark.parametrize("rs,n", [(6, 3), (4, 3)])
def test_box_3d_other_local_kernels(oracle, emu_lib, rs, n, mode, deformed):
    """HEXED_B200_OPT_PIPELINED_LOCAL = 0 (the general Local kernel) and 2 (the pipelined kernel with its earlier, fully staged
    shared-memory layout) stay available for A/B measurements, 3 = Cartesian elements in the lean layout with four resident CTAs:
    same parity bar as the default kernels"""
    basis = hb.gauss_legendre(rs)
    m = M.box_mesh(3, rs, n, basis, deformed=deformed, bc_kind=M.BC_FREESTREAM, bc_params=freestream_state(3))
    density_wave(m, basis)
    oracle.compute_write_face(basis, m)
    out, ref, dts, _ = run_euler_pair(oracle, emu_lib, m, basis, n_steps=2, options=((0, mode),))
    assert_euler_parity(out, ref, dts)


@pytest.mark.parametrize("nd,rs,n", [(2, 4, 6), (3, 2, 4)])
def test_max_dt_running_screen_random_states(oracle, emu_lib, nd, rs, n):
    from util import check_max_dt_running_screen_random
    check_max_dt_running_screen_random(oracle, emu_lib, nd, rs, n, range(12))


@pytest.mark.parametrize("nd,rs,n", [(2, 4, 6), (3, 2, 4)])
def test_max_dt_running_screen_navier_stokes(oracle, emu_lib, nd, rs, n):
    from util import check_max_dt_running_screen_ns
    check_max_dt_running_screen_ns(oracle, emu_lib, nd, rs, n, range(12))


@pytest.mark.parametrize("nd,rs,n", [(2, 6, 12), (3, 4, 5)])
def test_max_dt_running_screen_is_exact(oracle, emu_lib, nd, rs, n):
    from util import check_max_dt_running_screen
    check_max_dt_running_screen(oracle, emu_lib, nd, rs, n)


@pytest.mark.parametrize("nd,rs,pde", [(2, 4, "euler"), (3, 3, "euler"), (2, 4, "navier_stokes")])
def test_time_step_scale_write_skipped_only_when_known_one(oracle, emu_lib, nd, rs, pde):
    from util import check_tss_write_skipped
    check_tss_write_skipped(oracle, emu_lib, nd, rs, pde)


@pytest.mark.parametrize("nd,rs,n,deformed", [(2, 6, 4, False), (2, 3, 3, True), (3, 4, 2, True)])
def test_update_euler_device_time_step(oracle, emu_lib, nd, rs, n, deformed):
    from util import check_update_euler
    check_update_euler(oracle, emu_lib, nd, rs, n, n_steps=4, use_graph=False, deformed=deformed)


def test_update_euler_refined_mesh(oracle, emu_lib):
    from util import check_update_euler
    check_update_euler(oracle, emu_lib, 2, 4, 4, n_steps=3, use_graph=False, refined=True)


@pytest.mark.parametrize("nd,rs,n", [(2, 4, 4), (2, 3, 3), (3, 4, 2)])
def test_update_navier_stokes_device_time_step(oracle, emu_lib, nd, rs, n):
    from util import check_update_navier_stokes
    check_update_navier_stokes(oracle, emu_lib, nd, rs, n, n_steps=3, use_graph=False)


@pytest.mark.parametrize("n_cheby,n_steps", [(3, 8), (4, 9)])
def test_update_loops_with_chebyshev_steps(oracle, emu_lib, n_cheby, n_steps):
    """n_cheby_flow > 1 (the shock-capturing cases): Chebyshev factors cycle, one graph per cycle, a partial cycle at the end"""
    from util import check_update_euler, check_update_navier_stokes
    check_update_euler(oracle, emu_lib, 2, 4, 4, n_steps=n_steps, use_graph=False, deformed=True, n_cheby=n_cheby)
    check_update_navier_stokes(oracle, emu_lib, 2, 4, 3, n_steps=n_steps, use_graph=False, n_cheby=n_cheby)
