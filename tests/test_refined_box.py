"""Locally refined Cartesian boxes (hexed_b200.mesh.refined_box_mesh): real 2:1 hanging-node geometry built with the reference's
`Refined_connection<Element>` conventions (include/connection.hpp:195-262) -- the topology class of the adaptive BASELINE configs
(naca0012, onera_m6). Size-independent properties the discretisation guarantees, checked first on the oracle and then on the device:
  * a uniform free stream is preserved to round-off across hanging faces (prolong/restrict reproduce constants);
  * with slip walls, total mass and total energy are conserved to round-off (the mortar transfer is conservative: what
    `test_conservation` of the reference's test/test_Solver.cpp:610-639 asserts);
plus the usual parity of the CUDA path against the oracle on the same mesh."""
import numpy as np
import pytest

import hexed_b200 as hb
from hexed_b200 import mesh as M
from hexed_b200.kernels import Device
from hexed_b200.cases import density_wave, freestream_state
from pyoracle import EULER
from util import run_euler_pair, assert_euler_parity, rel_l2


def refine_pattern(nd, n, seed=0):
    rng = np.random.default_rng(seed)
    refine = rng.random((n,)*nd) < 0.3
    refine[(0,)*nd] = True           # a refined corner cell: hanging faces next to boundary faces
    refine[(n - 1,)*nd] = False
    return refine


def totals(m, basis):
    w = basis.weight
    W = w
    for _ in range(m.n_dim - 1):
        W = np.multiply.outer(W, w)
    W = W.reshape(-1)
    return np.array([(m.cell_volume[:, None]*W[None, :]*m.state()[:, v]).sum() for v in range(m.n_dim + 2)])


def advance(step_fn, n_steps):
    for _ in range(n_steps):
        step_fn()


@pytest.mark.parametrize("nd,rs,n", [(2, 4, 5), (3, 3, 3)])
def test_oracle_freestream_and_conservation(oracle, nd, rs, n):
    basis = hb.gauss_legendre(rs)
    refine = refine_pattern(nd, n)
    fs = freestream_state(nd)
    m = M.refined_box_mesh(nd, rs, n, basis, refine, bc_kind=M.BC_FREESTREAM, bc_params=fs)
    assert m.ref_face.shape[0] > 0
    m.state()[:] = fs[None, :, None]
    oracle.compute_write_face(basis, m); oracle.compute_prolong(basis, m)
    start = m.state().copy()

    def step(mm):
        dt = oracle.max_dt(EULER, basis, mm, 0.5, 0.5, False)
        for stage in (0, 1):
            oracle.apply_state_bcs(mm); oracle.compute_euler(basis, mm, dt=dt, i_stage=stage)
    advance(lambda: step(m), 3)
    assert np.abs(m.state() - start).max() <= 1e-13*np.abs(start).max()
    m = M.refined_box_mesh(nd, rs, n, basis, refine, bc_kind=M.BC_NONPENETRATION)
    density_wave(m, basis)
    oracle.compute_write_face(basis, m); oracle.compute_prolong(basis, m)
    t0, s0 = totals(m, basis), m.state().copy()
    advance(lambda: step(m), 5)
    t1 = totals(m, basis)
    assert np.abs(m.state() - s0).max() > 1e-2*np.abs(s0).max()  # something happened
    assert abs(t1[nd] - t0[nd]) <= 1e-13*t0[nd] and abs(t1[nd + 1] - t0[nd + 1]) <= 1e-13*t0[nd + 1]


def test_device_parity_emu(oracle, emu_lib):
    basis = hb.gauss_legendre(3)
    m = M.refined_box_mesh(2, 3, 4, basis, refine_pattern(2, 4), bc_kind=M.BC_NONPENETRATION)
    density_wave(m, basis)
    oracle.compute_write_face(basis, m); oracle.compute_prolong(basis, m)
    out, ref, dts, _ = run_euler_pair(oracle, emu_lib, m, basis, n_steps=2)
    assert_euler_parity(out, ref, dts)


@pytest.mark.gpu
@pytest.mark.parametrize("nd,rs,n", [(2, 6, 8), (3, 6, 4), (3, 4, 5)])
def test_device_parity(oracle, gpu_lib, nd, rs, n):
    basis = hb.gauss_legendre(rs)
    m = M.refined_box_mesh(nd, rs, n, basis, refine_pattern(nd, n), bc_kind=M.BC_NONPENETRATION)
    density_wave(m, basis)
    oracle.compute_write_face(basis, m); oracle.compute_prolong(basis, m)
    out, ref, dts, _ = run_euler_pair(oracle, gpu_lib, m, basis, n_steps=4)
    assert_euler_parity(out, ref, dts)


@pytest.mark.gpu
@pytest.mark.parametrize("nd,n", [(2, 96), (3, 20)])
def test_device_properties_at_size(gpu_lib, nd, n):
    """sizes the oracle does not do in seconds (2-D: 96^2 cells, 3-D: 20^3 cells, 30 % refined, row size 6): free-stream preservation
    and mass / energy conservation on the device alone"""
    rs = 6
    basis = hb.gauss_legendre(rs)
    refine = refine_pattern(nd, n, seed=3)
    fs = freestream_state(nd)
    m = M.refined_box_mesh(nd, rs, n, basis, refine, bc_kind=M.BC_FREESTREAM, bc_params=fs)
    m.state()[:] = fs[None, :, None]
    dev = Device(nd, rs, basis, lib_path=gpu_lib).load_mesh(m)
    dev.compute_write_face(); dev.compute_prolong()

    def step():
        dt = dev.max_dt_euler(0.5, 0.5, False)
        for stage in (0, 1):
            dev.apply_state_bcs(); dev.compute_euler(dt=dt, i_stage=stage)
    start = m.state().copy()
    advance(step, 3)
    dev.sync_to_host(m)
    assert np.abs(m.state() - start).max() <= 1e-13*np.abs(start).max()
    dev.close()
    m = M.refined_box_mesh(nd, rs, n, basis, refine, bc_kind=M.BC_NONPENETRATION)
    density_wave(m, basis)
    dev = Device(nd, rs, basis, lib_path=gpu_lib).load_mesh(m)
    dev.compute_write_face(); dev.compute_prolong()
    t0 = totals(m, basis)
    advance(step, 5)
    dev.sync_to_host(m)
    t1 = totals(m, basis)
    assert abs(t1[nd] - t0[nd]) <= 1e-12*t0[nd] and abs(t1[nd + 1] - t0[nd + 1]) <= 1e-12*t0[nd + 1]
    dev.close()
