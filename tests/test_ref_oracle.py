"""The restated CPU oracle (oracle/oracle_impl.hpp) against the REFERENCE'S OWN kernels (oracle/_ref/libhexed_ref.so:
src/kernels_convective.cpp, kernels_diffusive.cpp, kernels_max_dt.cpp, stabilizing_art_visc.cpp + include/Spatial.hpp, pde.hpp
compiled unmodified, recipe oracle/Makefile.ref) on identical meshes and states.

This is the pin that closes the "parity unpinned" rows of SURVEY section 8c: Advection / Smooth_art_visc / Fix_therm_admis
(pde.hpp:265-493), Stab_art_visc (stabilizing_art_visc.cpp:30-60), Neighbor_reconcile / Reconcile_ldg_flux, and the Euler / NS flux
and LLF dissipation (Spatial.hpp:671-674) at 1e-14 instead of the 1e-3 margin of the reference's marching test.
Tolerance: 1e-13 relative L2 per slot group (the two builds differ only in FMA contraction and summation order), 1e-14 for dt.
"""
import numpy as np
import pytest

import hexed_b200 as hb
from hexed_b200 import mesh as M
import pyoracle
from pyoracle import EULER, NAVIER_STOKES, ADVECTION, SMOOTH_AV, FIX_THERM_ADMIS
from util import rel_l2, prepare_pde_state

TOL = 1e-13
DT_TOL = 1e-14

pytestmark = pytest.mark.skipif(not (pyoracle.ref_available() or __import__("os").path.isdir("/root/reference")),
                                reason="oracle/_ref is built from /root/reference, which this machine does not have")


@pytest.fixture(scope="module")
def port():
    return pyoracle.Oracle()


@pytest.fixture(scope="module")
def ref():
    return pyoracle.RefOracle()


def drive(o, mesh, basis, pde, n_steps=1, local_time=False, use_filter=False, safety=0.3, compute_residual=False):
    """the call sequence Solver makes for each PDE (src/Solver.cpp:834-886, 484-518, 577-578, 1058-1074), on one oracle"""
    visc, cond = pyoracle.sutherland(1.7e-5, 273., 110.), pyoracle.constant(2.5e-2)
    dts = []
    for _ in range(n_steps):
        if pde == EULER:
            dt = o.max_dt(pde, basis, mesh, safety, safety, local_time)
            for stage in (0, 1):
                if compute_residual and stage:
                    break
                o.apply_state_bcs(mesh)
                o.compute_euler(basis, mesh, dt=dt, i_stage=stage, use_filter=use_filter, compute_residual=compute_residual)
        elif pde == NAVIER_STOKES:
            dt = o.max_dt(pde, basis, mesh, safety, safety, local_time, visc, cond)
            o.apply_state_bcs(mesh)
            o.compute_navier_stokes(basis, mesh, lambda: o.apply_flux_bcs(mesh), visc, cond, dt=dt, i_stage=0, use_filter=use_filter,
                                    compute_residual=compute_residual)
            if not compute_residual:
                o.apply_state_bcs(mesh)
                o.compute_euler(basis, mesh, dt=dt, i_stage=1, use_filter=use_filter)
        elif pde == ADVECTION:
            dt = o.max_dt(pde, basis, mesh, safety, safety, local_time, advect_length=0.7)
            for stage in (0, 1):
                o.compute_advection(basis, mesh, 0.7, dt=dt, i_stage=stage, use_filter=use_filter)
        elif pde == SMOOTH_AV:
            dt = o.max_dt(pde, basis, mesh, safety, safety, local_time)
            o.compute_smooth_av(basis, mesh, None, 0.4, 1.3, dt=dt, i_stage=0, use_filter=use_filter)
        else:
            dt = o.max_dt(pde, basis, mesh, safety, safety, local_time)
            o.compute_fix_therm_admis(basis, mesh, None, dt=dt, i_stage=0, use_filter=use_filter, compute_residual=compute_residual)
        dts.append(dt)
    return dts


def assert_same(a, b, tol=TOL):
    nd, rs = a.n_dim, a.row_size
    c = M.cache_slot(nd, rs)
    groups = {"state": (0, nd + 2), "tss": (nd + 2, nd + 3), "av": (nd + 3, nd + 5), "forcing": (nd + 5, nd + 9),
              "advection": (nd + 9, nd + 9 + rs), "cache": (c, a.n_slot)}
    for name, (lo, hi) in groups.items():
        x, y = a.elem_data[:, lo:hi], b.elem_data[:, lo:hi]
        assert np.isfinite(y).all(), name
        if np.linalg.norm(y) == 0:
            assert np.array_equal(x, y), name
        else:
            assert rel_l2(x, y) <= (10*tol if name == "cache" else tol), (name, rel_l2(x, y))
    for name in ("face_state", "face_ldg", "face_wide"):
        x, y = getattr(a, name), getattr(b, name)
        if x is not None and np.linalg.norm(y) > 0:
            assert rel_l2(x, y) <= tol, (name, rel_l2(x, y))
    assert rel_l2(a.uncert, b.uncert) <= tol if np.linalg.norm(b.uncert) else np.array_equal(a.uncert, b.uncert)


def soup(nd, rs, seed, pde, **kw):
    rng = np.random.default_rng(seed)
    m = M.soup_mesh(nd, rs, rng, with_ldg=True, with_wide=True, **kw)
    M.random_flow_state(m, rng)
    prepare_pde_state(m, rng, pde)
    return m


@pytest.mark.parametrize("row_size", range(2, 9))
def test_basis_tables_are_the_reference_generators(ref, row_size):
    """hexed_b200/data/basis_tables.json (parsed from the generator's text by oracle/gen_basis.py) against the reference's compiled
    Gauss_legendre: every table bit for bit; max_cfl / step_ratio (src/Basis.cpp:6-14) to the last bit as well"""
    b = hb.gauss_legendre(row_size)
    t = ref.basis_tables(row_size)
    for k in ("node", "weight", "diff_mat", "boundary", "orthogonal", "filter", "prolong", "restrict"):
        assert np.array_equal(np.asarray(getattr(b, k)), t[k]), k
    assert b.min_eig_diffusion == t["min_eig_diffusion"]
    assert -2*b.quadratic_safety/b.min_eig_convection == t["max_cfl"]
    assert .5/b.quadratic_safety == t["step_ratio"]


@pytest.mark.parametrize("pde", [EULER, NAVIER_STOKES, ADVECTION, SMOOTH_AV, FIX_THERM_ADMIS])
@pytest.mark.parametrize("nd,rs", [(1, 3), (2, 2), (2, 4), (2, 6), (3, 2), (3, 3), (3, 6)])
def test_stage_sequences_all_pdes(port, ref, pde, nd, rs):
    basis = hb.gauss_legendre(rs)
    m = soup(nd, rs, 100*nd + rs, pde)
    port.compute_write_face(basis, m)
    a, b = m.copy(), m.copy()
    da = drive(port, a, basis, pde, n_steps=2)
    db = drive(ref, b, basis, pde, n_steps=2)
    for x, y in zip(da, db):
        assert abs(x - y) <= DT_TOL*abs(y)
    assert_same(a, b)


@pytest.mark.parametrize("pde", [EULER, NAVIER_STOKES, ADVECTION, SMOOTH_AV, FIX_THERM_ADMIS])
@pytest.mark.parametrize("local_time,use_filter,compute_residual", [(True, False, False), (False, True, False), (True, True, True)])
def test_options(port, ref, pde, local_time, use_filter, compute_residual):
    if compute_residual and pde in (ADVECTION, SMOOTH_AV):
        pytest.skip("Solver never asks these PDEs for a residual")
    nd, rs = 2, 5
    basis = hb.gauss_legendre(rs)
    m = soup(nd, rs, 7, pde)
    port.compute_write_face(basis, m)
    a, b = m.copy(), m.copy()
    da = drive(port, a, basis, pde, local_time=local_time, use_filter=use_filter, compute_residual=compute_residual)
    db = drive(ref, b, basis, pde, local_time=local_time, use_filter=use_filter, compute_residual=compute_residual)
    assert da == db or abs(da[0] - db[0]) <= DT_TOL*abs(db[0])
    assert_same(a, b)


@pytest.mark.parametrize("nd,rs", [(1, 4), (2, 3), (2, 6), (3, 4), (3, 6)])
def test_stabilizing_art_visc(port, ref, nd, rs):
    """Stab_art_visc (src/stabilizing_art_visc.cpp:8-66) had no reference-held pin at all: rough density so that elements land on
    both plateaus and on the ramp"""
    basis = hb.gauss_legendre(rs)
    rng = np.random.default_rng(5)
    m = soup(nd, rs, 3, EULER, n_car=24, n_def=24)
    amp = np.logspace(-6, -0.3, m.n_elem)
    m.elem_data[:, nd] = 1.2*(1. + amp[:, None]*rng.standard_normal((m.n_elem, m.nq)))
    a, b = m.copy(), m.copy()
    port.stabilizing_art_visc(basis, a, 340.)
    ref.stabilizing_art_visc(basis, b, 340.)
    assert (b.uncert == 0).any() and (b.uncert > 0).any()
    ramp = (b.uncert > 0) & (b.uncert < (rs - 1)*340.*b.nom_size)
    assert ramp.any()
    # the indicator is a log of a difference of nearly equal projections: compare the ramp values with a tolerance scaled by the plateau
    assert np.abs(a.uncert - b.uncert).max() <= 1e-9*np.abs(b.uncert).max()
    assert np.array_equal(a.uncert == 0, b.uncert == 0)


@pytest.mark.parametrize("nd,rs", [(2, 4), (3, 3), (3, 6)])
@pytest.mark.parametrize("scale,offset", [(False, False), (True, False), (True, True), (False, True)])
def test_prolong_restrict_write_face(port, ref, nd, rs, scale, offset):
    basis = hb.gauss_legendre(rs)
    m = soup(nd, rs, 11, EULER)
    rng = np.random.default_rng(2)
    m.face_state[:] = rng.standard_normal(m.face_state.shape)
    m.face_ldg[:] = rng.standard_normal(m.face_ldg.shape)
    for fn in ("compute_prolong", "compute_restrict"):
        a, b = m.copy(), m.copy()
        getattr(port, fn)(basis, a, scale=scale, offset=offset)
        getattr(ref, fn)(basis, b, scale=scale, offset=offset)
        assert_same(a, b, 1e-14)
    for pde in (EULER, ADVECTION, SMOOTH_AV):
        a, b = m.copy(), m.copy()
        prepare_pde_state(a, np.random.default_rng(4), pde); prepare_pde_state(b, np.random.default_rng(4), pde)
        port.compute_write_face(basis, a, pde=pde)
        ref.compute_write_face(basis, b, pde=pde)
        assert_same(a, b, 1e-14)
    a, b = m.copy(), m.copy()
    port.compute_prolong(basis, a, pde=ADVECTION)
    ref.compute_prolong(basis, b, pde=ADVECTION)
    assert_same(a, b, 1e-14)


@pytest.mark.parametrize("nd,rs", [(2, 5), (3, 4), (3, 6)])
def test_face_permutation_bit_exact(port, ref, nd, rs):
    rng = np.random.default_rng(9)
    nfq = rs**(nd - 1)
    for direction in M.all_directions(nd):
        data = rng.standard_normal((nd + 2)*nfq)
        for restore in (False, True):
            a, b = data.copy(), data.copy()
            port.face_permutation(nd, rs, nd + 2, direction, a, restore=restore)
            ref.face_permutation(nd, rs, nd + 2, direction, b, restore=restore)
            assert np.array_equal(a, b), (direction.as_list(), restore)


@pytest.mark.parametrize("rs", range(2, 9))
def test_derivative(port, ref, rs):
    rng = np.random.default_rng(rs)
    basis = hb.gauss_legendre(rs)
    q, bv = rng.standard_normal((3, rs)), rng.standard_normal((3, 2))
    assert rel_l2(port.derivative(basis, q, bv), ref.derivative(basis, q, bv)) <= 1e-14


@pytest.mark.parametrize("nd", [1, 2, 3])
def test_characteristics_qr(port, ref, nd):
    """Navier_stokes::Pde::Characteristics (pde.hpp:181-256) with the shim's ColPivHouseholderQR against the oracle's 3x3 restatement"""
    rng = np.random.default_rng(40 + nd)
    for _ in range(50):
        mass = 0.5 + rng.random()
        veloc = 200*rng.standard_normal(nd)
        pres = 1e5*(0.5 + rng.random())
        state = np.concatenate([mass*veloc, [mass, pres/.4 + .5*mass*veloc@veloc]])
        state1 = state*(1 + 0.1*rng.standard_normal(nd + 2))
        direction = rng.standard_normal(nd)
        va, da = port.characteristics(state, direction, state1)
        vb, db = ref.characteristics(state, direction, state1)
        assert np.allclose(va, vb, rtol=1e-14, atol=0)
        assert rel_l2(da, db) <= 1e-12
        assert rel_l2(db.sum(1), state1) <= 1e-12  # test/test_Characteristics.cpp:20-24: the decomposition sums to the state
