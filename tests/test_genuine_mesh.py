"""Drop-in test on the reference's GENUINE classes: hexed_b200/host/adapter.cpp compiled with -DHEXED_B200_WITH_HEXED_HEADERS
against the reference's own include/kernels.hpp, driven on a mesh of genuine `Element` / `Deformed_element` /
`Element_face_connection` / `Refined_connection` / `Typed_bound_connection` objects (oracle/ref_genuine_mesh.cpp), against the
reference's own kernels on the identical object graph (closes "the HEXED_B200_WITH_HEXED_HEADERS branch has
never been compiled" and "never run against the genuine Element classes").

CPU: the adapter drives the host-thread emulation build of the CUDA sources (tests/emu); `-m gpu`: the product library on a B200.
Bars: state relative L2 <= 1e-11 after the stages, max_dt <= 1e-13 (north-star tolerances)."""
import os

import numpy as np
import pytest

import pyoracle
from pyoracle import EULER, NAVIER_STOKES
from util import rel_l2, STATE_TOL, MAX_DT_TOL

pytestmark = pytest.mark.skipif(not (pyoracle.ref_available() or os.path.isdir("/root/reference")),
                                reason="oracle/_ref is built from /root/reference, which this machine does not have")


@pytest.fixture(scope="module")
def ref_lib():
    import genuine
    pyoracle.RefOracle()  # builds oracle/_ref if need be
    return genuine.GenuineLib(genuine.REF)


@pytest.fixture(scope="module")
def adapter_emu(emu_lib):
    import subprocess
    import genuine
    if os.path.isdir("/root/reference"):
        subprocess.run(["make", "-C", pyoracle.HERE, "-f", "Makefile.ref", "-j8", "emu"], check=True, stdout=subprocess.DEVNULL)
    if not os.path.exists(genuine.ADAPTER_EMU):
        pytest.skip("libhexed_adapter_genuine_emu.so not built")
    return genuine.GenuineLib(genuine.ADAPTER_EMU)


@pytest.fixture(scope="module")
def adapter_gpu(gpu_lib):
    import genuine
    assert os.path.exists(genuine.ADAPTER), "oracle/_ref/libhexed_adapter_genuine.so must travel to the GPU box"
    return genuine.GenuineLib(genuine.ADAPTER)


def cell_kinds(nd, n, pattern, seed=0):
    rng = np.random.default_rng(seed)
    k = np.zeros((n,)*nd, np.int32)
    if pattern == "car":
        pass
    elif pattern == "def":
        k[:] = 1
    elif pattern == "mixed":
        k[:] = rng.integers(0, 2, k.shape)
    elif pattern == "refined_car":
        k[(1,)*nd] = 2
    elif pattern == "refined_def":
        k[:] = 1
        k[(1,)*nd] = 3
        k[(0,)*nd] = 3
    elif pattern == "refined_mixed":
        k[:] = rng.integers(0, 2, k.shape)
        k[(1,)*nd] |= 2
        k[(n - 1,)*nd] |= 2
    return k


def initial_state(pos, nd):
    """smooth subsonic flow with a density / velocity wave, evaluated at the genuine quadrature point positions"""
    ne, _, nq = pos.shape
    phase = sum((1 + 0.5*d)*pos[:, d] for d in range(nd))
    mass = 1.2*(1 + 0.1*np.sin(2*np.pi*phase))
    st = np.zeros((ne, nd + 2, nq))
    vel = [30.*(1 + d) + 10.*np.cos(2*np.pi*phase + d) for d in range(nd)]
    for d in range(nd):
        st[:, d] = mass*vel[d]
    st[:, nd] = mass
    st[:, nd + 1] = 1e5/0.4 + 0.5*mass*sum(v*v for v in vel)
    return st


def prepare(ref_lib, nd, rs, n, pattern, viscous=False):
    import genuine
    kind = cell_kinds(nd, n, pattern)
    g = genuine.GenuineMesh(ref_lib, nd, rs, n, kind)
    g.calc_jacobian()
    g.set_slots(0, initial_state(g.positions(), nd))
    if viscous:
        rng = np.random.default_rng(1)
        g.set_slots(nd + 3, 1e-4*rng.random((g.n_elem, 2, g.nq)))
    g.compute_write_face()
    return kind, g, g.export()


def run(g, viscous, n_steps=2, resident=False, safety=0.5):
    visc, cond = pyoracle.sutherland(1.7e-5, 273., 110.), pyoracle.constant(2.5e-2)
    dts = []

    def bcs():
        if resident:
            g.boundary_faces_to_host()
        g.bc_copy()
        if resident:
            g.ghost_faces_to_device()
    for _ in range(n_steps):
        if viscous:
            dt = g.max_dt(NAVIER_STOKES, safety, safety, False, visc, cond)
            bcs()
            g.compute_navier_stokes(visc, cond, dt, 0)
        else:
            dt = g.max_dt(EULER, safety, safety, False)
            bcs()
            g.compute_euler(dt, 0)
        bcs()
        g.compute_euler(dt, 1)
        dts.append(dt)
    return dts


def compare(alib, ref_lib, nd, rs, n, pattern, viscous, resident):
    import genuine
    kind, g_ref, blob0 = prepare(ref_lib, nd, rs, n, pattern, viscous)
    g_ad = genuine.GenuineMesh(alib, nd, rs, n, kind)
    try:
        assert (g_ad.n_car, g_ad.n_def, g_ad.n_car_con, g_ad.n_def_con, g_ad.n_ref, g_ad.n_bc) == \
               (g_ref.n_car, g_ref.n_def, g_ref.n_car_con, g_ref.n_def_con, g_ref.n_ref, g_ref.n_bc)
        g_ad.load(blob0)
        fc = g_ad.flatten_counts()  # the genuine pointer graph flattens to the expected table sizes
        assert fc[0] == g_ref.n_car and fc[1] == g_ref.n_def and fc[4] == g_ref.n_car_con and fc[5] == g_ref.n_def_con
        assert fc[6] == g_ref.n_ref and fc[7] == g_ref.n_bc
        alib.lib.hr_gm_set_sync_mode(int(resident))
        if resident:
            g_ad.to_device()
        dts_ref = run(g_ref, viscous)
        dts_ad = run(g_ad, viscous, resident=resident)
        if resident:
            g_ad.to_host()
        for a, b in zip(dts_ad, dts_ref):
            assert abs(a - b) <= MAX_DT_TOL*abs(b), (a, b)
        assert rel_l2(g_ad.state(), g_ref.state()) <= STATE_TOL
        assert rel_l2(g_ad.slots(nd + 2, 1), g_ref.slots(nd + 2, 1)) <= MAX_DT_TOL  # time step scale
        # every double either implementation can see (elements, Jacobian data, every connection's faces and normals)
        assert rel_l2(g_ad.export(), g_ref.export()) <= 10*STATE_TOL
        assert g_ad.work_units() == g_ref.work_units()  # Stopwatch_tree side-contract (include/kernel_factory.hpp:32-45)
    finally:
        alib.lib.hr_gm_set_sync_mode(0)
        alib.lib.hr_gm_release()
        g_ad.close(); g_ref.close()


CASES = [(2, 4, 3, "car", False), (2, 4, 3, "def", False), (2, 3, 4, "mixed", False), (2, 4, 3, "refined_car", False),
         (2, 4, 3, "refined_def", False), (3, 3, 3, "refined_mixed", False), (3, 4, 2, "def", False),
         (2, 4, 3, "def", True), (2, 3, 3, "refined_def", True), (3, 3, 2, "mixed", True)]


@pytest.mark.parametrize("resident", [False, True])
@pytest.mark.parametrize("nd,rs,n,pattern,viscous", CASES)
def test_adapter_on_genuine_classes_emulated(adapter_emu, ref_lib, nd, rs, n, pattern, viscous, resident):
    compare(adapter_emu, ref_lib, nd, rs, n, pattern, viscous, resident)


@pytest.mark.gpu
@pytest.mark.parametrize("resident", [False, True])
@pytest.mark.parametrize("nd,rs,n,pattern,viscous", CASES + [(3, 6, 4, "refined_mixed", False), (3, 6, 3, "def", True), (2, 6, 8, "refined_def", True)])
def test_adapter_on_genuine_classes_gpu(adapter_gpu, ref_lib, nd, rs, n, pattern, viscous, resident):
    compare(adapter_gpu, ref_lib, nd, rs, n, pattern, viscous, resident)


def test_genuine_layout_matches_flat_mesh_conventions(ref_lib):
    """the slot order / face aliasing the device mirror assumes, read off the genuine classes: the state written through
    `set_slots` comes back through the kernels' own view (compute_write_face extrapolates it to the faces the connections own)"""
    import genuine
    nd, rs, n = 2, 3, 2
    g = genuine.GenuineMesh(ref_lib, nd, rs, n, cell_kinds(nd, n, "def"))
    try:
        g.calc_jacobian()
        st = initial_state(g.positions(), nd)
        g.set_slots(0, st)
        assert np.array_equal(g.state(), st)
        assert np.array_equal(g.slots(nd + 2, 1), np.ones((g.n_elem, 1, g.nq)))  # time_step_scale initialised to 1 (src/Element.cpp:24)
        assert g.n_def == n**nd and g.n_car == 0 and g.n_bc == 2*nd*n**(nd - 1) and g.n_def_con == nd*(n - 1)*n**(nd - 1) + g.n_bc
    finally:
        g.close()
