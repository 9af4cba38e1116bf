"""CPU-side check of the CUDA kernels' logic: the product's .cu sources compiled for host threads
(tests/emu/cuda_emu.hpp) against the oracle. No GPU needed; small meshes only. The real parity tests are in
test_gpu_parity.py (-m gpu); this file exists so indexing bugs are caught in the GPU-less container."""
import numpy as np
import pytest

import hexed_b200 as hb
from hexed_b200 import mesh as M
from util import run_euler_pair, assert_euler_parity, density_wave, freestream_state


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
