"""Domain decomposition (hexed_b200/partition.py, halo.py): the partitioned run must reproduce the undivided run bit for bit.

CPU-only: the kernels are the oracle (numpy meshes) or the host-thread emulation of the CUDA sources; the exchange runs in
process and, for the world_size-2 test, over torch.distributed/gloo. The same driver logic runs on NCCL in bench.py."""
import copy
import os
import sys

import numpy as np
import pytest

import hexed_b200 as hb
from hexed_b200 import mesh as M
from hexed_b200 import partition as P
from hexed_b200.halo import exchange_in_process
from hexed_b200.cases import density_wave, freestream_state
from pyoracle import EULER


@pytest.fixture(scope="module")
def oracle():
    """these tests are about the partitioning logic: they run split phases through the restated oracle's individual kernels
    (neighbor / local, which oracle/_ref does not expose one by one) and demand BIT equality with the undivided run, so both
    sides must be the same arithmetic: the restated oracle only"""
    import pyoracle
    return pyoracle.Oracle()


def oracle_pre_prolong(oracle, basis, m):
    if len(m.pre_prolong):
        view = copy.copy(m)
        view.ref_face = np.ascontiguousarray(m.ref_face[m.pre_prolong])
        oracle.compute_prolong(basis, view)


def oracle_step_parts(oracle, basis, parts, exchange, safety=0.3):
    dt = min(oracle.max_dt(EULER, basis, m, safety, safety, False) for m in parts)
    for stage in (0, 1):
        for m in parts:
            oracle.apply_state_bcs(m)
        exchange()
        for m in parts:
            oracle_pre_prolong(oracle, basis, m)
            oracle.compute_euler(basis, m, dt=dt, i_stage=stage)
    return dt


def in_process_exchange(parts):
    def get(p, slots):
        return parts[p].face_state[slots].copy()

    def put(q, slots, data):
        parts[q].face_state[slots] = data
    return lambda: exchange_in_process(parts, get, put)


def make_case(kind, rng, oracle=None):
    basis, m = _make_case(kind, rng)
    if kind.startswith("soup"):
        # the soup's face data is random; make the mortar faces consistent with their coarse faces, as they are after any stage
        from pyoracle import Oracle
        (oracle or Oracle()).compute_prolong(basis, m)
    return basis, m


def _make_case(kind, rng):
    if kind == "soup2d":
        basis = hb.gauss_legendre(3)
        m = M.soup_mesh(2, 3, rng, n_car=8, n_def=14, n_ref=6, with_ldg=False)
        M.random_flow_state(m, rng)
    elif kind == "soup3d":
        basis = hb.gauss_legendre(2)
        m = M.soup_mesh(3, 2, rng, n_car=8, n_def=14, n_ref=8, with_ldg=False)
        M.random_flow_state(m, rng)
    elif kind.startswith("refined_box"):
        # C5 class: Cartesian hanging-node faces as `Refined_connection<Element>` builds them, cut by the partition
        nd = 3 if kind.endswith("3d") else 2
        basis = hb.gauss_legendre(3)
        n = 3 if nd == 3 else 5
        refine = np.zeros((n,)*nd, bool)
        refine[(1,)*nd] = True
        refine[(n - 1,)*nd] = True
        if nd == 2:
            refine[2, 3] = refine[3, 3] = True
        m = M.refined_box_mesh(nd, 3, n, basis, refine, bc_kind=M.BC_COPY)
        M.random_flow_state(m, rng, mach=0.2)
    elif kind == "box_def":
        basis = hb.gauss_legendre(3)
        m = M.box_mesh(3, 3, 4, basis, deformed=True, bc_kind=M.BC_FREESTREAM, bc_params=freestream_state(3))
        density_wave(m, basis)
    else:
        basis = hb.gauss_legendre(4)
        m = M.box_mesh(2, 4, 6, basis, deformed=False, bc_kind=M.BC_COPY)
        density_wave(m, basis)
    return basis, m


def reference_run(oracle, basis, m, n_steps, safety=0.3):
    ref = m.copy()
    if not hasattr(m, "box_n"):
        pass
    dts = []
    for _ in range(n_steps):
        dt = oracle.max_dt(EULER, basis, ref, safety, safety, False)
        dts.append(dt)
        for stage in (0, 1):
            oracle.apply_state_bcs(ref)
            oracle.compute_euler(basis, ref, dt=dt, i_stage=stage)
    return ref, dts


@pytest.mark.parametrize("kind,n_parts", [("soup2d", 2), ("soup2d", 3), ("soup3d", 4), ("box_def", 2), ("box_def", 8), ("box_car", 4),
                                          ("refined_box2d", 2), ("refined_box2d", 5), ("refined_box3d", 3)])
def test_partitioned_oracle_matches_undivided(oracle, kind, n_parts):
    rng = np.random.default_rng(42)
    basis, m = make_case(kind, rng)
    if kind.startswith("refined_box"):
        oracle.compute_write_face(basis, m)
        oracle.compute_prolong(basis, m)
        part = rng.integers(0, n_parts, m.n_elem)
    elif kind.startswith("box"):
        oracle.compute_write_face(basis, m)
        part = P.split_by_curve(P.morton_keys(m.elem_index), n_parts)
    else:
        part = rng.integers(0, n_parts, m.n_elem)  # arbitrary ownership: every kind of cut, including split hanging faces
    ref, dts = reference_run(oracle, basis, m, 2)
    parts = P.partition_mesh(m, part, n_parts)
    assert sum(p.n_elem for p in parts) == m.n_elem
    ex = in_process_exchange(parts)
    for step in range(2):
        dt = oracle_step_parts(oracle, basis, parts, ex)
        assert dt == dts[step]
    out = m.copy()
    P.gather_elements(parts, out)
    assert np.array_equal(out.elem_data, ref.elem_data)
    nf = 2*m.n_dim
    assert np.array_equal(out.face_state[:nf*m.n_elem], ref.face_state[:nf*m.n_elem])


def test_morton_split_is_balanced_and_compact():
    idx = np.stack(np.meshgrid(*[np.arange(8)]*3, indexing="ij"), -1).reshape(-1, 3)
    part = P.split_by_curve(P.morton_keys(idx), 8)
    assert np.bincount(part).tolist() == [64]*8
    for p in range(8):  # Z-order octants of a power-of-two box are sub-cubes
        sel = idx[part == p]
        assert (sel.max(0) - sel.min(0)).tolist() == [3, 3, 3]


@pytest.mark.parametrize("kind,n_parts", [("soup2d", 3), ("box_def", 2), ("soup3d", 4), ("refined_box3d", 3)])
def test_partitioned_device_begin_finish(oracle, emu_lib, kind, n_parts):
    """the split stage (compute_euler_begin / exchange / compute_euler_finish) of the CUDA sources, several parts in one process"""
    from hexed_b200.kernels import Device
    rng = np.random.default_rng(7)
    basis, m = make_case(kind, rng)
    if kind.startswith("refined_box"):
        oracle.compute_write_face(basis, m)
        oracle.compute_prolong(basis, m)
        part = rng.integers(0, n_parts, m.n_elem)
    elif kind.startswith("box"):
        oracle.compute_write_face(basis, m)
        part = P.split_by_curve(P.morton_keys(m.elem_index), n_parts)
    else:
        part = rng.integers(0, n_parts, m.n_elem)
    ref, dts = reference_run(oracle, basis, m, 1)
    parts = P.partition_mesh(m, part, n_parts)
    devs = [Device(m.n_dim, m.row_size, basis, lib_path=emu_lib).load_mesh(p) for p in parts]
    w = m.nv*m.nfq

    def get(p, slots):
        buf = np.empty((len(slots), w))
        devs[p].face_list_gather(devs[p].send_lists[[q for q, s in parts[p].halo.send.items() if s is slots][0]], buf)
        devs[p].synchronize()
        return buf

    def put(q, slots, data):
        peer = [p for p, s in parts[q].halo.recv.items() if s is slots][0]
        devs[q].face_list_scatter(devs[q].recv_lists[peer], np.ascontiguousarray(data))
        devs[q].synchronize()
    dt = min(d.max_dt_euler(0.3, 0.3, False) for d in devs)
    assert abs(dt/dts[0] - 1) <= 1e-13
    for stage in (0, 1):
        for d in devs:
            d.apply_state_bcs()
            d.compute_euler_begin()
        exchange_in_process(parts, get, put)
        for d in devs:
            d.compute_euler_finish(dt=dts[0], i_stage=stage)
    for d, p in zip(devs, parts):
        d.sync_to_host(p)
        d.close()
    out = m.copy()
    P.gather_elements(parts, out)
    from util import rel_l2
    assert rel_l2(out.state(), ref.state()) <= 1e-11
    nf = 2*m.n_dim
    assert rel_l2(out.face_state[:nf*m.n_elem], ref.face_state[:nf*m.n_elem]) <= 1e-11


def _gloo_worker(rank, world, port, result_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    here = os.path.dirname(os.path.abspath(__file__))
    for p in (os.path.dirname(here), os.path.join(os.path.dirname(here), "oracle"), here):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    from pyoracle import Oracle
    from hexed_b200.halo import MeshHalo, allreduce_min, allreduce_and
    dist.init_process_group("gloo", rank=rank, world_size=world)
    oracle = Oracle()
    rng = np.random.default_rng(3)
    basis, m = make_case("soup3d", rng)
    part = rng.integers(0, world, m.n_elem)
    mine = P.partition_mesh(m, part, world)[rank]
    halo = MeshHalo(mine)
    for _ in range(2):
        dt = allreduce_min(oracle.max_dt(EULER, basis, mine, 0.3, 0.3, False))
        for stage in (0, 1):
            oracle.apply_state_bcs(mine)
            halo.exchange()
            oracle_pre_prolong(oracle, basis, mine)
            oracle.compute_euler(basis, mine, dt=dt, i_stage=stage)
    # Solver::is_admissible of a partitioned mesh = AND over the ranks' own checks
    assert allreduce_and(True) and not allreduce_and(rank != 1)
    mine_ok = oracle.is_admissible(mine)[0]
    assert allreduce_and(mine_ok) == allreduce_and(mine_ok) and (allreduce_and(mine_ok) <= mine_ok)
    np.save(result_path % rank, mine.elem_data)
    dist.barrier()
    dist.destroy_process_group()


def test_partitioned_gloo_world2(oracle, tmp_path):
    """two processes, torch.distributed/gloo: partition, halo send/recv, dt allreduce(min) -- the N > 1 host logic of bench.py"""
    import torch.multiprocessing as mp
    world = 2
    port = 29500 + os.getpid() % 2000
    result = str(tmp_path / "part_%d.npy")
    mp.spawn(_gloo_worker, args=(world, port, result), nprocs=world, join=True)
    rng = np.random.default_rng(3)
    basis, m = make_case("soup3d", rng)
    part = rng.integers(0, world, m.n_elem)
    ref, _ = reference_run(oracle, basis, m, 2)
    parts = P.partition_mesh(m, part, world)
    for r in range(world):
        assert np.array_equal(np.load(result % r), ref.elem_data[parts[r].global_elem])


@pytest.mark.parametrize("deformed", [True, False])
@pytest.mark.parametrize("blocks", [(2, 1, 1), (2, 2, 1), (2, 2, 2)])
def test_box_blocks_match_undivided(oracle, deformed, blocks):
    """bench.py builds one block of the global box per rank (box_mesh(blocks=, block=)); the blocks together must reproduce the
    undivided box bit for bit, metric terms on the cuts included"""
    nd, rs, n = 3, 3, 2
    basis = hb.gauss_legendre(rs)
    fs = freestream_state(nd)
    world = int(np.prod(blocks))
    assert M.proc_grid(world) == blocks
    if blocks == (2, 2, 2):
        whole = M.box_mesh(nd, rs, 2*n, basis, deformed=deformed, bc_kind=M.BC_FREESTREAM, bc_params=fs)
        density_wave(whole, basis)
        oracle.compute_write_face(basis, whole)
        ref, dts = reference_run(oracle, basis, whole, 2)
    parts = []
    for rank in range(world):
        m = M.box_mesh(nd, rs, n, basis, deformed=deformed, bc_kind=M.BC_FREESTREAM, bc_params=fs, blocks=blocks, block=M.block_coords(rank, blocks))
        density_wave(m, basis)
        oracle.compute_write_face(basis, m)
        parts.append(m)
    # every send list has a matching receive list of the same length on the peer
    for r, m in enumerate(parts):
        for peer, s in m.halo.send.items():
            assert len(parts[peer].halo.recv[r]) == len(s)
    ex = in_process_exchange(parts)
    for step in range(2):
        dt = oracle_step_parts(oracle, basis, parts, ex)
        if blocks == (2, 2, 2):
            assert dt == dts[step]
    if blocks == (2, 2, 2):
        g = 2*n
        for m in parts:
            gid = (m.elem_index*np.array([g*g, g, 1])).sum(-1)
            assert np.array_equal(m.elem_data, ref.elem_data[gid])
    else:
        # conservation across the cuts: both ranks computed the same flux, so the faces of a cut hold identical numbers
        for r, m in enumerate(parts):
            assert np.isfinite(m.elem_data).all()


# ---------------------------------------------------------------------------------------------------------------
# viscous stage: two exchanges (state faces before Neighbor, LDG viscous-flux faces before Neighbor_reconcile)
# ---------------------------------------------------------------------------------------------------------------
from pyoracle import NAVIER_STOKES  # noqa: E402
import pyoracle  # noqa: E402

VISC_O, COND_O = pyoracle.sutherland(1.7e-5, 273., 110.), pyoracle.constant(2.5e-2)


def _sub_refs(m):
    view = copy.copy(m)
    view.ref_face = np.ascontiguousarray(m.ref_face[m.pre_prolong])
    return view


def oracle_ns_first_half(oracle, basis, m, dt):
    """src/kernels_diffusive.cpp:10-18 kernel by kernel (what compute_navier_stokes_begin + _middle do on the device)"""
    NS = NAVIER_STOKES
    if len(getattr(m, "pre_prolong", ())):
        oracle.compute_prolong(basis, _sub_refs(m), pde=NS)
    for deformed in (0, 1):
        oracle.neighbor(NS, deformed, m, 0, VISC_O, COND_O)
    oracle.compute_restrict(basis, m, True, False, pde=NS)
    oracle.compute_restrict(basis, m, False, True, pde=NS)
    for deformed in (0, 1):
        oracle.local(NS, deformed, basis, m, VISC_O, COND_O, dt=dt, i_stage=0)
    oracle.compute_prolong(basis, m, True, True, pde=NS)
    oracle.apply_flux_bcs(m)


def oracle_ns_second_half(oracle, basis, m, dt):
    """src/kernels_diffusive.cpp:19-25"""
    NS = NAVIER_STOKES
    if len(getattr(m, "pre_prolong", ())):
        oracle.compute_prolong(basis, _sub_refs(m), True, True, pde=NS)
    for deformed in (0, 1):
        oracle.neighbor_reconcile(NS, deformed, m)
    oracle.compute_restrict(basis, m, True, True, pde=NS)
    for deformed in (0, 1):
        oracle.reconcile_ldg_flux(NS, deformed, basis, m, VISC_O, COND_O, dt=dt, i_stage=0)
    oracle.compute_prolong(basis, m, pde=NS)


def make_ns_case(kind, rng, oracle):
    from util import prepare_pde_state
    if kind == "soup2d":
        basis = hb.gauss_legendre(3)
        m = M.soup_mesh(2, 3, rng, n_car=8, n_def=14, n_ref=6, with_ldg=True)
        M.random_flow_state(m, rng)
        oracle.compute_prolong(basis, m)
    else:
        basis = hb.gauss_legendre(3)
        m = M.box_mesh(3, 3, 4, basis, deformed=True, bc_kind=M.BC_NONPENETRATION, with_ldg=True)
        density_wave(m, basis)
        oracle.compute_write_face(basis, m)
    prepare_pde_state(m, rng, NAVIER_STOKES)
    return basis, m


def ns_reference_step(oracle, basis, m, safety=0.3):
    ref = m.copy()
    dt = oracle.max_dt(NAVIER_STOKES, basis, ref, safety, safety, False, VISC_O, COND_O)
    oracle.apply_state_bcs(ref)
    oracle.compute_navier_stokes(basis, ref, lambda: oracle.apply_flux_bcs(ref), VISC_O, COND_O, dt=dt, i_stage=0)
    oracle.apply_state_bcs(ref)
    oracle.compute_euler(basis, ref, dt=dt, i_stage=1)
    return ref, dt


def ldg_exchange(parts):
    def get(p, slots):
        return parts[p].face_ldg[slots].copy()

    def put(q, slots, data):
        parts[q].face_ldg[slots] = data
    return lambda: exchange_in_process(parts, get, put)


@pytest.mark.parametrize("kind,n_parts", [("soup2d", 1), ("soup2d", 3), ("box_def", 2), ("box_def", 4)])
def test_partitioned_viscous_oracle_matches_undivided(oracle, kind, n_parts):
    """the kernel-by-kernel viscous stage with its two exchanges reproduces compute_navier_stokes on the undivided mesh bit for bit
    (n_parts = 1: the decomposition itself; > 1: arbitrary ownership, split hanging faces included)"""
    rng = np.random.default_rng(17)
    basis, m = make_ns_case(kind, rng, oracle)
    ref, dt = ns_reference_step(oracle, basis, m)
    part = rng.integers(0, n_parts, m.n_elem) if kind.startswith("soup") else P.split_by_curve(P.morton_keys(m.elem_index), n_parts)
    parts = P.partition_mesh(m, part, n_parts)
    ex0, ex1 = in_process_exchange(parts), ldg_exchange(parts)
    assert min(oracle.max_dt(NAVIER_STOKES, basis, p, 0.3, 0.3, False, VISC_O, COND_O) for p in parts) == dt
    for p in parts:
        oracle.apply_state_bcs(p)
    ex0()
    for p in parts:
        oracle_ns_first_half(oracle, basis, p, dt)
    ex1()
    for p in parts:
        oracle_ns_second_half(oracle, basis, p, dt)
    for p in parts:
        oracle.apply_state_bcs(p)
    ex0()
    for p in parts:
        oracle_pre_prolong(oracle, basis, p)
        oracle.compute_euler(basis, p, dt=dt, i_stage=1)
    out = m.copy()
    P.gather_elements(parts, out)
    assert np.array_equal(out.elem_data, ref.elem_data)


@pytest.mark.parametrize("kind,n_parts", [("soup2d", 3), ("box_def", 2)])
def test_partitioned_viscous_device_split(oracle, emu_lib, kind, n_parts):
    """compute_navier_stokes_begin / middle / finish of the CUDA sources with both exchanges, several parts in one process"""
    from hexed_b200.kernels import Device, sutherland, constant_transport
    from util import rel_l2
    rng = np.random.default_rng(17)
    basis, m = make_ns_case(kind, rng, oracle)
    ref, dt = ns_reference_step(oracle, basis, m)
    part = rng.integers(0, n_parts, m.n_elem) if kind.startswith("soup") else P.split_by_curve(P.morton_keys(m.elem_index), n_parts)
    parts = P.partition_mesh(m, part, n_parts)
    devs = [Device(m.n_dim, m.row_size, basis, lib_path=emu_lib).load_mesh(p) for p in parts]
    visc, cond = sutherland(1.7e-5, 273., 110.), constant_transport(2.5e-2)
    w = m.nv*m.nfq

    def exchange(face_kind):
        def get(p, slots):
            buf = np.empty((len(slots), w))
            devs[p].face_list_gather(devs[p].send_lists[[q for q, s in parts[p].halo.send.items() if s is slots][0]], buf, face_kind)
            devs[p].synchronize()
            return buf

        def put(q, slots, data):
            peer = [p for p, s in parts[q].halo.recv.items() if s is slots][0]
            devs[q].face_list_scatter(devs[q].recv_lists[peer], np.ascontiguousarray(data), face_kind)
            devs[q].synchronize()
        exchange_in_process(parts, get, put)
    dt_d = min(d.max_dt_navier_stokes(0.3, 0.3, False, visc, cond) for d in devs)
    assert abs(dt_d/dt - 1) <= 1e-13
    for d in devs:
        d.apply_state_bcs()
        d.compute_navier_stokes_begin(visc, cond, dt=dt, i_stage=0)
    exchange(0)
    for d in devs:
        d.compute_navier_stokes_middle(d.apply_flux_bcs, visc, cond, dt=dt, i_stage=0)
    exchange(1)
    for d in devs:
        d.compute_navier_stokes_finish(visc, cond, dt=dt, i_stage=0)
    for d in devs:
        d.apply_state_bcs()
        d.compute_euler_begin()
    exchange(0)
    for d in devs:
        d.compute_euler_finish(dt=dt, i_stage=1)
    for d, p in zip(devs, parts):
        d.sync_to_host(p)
        d.close()
    out = m.copy()
    P.gather_elements(parts, out)
    assert rel_l2(out.state(), ref.state()) <= 1e-11
    nf = 2*m.n_dim
    assert rel_l2(out.face_state[:nf*m.n_elem], ref.face_state[:nf*m.n_elem]) <= 1e-11
