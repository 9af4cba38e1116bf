#!/usr/bin/env python3
"""N-GPU parity check of the partitioned path (run under torchrun on a multi-GPU box):
every rank advances its block of a deformed 3-D row-size-6 box on its GPU with NCCL halo exchange + dt allreduce; rank 0 repeats
the run with the CPU oracle on all blocks in process and compares. Prints one JSON line."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import hexed_b200 as hb  # noqa: E402
from hexed_b200 import mesh as M  # noqa: E402
from hexed_b200.cases import density_wave, freestream_state  # noqa: E402
from hexed_b200.halo import DeviceHalo, allreduce_min  # noqa: E402
from hexed_b200.kernels import Device  # noqa: E402


def refined_case(rank, world, local):
    """C5 class (onera_m6-like): a Cartesian 3-D row-size-6 box with hanging-node faces (`Refined_connection<Element>` layout), ownership drawn at
    random so that the partition cuts conforming faces, fine faces and coarse faces alike (`pre_prolong`); every rank builds the same global
    mesh, partitions it, keeps its part; NCCL halo exchange + dt allreduce; rank 0 repeats everything with the CPU oracle on the UNDIVIDED mesh."""
    from hexed_b200 import partition as P
    from pyoracle import Oracle, EULER
    nd, rs, n, steps = 3, 6, 4, 3
    basis = hb.gauss_legendre(rs)
    rng = np.random.default_rng(2024)
    refine = np.zeros((n,)*nd, bool)
    refine[1, 1, 1] = refine[3, 3, 3] = refine[0, 2, 3] = refine[2, 0, 1] = True
    m = M.refined_box_mesh(nd, rs, n, basis, refine, bc_kind=M.BC_COPY)
    M.random_flow_state(m, rng, mach=0.2)
    oracle = Oracle()
    oracle.compute_write_face(basis, m)
    oracle.compute_prolong(basis, m)
    owner = rng.integers(0, world, m.n_elem)
    parts = P.partition_mesh(m, owner, world)
    mine = parts[rank]
    dev = Device(nd, rs, basis, device=local).load_mesh(mine)
    halo = DeviceHalo(dev, mine)
    cuda = torch.device("cuda", local)
    dts = []
    for _ in range(steps):
        dt = allreduce_min(dev.max_dt_euler(0.3, 0.3, False), device=cuda)
        dts.append(dt)
        for stage in (0, 1):
            dev.apply_state_bcs()
            halo.start(); dev.compute_euler_begin(); halo.finish()
            dev.compute_euler_finish(dt=dt, i_stage=stage)
    dev.sync_to_host(mine)
    nq = mine.nq
    width = mine.nv*nq
    buf = np.zeros((max(p.n_elem for p in parts), width))
    buf[:mine.n_elem] = mine.state().reshape(mine.n_elem, width)
    t = torch.from_numpy(buf).to(cuda)
    gathered = [torch.empty_like(t) for _ in range(world)] if rank == 0 else None
    dist.gather(t, gathered, dst=0)
    if rank == 0:
        ref = m.copy()
        dt_errs = []
        for s in range(steps):
            dt_o = oracle.max_dt(EULER, basis, ref, 0.3, 0.3, False)
            dt_errs.append(abs(dts[s]/dt_o - 1))
            for stage in (0, 1):
                oracle.apply_state_bcs(ref)
                oracle.compute_euler(basis, ref, dt=dt_o, i_stage=stage)
        for r in range(world):
            parts[r].state()[:] = gathered[r].cpu().numpy()[:parts[r].n_elem].reshape(parts[r].state().shape)
        out = m.copy()
        P.gather_elements(parts, out)
        err = float(np.linalg.norm(out.state() - ref.state())/np.linalg.norm(ref.state()))
        per_elem = np.linalg.norm((out.state() - ref.state()).reshape(m.n_elem, -1), axis=1)/np.linalg.norm(ref.state().reshape(m.n_elem, -1), axis=1)
        ok = err <= 1e-11 and per_elem.max() <= 1e-11 and max(dt_errs) <= 1e-13
        print(json.dumps({"multigpu_check": "ok" if ok else "FAIL", "case": "refined box (hanging-node faces cut by a random partition)", "world": world,
                          "elements": int(m.n_elem), "refined_faces": int(m.ref_face.shape[0]), "elements_per_rank": [int(p.n_elem) for p in parts],
                          "pre_prolong_per_rank": [int(len(p.pre_prolong)) for p in parts], "cut_car_per_rank": [int(p.n_cut_car) for p in parts],
                          "state_rel_l2": err, "worst_element_rel_l2": float(per_elem.max()), "max_dt_rel_err": max(dt_errs),
                          "halo_bytes_per_exchange": halo.bytes_per_exchange, "steps": "%d Euler steps against the undivided oracle run" % steps}))
    dev.close()
    dist.barrier()
    dist.destroy_process_group()


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if "--refined" in sys.argv:
        return refined_case(rank, world, local)
    nd, rs, n, steps = 3, 6, 4, 3
    basis = hb.gauss_legendre(rs)
    fs = freestream_state(nd)
    blocks = M.proc_grid(world, nd)

    def build(r):
        m = M.box_mesh(nd, rs, n, basis, deformed=True, bc_kind=M.BC_FREESTREAM, bc_params=fs, blocks=blocks, block=M.block_coords(r, blocks), with_ldg=True)
        density_wave(m, basis)
        return m
    m = build(rank)
    dev = Device(nd, rs, basis, device=local).load_mesh(m)
    dev.compute_write_face()
    halo = DeviceHalo(dev, m)
    cuda = torch.device("cuda", local)
    dts = []
    for _ in range(steps):
        dt = allreduce_min(dev.max_dt_euler(0.7, 0.7, False), device=cuda)
        dts.append(dt)
        for stage in (0, 1):
            dev.apply_state_bcs()
            halo.start(); dev.compute_euler_begin(); halo.finish()
            dev.compute_euler_finish(dt=dt, i_stage=stage)
    # ---- one viscous step on top: compute_navier_stokes split around its two exchanges, then the inviscid second stage ----
    from hexed_b200.kernels import sutherland, constant_transport
    visc, cond = sutherland(1.7e-5, 273., 110.), constant_transport(2.5e-2)
    halo_ldg = DeviceHalo(dev, m, kind=1)
    dt_ns = allreduce_min(dev.max_dt_navier_stokes(0.3, 0.3, False, visc, cond), device=cuda)
    dev.apply_state_bcs()
    halo.start(); dev.compute_navier_stokes_begin(visc, cond, dt=dt_ns, i_stage=0); halo.finish()
    dev.compute_navier_stokes_middle(lambda: (dev.apply_flux_bcs(), halo_ldg.start()), visc, cond, dt=dt_ns, i_stage=0)
    halo_ldg.finish()
    dev.compute_navier_stokes_finish(visc, cond, dt=dt_ns, i_stage=0)
    dev.apply_state_bcs()
    halo.start(); dev.compute_euler_begin(); halo.finish()
    dev.compute_euler_finish(dt=dt_ns, i_stage=1)
    dev.sync_to_host(m)
    mine = torch.from_numpy(m.state().copy()).to(cuda)
    gathered = [torch.empty_like(mine) for _ in range(world)] if rank == 0 else None
    dist.gather(mine, gathered, dst=0)
    if rank == 0:
        from pyoracle import Oracle
        from test_partition import (oracle_step_parts, in_process_exchange, ldg_exchange, oracle_ns_first_half, oracle_ns_second_half,
                                    oracle_pre_prolong, VISC_O, COND_O)
        from pyoracle import NAVIER_STOKES
        oracle = Oracle()
        parts = [build(r) for r in range(world)]
        for p in parts:
            oracle.compute_write_face(basis, p)
        ex = in_process_exchange(parts)
        errs, dt_errs = [], []
        for s in range(steps):
            dt_o = oracle_step_parts(oracle, basis, parts, ex, safety=0.7)
            dt_errs.append(abs(dts[s]/dt_o - 1))
        ex1 = ldg_exchange(parts)
        dt_o = min(oracle.max_dt(NAVIER_STOKES, basis, p, 0.3, 0.3, False, VISC_O, COND_O) for p in parts)
        dt_errs.append(abs(dt_ns/dt_o - 1))
        for p in parts:
            oracle.apply_state_bcs(p)
        ex()
        for p in parts:
            oracle_ns_first_half(oracle, basis, p, dt_o)
        ex1()
        for p in parts:
            oracle_ns_second_half(oracle, basis, p, dt_o)
        for p in parts:
            oracle.apply_state_bcs(p)
        ex()
        for p in parts:
            oracle_pre_prolong(oracle, basis, p)
            oracle.compute_euler(basis, p, dt=dt_o, i_stage=1)
        for r in range(world):
            ref = parts[r].state()
            got = gathered[r].cpu().numpy()
            errs.append(float(np.linalg.norm(got - ref)/np.linalg.norm(ref)))
        ok = max(errs) <= 1e-11 and max(dt_errs) <= 1e-13
        print(json.dumps({"multigpu_check": "ok" if ok else "FAIL", "world": world, "blocks": blocks, "state_rel_l2_per_rank": errs,
                          "max_dt_rel_err": max(dt_errs), "halo_bytes_per_exchange": halo.bytes_per_exchange,
                          "steps": "%d Euler steps + 1 viscous step (Navier-Stokes stage 0 with both exchanges, Euler stage 1)" % steps}))
    dev.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
