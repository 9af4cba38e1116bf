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


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
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
