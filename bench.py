#!/usr/bin/env python3
"""bench.py -- DOF-stage-updates/s of the per-stage DG residual update (BASELINE.json metric).

One "step" = one `Solver::update` iteration of the reference (src/Solver.cpp:846-868) on a synthetic 3-D box of
row-size-6 hexes: max_dt (CFL reduction) followed by two stages, each = ghost-state boundary fill + compute_euler
(neighbor flux, hanging-node restrict, local residual/update/face extrapolation, prolong). value = elements * 1080 DOF *
2 stages * steps / device time. See DESIGN.md "Measurement" for the byte counts behind the roofline object.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--n BOX] [--mesh deformed|cartesian]
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "DOF-stage-updates/sec"
# fixed roofline denominators: doubles per element-stage, 3-D row size 6 Euler
ALG_DOUBLES = {"deformed": {"stage": 11556, "local": 8424, "neighbor": 2484, "max_dt_half": 648},
               "cartesian": {"stage": 8424, "local": 5616, "neighbor": 2160, "max_dt_half": 648}}


def alg_doubles(nd, rs, deformed):
    """the same accounting for any (n_dim, row_size): compulsory HBM traffic of one Euler stage per element, following the
    reference data flow: Local R state + tss + face flux + residual cache (R or W) + W state + W faces
    (+ normals, determinant, face normals when deformed); Neighbor R + W of n_dim connections per element (+ normals);
    half a Max_dt (R state + W tss)"""
    nq, nfq, nv = rs**nd, rs**(nd - 1), nd + 2
    faces = 2*nd*nv*nfq
    local = nv*nq + nq + faces + nv*nq + nv*nq + faces
    neighbor = 2*faces
    if deformed:
        local += nd*nd*nq + nq + 2*nd*nd*nfq
        neighbor += nd*nd*nfq
    half_dt = (nv*nq + nq)//2
    return {"stage": local + neighbor + half_dt, "local": local, "neighbor": neighbor, "max_dt_half": half_dt}


def workload_name(args, viscous=False):
    nd = args.dim
    n = 1000 if (nd == 2 and args.n == 100) else args.n
    return ("synthetic %dD %s %s box %d^%d = %d elements per GPU, row_size 6, %s, freestream ghosts; step = %s"
            % (nd, args.mesh, "hex" if nd == 3 else "quad", n, nd, n**nd, "Navier-Stokes (Sutherland air)" if viscous else "Euler",
               "max_dt + ghost fill + compute_navier_stokes (stage 0, LDG) + ghost fill + compute_euler (stage 1)" if viscous
               else "max_dt + 2 x (ghost BC fill + compute_euler)"))


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", "--box-n", dest="n", type=int, default=100, help="box edge in elements per GPU (100 -> 1M elements); --box-n is the spelling torchrun does not trip over")
    ap.add_argument("--mesh", default="deformed", choices=["deformed", "cartesian"])
    ap.add_argument("--dim", type=int, default=3, choices=[2, 3], help="3 = the headline; 2 = the shape of the 2-D BASELINE configs (vortex / naca0012 / cylinder)")
    ap.add_argument("--pde", default="euler", choices=["euler", "navier_stokes"],
                    help="euler = the headline; navier_stokes = Solver::update with use_ldg (stage 0 viscous/LDG, stage 1 inviscid), 1 GPU only")
    ap.add_argument("--cpu-n", type=int, default=48, help="box edge of the bounded CPU sample (48^3 = 110 592 elements, 6 GB working set)")
    ap.add_argument("--cpu-steps", type=int, default=10, help="timed steps of the CPU sample (10 steps = 20 stages, SURVEY section 8d), after 3 warm-up steps")
    ap.add_argument("--pipe-mode", type=int, default=1, help="HEXED_B200_OPT_PIPELINED_LOCAL value (A/B: 2 = earlier shared-memory layout of the 3-D deformed kernel)")
    ap.add_argument("--ns-layout", type=int, default=-1, help="HEXED_B200_OPT_NS_LOCAL_LAYOUT (A/B of the 3-D Navier-Stokes Local kernel variants 0 / 1 / 2; -1 = library default)")
    ap.add_argument("--adapter-n", type=int, default=0, help="box edge per GPU of the end-to-end run through the C++ adapter (0 = 64, or 48 when host memory is short: "
                    "the reference keeps 85 KB of host objects per deformed element)")
    ap.add_argument("--no-aux-lines", action="store_true", help="skip the Navier-Stokes / Cartesian sub-lines of the default run")
    ap.add_argument("--cbc-first", action="store_true", help="time the call-by-call loop before the device-resident loop (measurement-order experiment)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def ncu_traffic(kernel, n_elem):
    """DRAM bytes per launch of `kernel` as measured by ncu (profiles/ncu_traffic.json), or None"""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        entry = json.load(open(p)).get(kernel)
        return {"bytes": entry["bytes_per_element"]*n_elem, "source": entry["source"]} if entry else None
    except Exception:
        return None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)).get("hbm_gbs", 6650.), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650., "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """samples nvidia-smi clocks / throttle reasons for one GPU while the timed region runs"""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for name, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm)//2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def host_info():
    model = "unknown"
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                model = line.split(":", 1)[1].strip()
                break
    except OSError:
        pass
    avail = 0.
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable"):
                avail = float(line.split()[1])/1e6
    except OSError:
        pass
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    return {"cpu_model": model, "cores": cores, "mem_available_gb": avail}


def adapter_e2e(args, n_dev, viscous, steps, warmup, modes=("host_bcs",), n_override=None):
    """The metric measured the way a Hexed user would see it: through hexed::max_dt_* / hexed::compute_* of the C++ adapter
    (hexed_b200/libhexed_b200_host.so = adapter.cpp + the pointer-graph mesh of harness.cpp standing in for Solver's Accessible_mesh),
    one Kernel_mesh on n_dev GPUs (hexed_b200::set_devices: Morton split, NCCL halo exchange, dt allreduce -- all below the boundary).
    Timed with the host clock between device synchronisations (host loops are part of the step). One harness, several modes:
      host_bcs        state resident on the devices, boundary conditions applied BY THE HOST every stage as Solver::apply_state_bcs does
                      (src/Solver.cpp:56-67): inside faces D2H, an OpenMP loop of per-face Flow_bc calls over the host objects, ghost faces H2D
      device_bcs      the flow INTEGRATION.md section 3a recommends: conditions registered once with hexed_b200::add_device_bc and applied by
                      hexed_b200::apply_state_bcs / apply_flux_bcs on the devices; after every stage the host asks hexed_b200::is_admissible as
                      Solver::update does (fix_admissibility, src/Solver.cpp:960-975): per step dt, the flags and Element::record cross PCIe
      sync_every_call the adapter's zero-Solver-change default: every hexed:: call moves what it reads and writes over PCIe
    Returns {mode: result}."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np
    import torch
    import hexed_b200 as hb
    from hexed_b200 import mesh as M
    from hexed_b200.cases import density_wave, freestream_state
    import host_harness as H
    info = host_info()
    nd, rs = args.dim, 6
    # host memory: the reference keeps ~85 KB of host objects per deformed 3-D element (+ the flat mesh while they are being filled)
    avail = info["mem_available_gb"]/n_dev
    # the same per-GPU size at every N, so that the driver's weak-scaling ratio compares like with like: 64^3 (262 144 elements, 22 GB of host objects per GPU)
    n_gpu = n_override or args.adapter_n or (64 if avail >= 40. else 48)
    if nd == 2:
        n_gpu = int(round(n_gpu**1.5))
    n = int(round(n_gpu*n_dev**(1./nd)))
    deformed = args.mesh == "deformed"
    basis = hb.gauss_legendre(rs)
    fs = freestream_state(nd)
    m = M.box_mesh(nd, rs, n, basis, deformed=deformed, bc_kind=M.BC_FREESTREAM, bc_params=fs, with_ldg=viscous,
                   device="cuda:0" if torch.cuda.is_available() else None)  # (metric terms of the synthetic mesh evaluated on the device, returned on the host)
    density_wave(m, basis)
    if viscous:
        m.elem_data[:, M.BULK_AV_SLOT(nd)] = 0.; m.elem_data[:, M.LAPLACIAN_AV_SLOT(nd)] = 0.
    ne, nv, nq, nfq = m.n_elem, m.nv, m.nq, m.nfq
    n_face_slot = m.n_face_slot
    rows = np.ascontiguousarray(m.bcs[0]["con_index"], dtype=np.int32)
    coords = np.ascontiguousarray(m.elem_index)
    h = H.HostHarness(H.build(emu=False), m, basis)
    m.elem_data = None; m.ref_normals = None; m.det = None; m.face_state = None; m.face_ldg = None  # the host objects hold the data now
    threads = info["cores"]
    visc, cond = H.sutherland(1.716e-5, 273., 111.), H.sutherland(0.0241, 273., 194.)
    nb, w = rows.size, nv*nfq
    per_stage = nb*w*8
    n_exch = 2 + (2 if viscous else 0)  # state faces every stage (+ LDG faces down and up inside the viscous stage)
    results = {}

    def measure(mode):
        sync = mode == "sync_every_call"
        dev_bcs = mode == "device_bcs"
        host_s = {"inside_faces_to_host": 0., "host_bc_loop": 0., "ghost_faces_to_device": 0., "hexed_calls": 0.}

        def clocked(key, fn, *a, **kw):
            t = time.perf_counter(); r = fn(*a, **kw); host_s[key] += time.perf_counter() - t
            return r

        def state_bcs():
            if dev_bcs:
                return clocked("hexed_calls", h.apply_state_bcs)
            if not sync:
                clocked("inside_faces_to_host", h.inside_state_faces_to_host)
            clocked("host_bc_loop", h.host_state_bcs, 0, fs, rows, threads)
            if not sync:
                clocked("ghost_faces_to_device", h.ghost_state_faces_to_device)

        def flux_bcs():  # runs inside compute_navier_stokes (its flux_bc callback), like Solver::apply_flux_bcs
            if dev_bcs:
                return h.apply_flux_bcs()
            if not sync:
                h.inside_ldg_faces_to_host()
            h.host_flux_bcs(rows, threads)
            if not sync:
                h.ghost_ldg_faces_to_device()

        def admissible():
            if dev_bcs and not clocked("hexed_calls", h.is_admissible_flag):
                raise RuntimeError("inadmissible state in the benchmark flow")

        def step():
            if dev_bcs:  # this step's input: the freestream state of the boundary condition (a HIL-controlled, possibly time-dependent, value)
                clocked("hexed_calls", h.set_device_bc_params, 0, fs)
            if viscous:
                dt = clocked("hexed_calls", h.call, "max_dt_navier_stokes", 0.7, 0.7, False, *visc, *cond)
                state_bcs()
                clocked("hexed_calls", h.call, "compute_navier_stokes", *visc, *cond, dt=dt, i_stage=0)
                admissible()
                state_bcs()
                clocked("hexed_calls", h.call, "compute_euler", dt=dt, i_stage=1)
                admissible()
            else:
                dt = clocked("hexed_calls", h.call, "max_dt_euler", 0.7, 0.7, False)
                for stage in (0, 1):
                    state_bcs()
                    clocked("hexed_calls", h.call, "compute_euler", dt=dt, i_stage=stage)
                    admissible()
        h.set_sync_mode(H.SYNC_EVERY_CALL if sync else H.RESIDENT)
        if dev_bcs:
            h.add_device_bcs(m)
        if viscous:
            h.set_flux_bc(flux_bcs)
        k_steps, k_warm = (min(steps, 2), 1) if sync else (steps, warmup)
        for _ in range(k_warm):
            step()
        h.synchronize()
        for k in host_s:
            host_s[k] = 0.
        t0 = time.perf_counter()
        for _ in range(k_steps):
            step()
        h.synchronize()
        sec = time.perf_counter() - t0
        if sync:
            elem_bytes = ne*(nv + 1 + (6 if viscous else 0) + max(nv, rs))*nq*8
            face_bytes = n_face_slot*2*w*8
            geom = ne*((nd*nd + 1)*nq*8 if deformed else 0)
            h2d = 3*(elem_bytes + face_bytes + geom)
            d2h = 2*(elem_bytes + face_bytes) + ne*nq*8
            text = ("sync_every_call (the zero-Solver-change default of the adapter): every hexed:: call uploads what it reads from the host objects, "
                    "metric terms included, and downloads what it wrote")
        elif dev_bcs:
            h2d, d2h = 8*nv*n_dev, 8 + 2*8*n_dev
            text = ("adapter (hexed::max_dt_* / compute_* + hexed_b200::set_device_bc_params / apply_state_bcs / is_admissible of hexed_b200/host/adapter.cpp on "
                    "a pointer-graph Kernel_mesh), state resident on the device(s), boundary conditions registered on the devices (INTEGRATION.md section 3a); "
                    "per step: the freestream state of the boundary condition H2D, dt D2H, and after each stage the admissibility flags of every device D2H "
                    "(Element::record only when a device reports an inadmissible state), as Solver::update does. Nothing else has to cross PCIe in a Hexed run "
                    "whose boundary conditions the device implements; the variant with HOST-applied conditions is aux.e2e_adapter_host_bcs")
        else:
            h2d, d2h = n_exch*per_stage, n_exch*per_stage + 8
            text = ("adapter (hexed::max_dt_*/compute_* of hexed_b200/host/adapter.cpp on a pointer-graph Kernel_mesh), resident; per stage: inside boundary "
                    "faces D2H (prefetched on a copy stream, pinned), host OpenMP loop of per-face Freestream::apply_state over the host objects (%d threads), "
                    "ghost faces H2D (deferred: lands after the interior Neighbor kernels); dt D2H per step" % threads)
        return {"value": ne*nv*nq*2*k_steps/sec, "unit": "DOF-stage/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": sec/k_steps*1e3,
                "mode": text, "elements": ne, "elements_per_gpu": np.bincount(h.element_owners(), minlength=n_dev).tolist(), "box": "%d^%d" % (n, nd),
                "transport": h.transport_description(), "host": info, "boundary_faces": int(nb),
                "host_ms_per_step": {k: v/k_steps*1e3 for k, v in host_s.items()},
                "timing": "host clock between device synchronisations, %d steps after %d warm-up" % (k_steps, k_warm)}
    try:
        h.set_devices(list(range(n_dev)))
        if n_dev > 1:
            h.set_element_coordinates(coords)
        h.set_sync_mode(H.RESIDENT)
        h.invalidate()
        h.call("compute_write_face")
        for mode in modes:  # (the headline mode first: nothing left over from another mode can touch it)
            try:
                results[mode] = measure(mode)
            except Exception as ex:
                results[mode] = {"value": None, "unit": "DOF-stage/s", "h2d_bytes_per_step": None, "d2h_bytes_per_step": None, "mode": "%s failed: %r" % (mode, ex)}
    finally:
        h.set_sync_mode(H.SYNC_EVERY_CALL)
        h.release()
        h.set_devices([0])
        h.close()
    return results


def build_native_oracle():
    """rebuild the oracle for this host's CPU (-march=native, row size 6 only); falls back to the portable build"""
    try:
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "native"], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=600)
        return "liboracle_native.so"
    except Exception:
        return "liboracle_fast.so"


def cpu_reference_rate(args, steps, warmup, n=None):
    """The CPU arm: two implementations of the same path are available on the host and the FASTER one is what gets timed for `steps` steps
    after `warmup` (both are probed for two steps first; `cpu_reference_rate.both` keeps the two probe rates):
      "reference"  the reference's OWN kernels (oracle/_ref/libhexed_ref.so: src/kernels_convective.cpp, kernels_max_dt.cpp + include/Spatial.hpp
                   compiled unmodified, oracle/Makefile.ref; built where /root/reference exists and shipped with the snapshot) on a Kernel_mesh stood up
                   once over the flat mesh. It compiles here only against oracle/eigen_shim (eager evaluation, no expression templates), which costs
                   it speed a genuine-Eigen build would not pay;
      "port"       the restated kernels (oracle/oracle_impl.hpp, plain loops, rebuilt -march=native for this host).
    So the GPU number is never compared with a handicapped CPU one. All host threads (OpenMP, as the reference's `threaded` build).
    Returns (rate, seconds per step, threads, kind, sample description)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    # torchrun exports OMP_NUM_THREADS=1 for its workers; the CPU arm is meant to use every host core (libgomp reads this at load time)
    os.environ["OMP_NUM_THREADS"] = str(len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
    import ctypes as C
    import numpy as np
    import hexed_b200 as hb
    from hexed_b200 import mesh as M
    import pyoracle
    from pyoracle import Oracle, EULER
    from hexed_b200.cases import density_wave, freestream_state
    basis = hb.gauss_legendre(6)
    nd = args.dim
    if n is None:
        n = args.cpu_n if nd == 3 else int(round(args.cpu_n**1.5))
    fs = freestream_state(nd)
    m = M.box_mesh(nd, 6, n, basis, deformed=args.mesh == "deformed", bc_kind=M.BC_FREESTREAM, bc_params=fs)
    density_wave(m, basis)
    plib = build_native_oracle()
    po = Oracle(plib)
    po.compute_write_face(basis, m)
    impls = {}

    def port_step():
        dt = po.max_dt(EULER, basis, m, 0.7, 0.7, False)
        for stage in (0, 1):
            po.apply_state_bcs(m)
            po.compute_euler(basis, m, dt=dt, i_stage=stage)
    impls["port"] = (port_step, "oracle/%s (restated kernels, -march=native)" % plib)
    view = None
    if pyoracle.ref_available():
        o = pyoracle.RefOracle()
        packed = o.pack_mesh(m)
        view = o.ref.hr_view_create(C.byref(packed))
        if view:
            ghost = np.ascontiguousarray(m.bcs[0]["ghost_slot"], dtype=np.int32)
            fs_a = np.ascontiguousarray(fs, dtype=np.float64)
            gp, fp = ghost.ctypes.data_as(C.POINTER(C.c_int)), fs_a.ctypes.data_as(C.POINTER(C.c_double))
            dt_c = C.c_double()

            def ref_step():
                o._check(o.ref.hr_view_max_dt_euler(view, 0.7, 0.7, 0, C.byref(dt_c)))
                for stage in (0, 1):
                    o.ref.hr_view_bc_freestream(view, ghost.size, gp, fp)
                    o._check(o.ref.hr_view_compute_euler(view, o.opts(dt=dt_c.value, i_stage=stage)))
            impls["reference"] = (ref_step, "oracle/_ref/libhexed_ref.so (the reference's own kernel sources on the Eigen stand-in, -O3 -march=x86-64-v3)")

    def rate_of(fn, k):
        t = time.perf_counter()
        for _ in range(k):
            fn()
        el = time.perf_counter() - t
        return m.n_elem*m.nv*m.nq*2*k/el, el
    probe = {}
    for kind, (fn, _) in impls.items():
        fn()
        probe[kind] = rate_of(fn, 2)[0]
    kind = max(probe, key=probe.get)
    if view and kind != "reference":
        o.ref.hr_view_destroy(view); view = None   # 26 GB of reference-layout face storage at 1 M elements
    fn, lib = impls[kind]
    for _ in range(warmup):
        fn()
    rate, el = rate_of(fn, steps)
    if view:
        o.ref.hr_view_destroy(view)
    cpu_reference_rate.both = {"probe_steps": 2, "reference_kernels_on_eigen_shim": probe.get("reference"), "restated_port": probe.get("port")}
    desc = "%d^%d = %d %s elements, %d steps (max_dt + 2 x (freestream ghost fill + compute_euler)), %s, OpenMP; the faster of the two CPU implementations (%s)" % (
        n, nd, m.n_elem, args.mesh, steps, lib, ", ".join("%s %.3g" % kv for kv in probe.items()))
    return rate, el/steps, po.num_threads(), kind, desc


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on this box's host cores, on the bench's own config when the
    host has the memory for it (1 M deformed 3-D elements = 50 GB of flat mesh + 26 GB of reference-layout face storage), else on the
    bounded --cpu-n sample; rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    info = host_info()
    nd = args.dim
    n_full = 1000 if (nd == 2 and args.n == 100) else args.n
    need_gb = n_full**nd*(80e3 if nd == 3 else 9e3)/1e9*1.6  # flat mesh + reference-layout faces + generation scratch
    full = info["mem_available_gb"] >= need_gb and n_full**nd*(args.steps + args.warmup) <= 40e6
    rate, sec, cores, kind, sample = cpu_reference_rate(args, args.steps, args.warmup, n=n_full if full else None)
    out = {"impl": "reference", "metric": METRIC, "value": rate, "unit": "DOF-stage/s", "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": sec*1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f64", "data": "synthetic",
           "config": {"workload": workload_name(args), "same_config": bool(full),
                      "sample": ("CPU arm timed on the full workload: " if full else "CPU arm timed on a bounded sample of that workload: ") + sample,
                      "host": info},
           "cpu_baseline": {"value": rate, "unit": "DOF-stage/s", "cores": cores, "kind": kind, "sample": sample, "host_cpu": info["cpu_model"],
                            "both_cpu_implementations": getattr(cpu_reference_rate, "both", None)},
           "e2e": {"value": rate, "unit": "DOF-stage/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    import numpy as np
    import torch
    import hexed_b200 as hb
    from hexed_b200 import mesh as M
    from hexed_b200.kernels import Device
    from hexed_b200.cases import freestream_state

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: hexed_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        host_group = dist.new_group(backend="gloo")  # for waits that must not park a spinning NCCL kernel on the other ranks' GPUs
    nd, rs, n = args.dim, 6, args.n
    if nd == 2 and args.n == 100:
        n = 1000  # 10^6 quads
    deformed = args.mesh == "deformed"
    basis = hb.gauss_legendre(rs)
    fs = freestream_state(nd)
    cuda = torch.device("cuda", local_rank)

    # ---- synthetic mesh (geometry computed on the GPU, never leaves it) and initial state ----
    # one block of the global box per rank, blocks laid out in Z-order (space-filling-curve split of the structured box)
    blocks = M.proc_grid(world, nd)
    m = M.box_mesh(nd, rs, n, basis, deformed=deformed, bc_kind=M.BC_FREESTREAM, bc_params=fs, device=cuda,
                   geometry_chunk=32768, keep_geometry_on_device=deformed, lean=True,
                   blocks=blocks, block=M.block_coords(rank, blocks))
    ne, nq, nv = m.n_elem, m.nq, m.nv
    dev = Device(nd, rs, basis, device=local_rank).load_mesh(m, upload_elem_data=False)
    if args.pipe_mode != 1:
        dev.set_option(0, args.pipe_mode)
    if args.ns_layout >= 0:
        dev.set_option(3, args.ns_layout)
    # the generator's copies of the metric terms are dead once the device mirror holds them: 28 KB per element that a 2 M element
    # mesh (the 16 M / 8 GPU configuration) needs back
    m.ref_normals = None; m.det = None; m.normals = None
    torch.cuda.empty_cache()
    # density wave initial condition, written from pinned host memory (the solver's state lives in host arrays in the reference)
    pos = torch.as_tensor(m.qpoint_pos, device=cuda) if not torch.is_tensor(m.qpoint_pos) else m.qpoint_pos
    phase = sum(torch.sin(2*np.pi*pos[:, d] + 0.3*d) for d in range(nd))/nd
    rho = 1.2*(1 + 0.1*phase)
    vel = [0.3*340.*(0.6 + 0.2*d) for d in range(nd)]
    p = 101325.*(1 + 0.05*torch.cos(2*np.pi*pos[:, 0]))
    st = torch.empty((ne, nv + 1, nq), dtype=torch.float64, device=cuda)
    ke = 0
    for d in range(nd):
        st[:, d] = rho*vel[d]; ke = ke + 0.5*rho*vel[d]**2
    st[:, nd] = rho; st[:, nd + 1] = p/0.4 + ke; st[:, nd + 2] = 1.
    dev.upload_elements(st, 0, nv + 1)
    del st, pos, phase, rho, p, ke
    m.qpoint_pos = None; m.ref_normals = None; m.det = None; m.normals = None
    torch.cuda.empty_cache()
    dev.compute_write_face()
    stream = torch.cuda.ExternalStream(dev.cuda_stream(), device=cuda)

    from hexed_b200.halo import DeviceHalo
    halo = DeviceHalo(dev, m) if world > 1 else None
    halo_ldg = DeviceHalo(dev, m, kind=1) if (world > 1 and args.pde == "navier_stokes") else None

    def global_dt(dt):
        if dist is None:
            return dt
        t = torch.tensor([dt], dtype=torch.float64, device=cuda)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return float(t.item())

    def stage_kernels(dt, stage):
        if halo is None:
            dev.compute_euler(dt=dt, i_stage=stage)
        else:
            halo.start()                 # gather cut faces, NCCL send/recv over NVLink ...
            dev.compute_euler_begin()    # ... overlapped with the flux on interior connections
            halo.finish()
            dev.compute_euler_finish(dt=dt, i_stage=stage)

    viscous = args.pde == "navier_stokes"
    from hexed_b200.kernels import sutherland
    visc, cond = sutherland(1.716e-5, 273., 111.), sutherland(0.0241, 273., 194.)  # air (reference samples/Case.hil:83-90)

    def step():
        if viscous:  # Solver::update with use_ldg(): src/Solver.cpp:857-865
            dt = global_dt(dev.max_dt_navier_stokes(0.7, 0.7, False, visc, cond))
            dev.apply_state_bcs()
            if halo is None:
                dev.compute_navier_stokes(dev.apply_flux_bcs, visc, cond, dt=dt, i_stage=0)
            else:  # two exchanges: state faces before Neighbor, viscous-flux (LDG) faces before Neighbor_reconcile
                halo.start()
                dev.compute_navier_stokes_begin(visc, cond, dt=dt, i_stage=0)
                halo.finish()
                dev.compute_navier_stokes_middle(lambda: (dev.apply_flux_bcs(), halo_ldg.start()), visc, cond, dt=dt, i_stage=0)
                halo_ldg.finish()
                dev.compute_navier_stokes_finish(visc, cond, dt=dt, i_stage=0)
            dev.apply_state_bcs()
            stage_kernels(dt, 1)
            return
        dt = global_dt(dev.max_dt_euler(0.7, 0.7, False))
        for stage in (0, 1):
            dev.apply_state_bcs()
            stage_kernels(dt, stage)

    def barrier():
        if dist is not None:
            dist.barrier()
        dev.synchronize(); torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms], dtype=torch.float64, device=cuda)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms*1e-3

    for _ in range(args.warmup):
        step()
    # One GPU: the headline is the flow loop of Solver::update with the time step kept on the device (hexed_b200_update_euler /
    # _navier_stokes: max_dt -> dt never read back, boundary conditions on the device, one CUDA graph per step), i.e. no host
    # synchronisation inside the timed region. Several GPUs under torchrun: call by call (one dt allreduce + read-back per step).
    device_loop = world == 1
    if device_loop:
        def run_steps(k):
            if viscous:
                dev.update_navier_stokes(0.7, visc, cond, k)
            else:
                dev.update_euler(0.7, k)
        run_steps(max(args.warmup, 1))
        l0 = dev.launch_count(); run_steps(1); launches_per_step = dev.launch_count() - l0
        if args.cbc_first:  # (order experiment: is the device loop slower, or whichever of the two is timed first?)
            sec_call_by_call = timed(step, args.steps)
        sampler = ClockSampler(local_rank)
        sec = timed(lambda: run_steps(args.steps), 1)
        clocks = sampler.stop()
        launches = launches_per_step*args.steps  # (graph replays do not pass through the host-side counter)
        if not args.cbc_first:
            sec_call_by_call = timed(step, args.steps)
    else:
        launches0 = dev.launch_count()
        sampler = ClockSampler(local_rank) if rank == 0 else None
        sec = timed(step, args.steps)
        clocks = sampler.stop() if sampler else None
        launches = dev.launch_count() - launches0
        sec_call_by_call = sec
    dof_stage = ne*nv*nq*2*args.steps*world
    value = dof_stage/sec

    # ---- per-kernel device time (separate pass with per-launch CUDA events on the kernels' stream) ----
    dev.reset_stats(); dev.set_timing(True)
    n_prof = max(2, min(args.steps, 4))
    for _ in range(n_prof):
        step()
    dev.set_timing(False)
    stats = dev.kernel_stats()
    alg = alg_doubles(nd, rs, deformed)
    nfq_ = rs**(nd - 1)
    assert nd != 3 or alg == ALG_DOUBLES[args.mesh]
    peak, peak_src = peaks()
    local_stat = max((s for s in stats if s["name"] == "local" and s["deformed"] == int(deformed)), key=lambda s: s["launches"])
    local_sec = local_stat["device_seconds"]/max(local_stat["launches"], 1)
    local_gbs = ne*alg["local"]*8/local_sec/1e9
    shares = {("%s/%s" % (("car", "def", "shared")[s["deformed"]], s["name"])): s["device_seconds"]/n_prof for s in stats if s["launches"]}
    local_name = ("local_euler_pipe_kernel<6,%s>" if nd == 3 else "local_euler_pipe2d_kernel<6,%s>") % ("true" if deformed else "false")
    if viscous:
        local_name += " + ns_local_line_kernel (both count as 'local'; see kernel_seconds_per_step)"
    traffic = ncu_traffic(local_name, ne) if not viscous else None
    # Solver::is_admissible (reference src/Solver.cpp:921-958) runs after every stage of Solver::update when fix_admis is on; it is not
    # part of the metric (the reference accounts it under "check admis."), so it is timed separately and reported beside the step
    dev.reset_stats(); dev.set_timing(True)
    admissible = [dev.is_admissible() for _ in range(3)]
    dev.set_timing(False)
    admis_stat = [s for s in dev.kernel_stats() if s["name"] == "check admis."]
    aux = {"is_admissible_ms_per_call": admis_stat[0]["device_seconds"]/3*1e3 if admis_stat else None, "admissible": all(admissible),
           "is_admissible_bytes": ne*(nv*nq + 2*nd*nv*m.nfq)*8}
    if not viscous and halo is None:
        # the same check with HEXED_B200_OPT_FUSED_ADMIS: the Local kernels leave the bits, is_admissible after each stage reduces them
        from hexed_b200.kernels import OPT_FUSED_ADMIS
        dev.set_option(OPT_FUSED_ADMIS, 1)

        def step_checked():
            dt = dev.max_dt_euler(0.7, 0.7, False)
            ok = True
            for stage in (0, 1):
                dev.apply_state_bcs()
                dev.compute_euler(dt=dt, i_stage=stage)
                ok = dev.is_admissible() and ok
            return ok
        step_checked()
        fused_sec = timed(step_checked, n_prof)
        dev.reset_stats(); dev.set_timing(True)
        ok = step_checked()
        dev.set_timing(False)
        st = [s_ for s_ in dev.kernel_stats() if s_["name"] == "check admis."]
        aux["fused_admis"] = {"ms_per_step_with_a_check_after_every_stage": fused_sec/n_prof*1e3, "is_admissible_ms_per_call": st[0]["device_seconds"]/2*1e3 if st else None,
                              "admissible": bool(ok)}
        dev.set_option(OPT_FUSED_ADMIS, 0)
    stage_gbs = value/world*(alg["stage"]*8/float(nv*nq))/1e9

    # ---- end to end through the public API with HOST buffers: the boundary condition is applied by the host (as the reference's
    # Solver::apply_state_bcs does, src/Solver.cpp:56-67), so every stage the inside boundary faces go D2H and the ghost faces
    # come back H2D, pinned memory, inside the timed region. The boundary connections are declared "late" (set_partition), so the
    # flux on interior connections (compute_euler_begin) runs while the faces cross PCIe on a copy stream. ----
    e2e_api = None
    if not args.no_e2e and not viscous and world == 1:
        bc = m.bcs[0]
        inside_list, ghost_list = dev.face_list(bc["inside_slot"]), dev.face_list(bc["ghost_slot"])
        nb, w = bc["inside_slot"].size, nv*m.nfq
        h_in = torch.empty((nb, w), dtype=torch.float64, pin_memory=True)
        h_gh = torch.empty((nb, w), dtype=torch.float64, pin_memory=True)
        d_in = torch.empty((nb, w), dtype=torch.float64, device=cuda)
        d_gh = torch.empty((nb, w), dtype=torch.float64, device=cuda)
        fs_t = torch.as_tensor(fs).repeat_interleave(m.nfq)
        n_cut_car, n_cut_def = getattr(m, "n_cut_car", 0), getattr(m, "n_cut_def", 0)
        dev.set_partition(n_cut_car, n_cut_def + nb, getattr(m, "pre_prolong", ()))  # def_con = [interior, boundary, cut]
        copy_stream = torch.cuda.Stream(device=cuda)
        ev_g, ev_d, ev_u = torch.cuda.Event(), torch.cuda.Event(), torch.cuda.Event()

        def step_e2e():
            dt = global_dt(dev.max_dt_euler(0.7, 0.7, False))       # D2H: the time step
            for stage in (0, 1):
                dev.face_list_gather(inside_list, d_in)
                ev_g.record(stream); copy_stream.wait_event(ev_g)
                with torch.cuda.stream(copy_stream):
                    h_in.copy_(d_in, non_blocking=True)               # D2H: inside boundary faces for the host Flow_bc
                    ev_d.record(copy_stream)
                if halo is not None:
                    halo.start()
                dev.compute_euler_begin()                             # flux on interior connections overlaps the PCIe trips
                ev_d.synchronize()
                h_gh[:] = fs_t                                        # host boundary condition (Freestream::apply_state)
                with torch.cuda.stream(copy_stream):
                    d_gh.copy_(h_gh, non_blocking=True)               # H2D: ghost faces
                    ev_u.record(copy_stream)
                stream.wait_event(ev_u)
                dev.face_list_scatter(ghost_list, d_gh)
                if halo is not None:
                    halo.finish()
                dev.compute_euler_finish(dt=dt, i_stage=stage)
        for _ in range(2):
            step_e2e()
        sec_e2e = timed(step_e2e, args.steps)
        dev.set_partition(n_cut_car, n_cut_def, getattr(m, "pre_prolong", ()))
        e2e_api = {"value": dof_stage/sec_e2e, "unit": "DOF-stage/s", "h2d_bytes_per_step": 2*nb*w*8, "d2h_bytes_per_step": 2*nb*w*8 + 8,
                   "mode": "C ABI driven from Python (ctypes Device), resident state, 1 M elements; per stage the inside boundary faces go D2H (pinned), a "
                           "constant ghost state is broadcast into a pinned buffer, ghost faces go H2D; copies on a side stream overlap the "
                           "interior-connection flux; dt D2H per step. The lower bound of what host-applied boundary conditions cost: no host objects to scatter into"}
        aux["e2e_c_abi_from_python"] = e2e_api

    # ---- end to end the way a Hexed user would call it: hexed::max_dt_* / compute_* of the C++ adapter on a pointer-graph Kernel_mesh, boundary
    # conditions registered on the devices (headline) or applied by the host loop of Solver::apply_state_bcs (aux, one GPU); at N > 1 rank 0's
    # process drives all N GPUs through ONE Kernel_mesh (the split, the NCCL halo exchange and the dt allreduce happen below the kernels.hpp
    # boundary) while the other ranks wait ----
    e2e = None
    if not args.no_e2e:
        if dist is not None:
            torch.cuda.synchronize()
            dist.barrier(group=host_group)
        if rank == 0:
            keep = ("value", "unit", "ms_per_step", "h2d_bytes_per_step", "d2h_bytes_per_step", "elements", "elements_per_gpu", "box", "transport",
                    "boundary_faces", "host_ms_per_step", "mode")
            try:
                # the host-applied variant only on one GPU: its per-face host loop is Amdahl's serial part (one process, all boundary faces of all devices)
                res = adapter_e2e(args, world, viscous, args.steps, 3, modes=("device_bcs", "host_bcs") if (world == 1 and not args.no_aux_lines) else ("device_bcs",))
                e2e = res["device_bcs"]
                if "host_bcs" in res:
                    aux["e2e_adapter_host_bcs"] = {k: res["host_bcs"].get(k) for k in keep}
            except Exception as ex:
                e2e = {"value": None, "unit": "DOF-stage/s", "h2d_bytes_per_step": None, "d2h_bytes_per_step": None, "mode": "adapter run failed: %r" % (ex,)}
            if world == 1 and not args.no_aux_lines and not viscous:
                try:  # what the adapter's zero-Solver-change default costs (every call moves everything over PCIe), once, on a small mesh
                    sc = adapter_e2e(args, 1, False, 2, 1, modes=("sync_every_call",), n_override=24)["sync_every_call"]
                    aux["e2e_adapter_sync_every_call"] = {k: sc.get(k) for k in keep}
                except Exception as ex:
                    aux["e2e_adapter_sync_every_call"] = {"value": None, "mode": "failed: %r" % (ex,)}
        if dist is not None:
            dist.barrier(group=host_group)

    # ---- the other workloads of the BASELINE config list on the same box, as sub-lines of the default run (own processes, 1 M elements each) ----
    if rank == 0 and world == 1 and not args.no_aux_lines and args.n == 100 and nd == 3 and not viscous and deformed:
        sub = {}
        for name, extra in (("3d_cartesian_euler", ["--mesh", "cartesian", "--no-e2e"]), ("3d_deformed_navier_stokes", ["--pde", "navier_stokes"]),
                            ("3d_cartesian_navier_stokes", ["--pde", "navier_stokes", "--mesh", "cartesian", "--no-e2e"])):
            try:  # (the deformed Navier-Stokes sub-line also carries its end-to-end number through the adapter: viscous stage with the flux_bc callback)
                r = subprocess.run([sys.executable, os.path.abspath(__file__), "--steps", str(args.steps), "--warmup", str(max(args.warmup, 3)),
                                    "--no-cpu-baseline", "--no-aux-lines"] + extra, capture_output=True, text=True, timeout=900)
                line = json.loads([x for x in r.stdout.splitlines() if x.startswith("{")][-1])
                sub[name] = {"value": line["value"], "unit": line["unit"], "ms_per_step": line["ms_per_step"], "workload": line["config"]["workload"],
                             "e2e": ({k: line["e2e"].get(k) for k in ("value", "unit", "ms_per_step", "h2d_bytes_per_step", "d2h_bytes_per_step", "elements", "box")}
                                     if line.get("e2e") else None),
                             "timed_path": line["config"].get("timed_path"), "gpu_launches": line.get("gpu_launches"), "clocks": line.get("clocks"),
                             "roofline": {k: line["roofline"].get(k) for k in ("kernel", "achieved", "peak", "frac", "frac_traffic", "avg_launch_ms", "whole_stage",
                                                                                  "kernel_seconds_per_step", "accounting")}}
            except Exception as ex:
                sub[name] = {"value": None, "error": repr(ex)}
        aux["sub_lines"] = sub

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            rate, cs, cores, kind, sample = cpu_reference_rate(args, args.cpu_steps, 3)
            cpu = {"value": rate, "unit": "DOF-stage/s", "cores": cores, "kind": kind, "sample": sample, "host_cpu": host_info()["cpu_model"],
                   "both_cpu_implementations": getattr(cpu_reference_rate, "both", None)}
        except Exception as ex:  # the baseline is informative; never let it take the GPU number down
            cpu = {"value": None, "unit": "DOF-stage/s", "cores": None, "kind": "port", "sample": "failed: %r" % (ex,)}

    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": "DOF-stage/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": sec/args.steps*1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args, viscous),
                       "timed_path": ("hexed_b200_update_%s: the flow loop of Solver::update with dt kept on the device, device boundary conditions, one CUDA graph per "
                                      "step, no host synchronisation in the timed region (call by call through compute_euler/max_dt_euler with a dt read-back per step: "
                                      "%.2f ms/step)" % ("navier_stokes" if viscous else "euler", sec_call_by_call/args.steps*1e3)) if device_loop else
                                     "call by call: max_dt (NCCL allreduce + read-back) + 2 x (device boundary conditions + split stage around the halo exchange)",
                       "elements_per_gpu": ne, "dof_per_element": nv*nq, "stages_per_step": 2,
                       "l2": "working set %.1f GB per GPU >> 126 MB L2 (no flush needed)" % (ne*(50e3 if nd == 3 else 6.2e3)/1e9),
                       "parallelism": "1 GPU" if world == 1 else
                       "%s blocks of one global box (Z-order split), cut faces exchanged by NCCL send/recv (%d B per rank and stage) overlapped with interior flux work, dt by NCCL allreduce(min)"
                       % ("x".join(map(str, blocks)), halo.bytes_per_exchange)},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": local_name,
                         "achieved": local_gbs, "peak": peak, "unit": "GB/s", "frac": local_gbs/peak,
                         "traffic": traffic["bytes"] if traffic else None, "traffic_source": traffic["source"] if traffic else None,
                         "frac_traffic": (traffic["bytes"]/local_sec/1e9/peak) if traffic else None,
                         "algorithmic_bytes_note": "frac uses the fixed SURVEY 8(d) denominator, which charges the Euler Local kernel bytes that are never moved: 648 "
                                                   "doubles/element of face normals (neither the reference's Euler Local nor ours reads them) and the time-step scale "
                                                   "(216 doubles/element; not read while it is known to hold 1 after a global-time-step max_dt). So frac can exceed "
                                                   "1; frac_traffic (DRAM bytes measured by ncu per launch / the same time) is the fraction of the measured HBM "
                                                   "bandwidth the kernel really sustains",
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": ne*alg["local"]*8, "avg_launch_ms": local_sec*1e3,
                         "whole_stage": {"achieved": stage_gbs, "frac": stage_gbs/peak, "bytes_per_dof_stage": alg["stage"]*8/float(nv*nq),
                                         "frac_moved": (stage_gbs/peak*(1. - (2*deformed*nd*nd*nfq_ + 1.5*nq)/float(alg["stage"]))) if not viscous else None,
                                         "frac_moved_note": "the same with the bytes that are not moved taken out of the denominator: face normals in Local, "
                                                            "the time-step scale read in Local and its write in max_dt"},
                         "kernel_seconds_per_step": shares,
                         "accounting": {"sum_kernels_ms": sum(shares.values())*1e3, "ms_per_step": sec/args.steps*1e3,
                                        "unaccounted_frac": 1. - sum(shares.values())/(sec/args.steps),
                                        "note": "kernel times from a separate pass with per-launch CUDA events; the step itself is timed without them"}},
            "e2e": e2e, "cpu_baseline": cpu, "aux": aux,
        }
        print(json.dumps(out))
    dev.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
