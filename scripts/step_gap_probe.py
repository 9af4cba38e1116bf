"""Where do the ~1.5 ms per step between the sum of the kernel times and the step time of the headline bench go?
Times, on one box: the bench's host loop (max_dt with its read-back + 2 x (ghost fill + compute_euler)), the same with the time
step left on the device (update_euler, eager) and as one CUDA graph per step. Prints one JSON line."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hexed_b200 as hb  # noqa: E402
from hexed_b200 import mesh as M  # noqa: E402
from hexed_b200.kernels import Device  # noqa: E402
from hexed_b200.cases import freestream_state  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
steps = 10
nd, rs = 3, 6
cuda = torch.device("cuda", 0)
basis = hb.gauss_legendre(rs)
fs = freestream_state(nd)
m = M.box_mesh(nd, rs, n, basis, deformed=True, bc_kind=M.BC_FREESTREAM, bc_params=fs, device=cuda, geometry_chunk=32768,
               keep_geometry_on_device=True, lean=True)
ne, nq, nv = m.n_elem, m.nq, m.nv
dev = Device(nd, rs, basis, device=0).load_mesh(m, upload_elem_data=False)
m.ref_normals = None; m.det = None; m.normals = None
pos = m.qpoint_pos if torch.is_tensor(m.qpoint_pos) else torch.as_tensor(m.qpoint_pos, device=cuda)
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
m.qpoint_pos = None
torch.cuda.empty_cache()
dev.compute_write_face()
stream = torch.cuda.ExternalStream(dev.cuda_stream(), device=cuda)


def host_loop(k):
    for _ in range(k):
        dt = dev.max_dt_euler(0.7, 0.7, False)
        for stage in (0, 1):
            dev.apply_state_bcs()
            dev.compute_euler(dt=dt, i_stage=stage)


def timed(fn):
    dev.synchronize(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    fn()
    e1.record(stream)
    dev.synchronize(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/steps


out = {"n_elem": ne, "steps": steps}
host_loop(3)
for rep in range(2):
    out["host_loop_ms_%d" % rep] = timed(lambda: host_loop(steps))
    out["update_eager_ms_%d" % rep] = timed(lambda: dev.update_euler(0.7, steps, use_graph=False))
    out["update_graph_ms_%d" % rep] = timed(lambda: dev.update_euler(0.7, steps + 1, use_graph=True))*steps/(steps + 1)
print(json.dumps(out))
