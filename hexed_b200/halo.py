"""Halo exchange of cut-face states between the ranks of a partitioned mesh (see partition.py).

`torch.distributed` is the transport (NCCL send/recv over NVLink on the GPUs, gloo in the CPU tests); packing and unpacking of
the face slots is done by the library's gather / scatter kernels straight into / out of the communication buffers, on the
context's own stream, so that `compute_euler_begin` (interior connections) overlaps the transfer:

    halo.start()                      # gather cut faces -> send buffers, post isend / irecv
    dev.compute_euler_begin()         # Neighbor on connections that touch no halo face
    halo.finish()                     # wait, scatter received faces into the halo slots
    dev.compute_euler_finish(...)     # pre-prolong, Neighbor on cut connections, Restrict, Local, Prolong
"""
import numpy as np
import torch


def _dist():
    import torch.distributed as dist
    return dist


class DeviceHalo:
    """exchange for a `Device` whose mesh came from `partition_mesh` / `box_partition`; one instance per face kind"""

    def __init__(self, dev, mesh, kind=0, group=None):
        self.dev, self.kind, self.group = dev, kind, group
        width = {0: mesh.nv*mesh.nfq, 1: mesh.nv*mesh.nfq, 2: (mesh.n_dim + mesh.row_size)*mesh.nfq}[kind]
        cuda = torch.device("cuda", torch.cuda.current_device())
        self.stream = torch.cuda.ExternalStream(dev.cuda_stream(), device=cuda)
        self.send_buf = {p: torch.empty((len(s), width), dtype=torch.float64, device=cuda) for p, s in mesh.halo.send.items()}
        self.recv_buf = {p: torch.empty((len(s), width), dtype=torch.float64, device=cuda) for p, s in mesh.halo.recv.items()}
        self.bytes_per_exchange = sum(b.numel() for b in self.send_buf.values())*8
        self.reqs = []

    def start(self):
        dist = _dist()
        for p, buf in self.send_buf.items():
            self.dev.face_list_gather(self.dev.send_lists[p], buf, self.kind)
        ops = []
        for p in sorted(set(self.send_buf) | set(self.recv_buf)):
            if p in self.recv_buf:
                ops.append(dist.P2POp(dist.irecv, self.recv_buf[p], p, group=self.group))
            if p in self.send_buf:
                ops.append(dist.P2POp(dist.isend, self.send_buf[p], p, group=self.group))
        if ops:
            with torch.cuda.stream(self.stream):  # NCCL orders itself after the gather kernels on the context's stream
                self.reqs = dist.batch_isend_irecv(ops)

    def finish(self):
        with torch.cuda.stream(self.stream):
            for r in self.reqs:
                r.wait()  # stream-ordered: the context's stream waits for the transfer, the host does not
        self.reqs = []
        for p, buf in self.recv_buf.items():
            self.dev.face_list_scatter(self.dev.recv_lists[p], buf, self.kind)


class MeshHalo:
    """the same exchange for a numpy `FlatMesh` (CPU oracle runs in the gloo tests)"""

    def __init__(self, mesh, kind=0, group=None):
        self.mesh, self.kind, self.group = mesh, kind, group

    def _faces(self):
        return {0: self.mesh.face_state, 1: self.mesh.face_ldg, 2: self.mesh.face_wide}[self.kind]

    def exchange(self):
        dist = _dist()
        faces = self._faces()
        m = self.mesh
        send = {p: torch.from_numpy(np.ascontiguousarray(faces[s])) for p, s in m.halo.send.items()}
        recv = {p: torch.empty((len(s), faces.shape[1]), dtype=torch.float64) for p, s in m.halo.recv.items()}
        reqs = []
        for p in sorted(set(send) | set(recv)):
            if p in recv:
                reqs.append(dist.irecv(recv[p], p, group=self.group))
            if p in send:
                reqs.append(dist.isend(send[p], p, group=self.group))
        for r in reqs:
            r.wait()
        for p, s in m.halo.recv.items():
            faces[s] = recv[p].numpy()


def exchange_in_process(parts, get, put):
    """all parts live in this process (single-GPU tests of the partition logic): `get(p, slots)` returns the packed faces of part p,
    `put(q, slots, data)` stores them into part q"""
    staged = []
    for p, m in enumerate(parts):
        for q, slots in m.halo.send.items():
            staged.append((q, parts[q].halo.recv[p], get(p, slots)))
    for q, slots, data in staged:
        put(q, slots, data)


def allreduce_min(value, device=None, group=None):
    dist = _dist()
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
    return float(t.item())


def allreduce_and(flag, device=None, group=None):
    """`Solver::is_admissible` of a partitioned mesh: every rank's `Device.is_admissible()` must hold"""
    dist = _dist()
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return bool(flag)
    t = torch.tensor([1 if flag else 0], dtype=torch.int32, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
    return bool(int(t.item()))
