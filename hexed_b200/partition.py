"""Domain decomposition of a flattened kernel mesh across the GPUs of one box.

The reference is single-process; this module adds the one thing a multi-GPU run needs on the host side:
split the element set by a space-filling curve, give every rank a self-contained `FlatMesh` whose cut faces are backed by
HALO face slots, and list which face slots have to be sent to / received from which peer before the `Neighbor` kernels run.

Rules (they keep every kernel unchanged and the result bit-identical to the undivided run):
  * an element belongs to exactly one rank; its 2*n_dim faces live on that rank;
  * a connection whose two faces belong to different ranks is REPLICATED on both ranks with the remote face replaced by a halo
    slot. Both ranks compute the same flux from the same two inputs (the LLF flux has no rank-dependent term), so no flux has
    to be sent back. Cut connections are placed at the end of `car_con` / `def_con` (`n_cut_car`, `n_cut_def`) so the interior
    ones can run while the exchange is in flight;
  * a hanging-node face (`Refined_face`: coarse face + mortar faces + fine connections, reference include/Refined_face.hpp:9-15,
    include/connection.hpp:218-269) is replicated on every rank that owns one of its participants. Where the coarse element is
    remote its face state arrives through a halo slot and is prolonged locally before the fine connections are evaluated
    (`pre_prolong`), where a fine element is remote its face arrives through a halo slot; `Restrict_refined` runs everywhere and
    its result is only consumed on the rank that owns the coarse element;
  * connection-owned storage (boundary ghosts, mortar faces, connection normals) is copied to every rank that uses it.
"""
import numpy as np

from .mesh import FlatMesh


def morton_keys(index, bits=21):
    """Z-order key of integer coordinates (n, n_dim): interleaved bits, last dimension least significant"""
    index = np.asarray(index, dtype=np.uint64)
    n, nd = index.shape
    key = np.zeros(n, dtype=np.uint64)
    for b in range(bits):
        for d in range(nd):
            key |= ((index[:, d] >> np.uint64(b)) & np.uint64(1)) << np.uint64(b*nd + (nd - 1 - d))
    return key


def split_by_curve(keys, n_parts, weights=None):
    """contiguous equal-weight ranges along the curve: returns part id per element"""
    order = np.argsort(keys, kind="stable")
    w = np.ones(len(keys)) if weights is None else np.asarray(weights, dtype=float)
    cum = np.cumsum(w[order])
    part_sorted = np.minimum((cum - 0.5*w[order])*n_parts/cum[-1], n_parts - 1).astype(np.int64)
    part = np.empty(len(keys), dtype=np.int64)
    part[order] = part_sorted
    return part


class Halo:
    """what one rank exchanges before `Neighbor`: per peer, the local face slots to send and the halo slots to fill"""

    def __init__(self):
        self.send = {}  # peer -> int32 array of local face slots (ordered by global slot id)
        self.recv = {}  # peer -> int32 array of local halo slots (same order as the peer's send list)

    def peers(self):
        return sorted(set(self.send) | set(self.recv))


def partition_mesh(mesh, part, n_parts):
    """split `mesh` (FlatMesh with numpy arrays) into `n_parts` self-contained meshes; `part[e]` = owner of element e.
    Returns a list of FlatMesh with the extra attributes `global_elem`, `halo`, `n_cut_car`, `n_cut_def`, `pre_prolong`."""
    nd, rs = mesh.n_dim, mesh.row_size
    nf = 2*nd
    ne_g = mesh.n_elem
    part = np.asarray(part)
    n_elem_slots = nf*ne_g
    owner_of_slot = lambda g: int(part[g//nf]) if g < n_elem_slots else -1  # noqa: E731
    # which refined face a mortar slot belongs to, and the participants of every refined face
    mortar_ref = {}
    for r, row in enumerate(mesh.ref_face):
        for g in row[1:5]:
            if g >= 0:
                mortar_ref[int(g)] = r
    ref_ranks = [set() for _ in range(len(mesh.ref_face))]
    for r, row in enumerate(mesh.ref_face):
        ref_ranks[r].add(owner_of_slot(int(row[0])))
    for table in (mesh.car_con, mesh.def_con):  # Cartesian hanging faces keep their fine connections in car_con
        for con in table:
            for side in (0, 1):
                if int(con[side]) in mortar_ref:
                    ref_ranks[mortar_ref[int(con[side])]].add(owner_of_slot(int(con[1 - side])))
    bc_of_con = {}
    for ib, bc in enumerate(mesh.bcs):
        for k, ci in enumerate(bc["con_index"]):
            bc_of_con[int(ci)] = (ib, k)

    parts = []
    for p in range(n_parts):
        mine = np.nonzero(part == p)[0]
        car = mine[mine < mesh.n_car]
        dfm = mine[mine >= mesh.n_car]
        glob = np.concatenate([car, dfm])
        local_of = {int(g): i for i, g in enumerate(glob)}
        n_car, n_def = len(car), len(dfm)
        extra_slots = {}   # global face slot -> local extra slot (halo / ghost / mortar)
        halo_recv = {}     # peer -> list of global slots received from it
        extra_normals = {}  # global normal slot -> local extra normal slot

        def lslot(g):
            g = int(g)
            if g < n_elem_slots and int(part[g//nf]) == p:
                return local_of[g//nf]*nf + g % nf
            if g not in extra_slots:
                extra_slots[g] = nf*len(glob) + len(extra_slots)
                if g < n_elem_slots:
                    halo_recv.setdefault(int(part[g//nf]), []).append(g)
            return extra_slots[g]

        def lnormal(g):
            g = int(g)
            if g < nf*mesh.n_def:
                e = mesh.n_car + g//nf
                if int(part[e]) == p:
                    return (local_of[e] - n_car)*nf + g % nf
            if g not in extra_normals:
                extra_normals[g] = nf*n_def + len(extra_normals)
            return extra_normals[g]

        car_int, car_cut, def_int, def_cut = [], [], [], []
        bc_local = [dict(kind=bc["kind"], params=bc.get("params"), inside_slot=[], ghost_slot=[], normal_slot=[], con_index=[]) for bc in mesh.bcs]
        def classify(s0, s1):
            """(kept on rank p, must wait for the exchange) for a connection between face slots s0 and s1"""
            r = mortar_ref.get(s0, mortar_ref.get(s1))
            if r is not None:
                # fine connection of a hanging-node face: kept where the fine element is local, or where the coarse element is local;
                # on the fine element's rank the mortar face is only valid after the coarse face arrived and was prolonged locally
                fine_owner = owner_of_slot(s1 if s0 in mortar_ref else s0)
                coarse_owner = owner_of_slot(int(mesh.ref_face[r][0]))
                return p in (fine_owner, coarse_owner), (fine_owner != p or coarse_owner != p)
            o0, o1 = owner_of_slot(s0), owner_of_slot(s1)
            if o0 < 0 or o1 < 0:
                return (o1 if o0 < 0 else o0) == p, False  # boundary connection: belongs to its element's rank
            return p in (o0, o1), o0 != o1

        for con in mesh.car_con:
            keep, cut = classify(int(con[0]), int(con[1]))
            if not keep:
                continue
            row = [lslot(con[0]), lslot(con[1]), int(con[2])]
            (car_cut if cut else car_int).append(row)
        def_rows = []  # (is_cut, row, bc tag)
        for ci, con in enumerate(mesh.def_con):
            s0, s1 = int(con[0]), int(con[1])
            keep, cut = classify(s0, s1)
            if not keep:
                continue
            row = [lslot(s0), lslot(s1), int(con[2]), int(con[3]), int(con[4]), int(con[5]), lnormal(con[6])]
            def_rows.append((cut, row, bc_of_con.get(ci)))
        def_rows.sort(key=lambda x: x[0])  # stable: interior first, cut last
        for i, (cut, row, tag) in enumerate(def_rows):
            (def_cut if cut else def_int).append(row)
            if tag is not None:
                ib, k = tag
                b = bc_local[ib]
                src = mesh.bcs[ib]
                b["inside_slot"].append(lslot(src["inside_slot"][k])); b["ghost_slot"].append(lslot(src["ghost_slot"][k]))
                b["normal_slot"].append(lnormal(src["normal_slot"][k])); b["con_index"].append(i)
        refs, pre_prolong = [], []
        for r, row in enumerate(mesh.ref_face):
            if p not in ref_ranks[r]:
                continue
            if owner_of_slot(int(row[0])) != p:
                pre_prolong.append(len(refs))
            refs.append([lslot(row[0])] + [lslot(g) if g >= 0 else -1 for g in row[1:5]] + [int(row[5]), int(row[6])])

        m = FlatMesh(nd, rs, n_car, n_def, n_ghost=len(extra_slots), n_extra_normal=len(extra_normals),
                     with_ldg=mesh.face_ldg is not None, with_wide=mesh.face_wide is not None)
        m.global_elem = glob
        m.elem_data[:] = mesh.elem_data[glob]
        m.nom_size[:] = mesh.nom_size[glob]
        m.vertex_tss[:] = mesh.vertex_tss[glob]
        m.uncert[:] = mesh.uncert[glob]
        if n_def:
            m.ref_normals[:] = mesh.ref_normals[dfm - mesh.n_car]
            m.det[:] = mesh.det[dfm - mesh.n_car]
            m.normals[:nf*n_def] = mesh.normals[:nf*mesh.n_def].reshape(mesh.n_def, nf, nd, mesh.nfq)[dfm - mesh.n_car].reshape(nf*n_def, nd, mesh.nfq)
        for g, l in extra_normals.items():
            m.normals[l] = mesh.normals[g]
        if hasattr(mesh, "qpoint_pos") and mesh.qpoint_pos is not None:
            m.qpoint_pos = np.asarray(mesh.qpoint_pos)[glob]
        # face data: local element faces + copies of everything held in extra slots (so a freshly partitioned mesh is consistent)
        src_elem = (glob[:, None]*nf + np.arange(nf)[None, :]).reshape(-1)
        for name in ("face_state", "face_ldg", "face_wide"):
            a = getattr(mesh, name)
            if a is None:
                continue
            b = getattr(m, name)
            b[:nf*len(glob)] = a[src_elem]
            for g, l in extra_slots.items():
                b[l] = a[g]
        m.car_con = np.array(car_int + car_cut, np.int32).reshape(-1, 3)
        m.def_con = np.array(def_int + def_cut, np.int32).reshape(-1, 7)
        m.n_cut_car, m.n_cut_def = len(car_cut), len(def_cut)
        m.ref_face = np.array(refs, np.int32).reshape(-1, 7)
        m.pre_prolong = np.array(pre_prolong, np.int32)
        m.bcs = [dict(kind=b["kind"], params=b["params"], inside_slot=np.array(b["inside_slot"], np.int32),
                      ghost_slot=np.array(b["ghost_slot"], np.int32), normal_slot=np.array(b["normal_slot"], np.int32),
                      con_index=np.array(b["con_index"], np.int32)) for b in bc_local]
        m._extra_slots = extra_slots
        m._halo_recv_global = {peer: sorted(v) for peer, v in halo_recv.items()}
        m._local_of = local_of
        parts.append(m)

    # send lists: what peer q receives from p, in the same (global slot) order
    for p, m in enumerate(parts):
        m.halo = Halo()
        for peer, globs in m._halo_recv_global.items():
            m.halo.recv[peer] = np.array([m._extra_slots[g] for g in globs], np.int32)
    for p, m in enumerate(parts):
        for q, other in enumerate(parts):
            globs = other._halo_recv_global.get(p)
            if globs:
                m.halo.send[q] = np.array([m._local_of[g//nf]*nf + g % nf for g in globs], np.int32)
    return parts


def gather_elements(parts, global_mesh):
    """write the element data of the parts back into the global mesh (test helper)"""
    for m in parts:
        global_mesh.elem_data[m.global_elem] = m.elem_data
        nf = 2*m.n_dim
        dst = (m.global_elem[:, None]*nf + np.arange(nf)[None, :]).reshape(-1)
        global_mesh.face_state[dst] = m.face_state[:nf*m.n_elem]
    return global_mesh
