"""Flattened kernel mesh: the plain-array form of the reference's `Kernel_mesh`.

The reference hands its kernels six `Sequence<>` views over heap objects that alias each other through
raw pointers (include/Kernel_mesh.hpp:14-25, include/connection.hpp:111-123). The device cannot chase that
pointer graph, so the host side flattens it ONCE per mesh epoch into slot-indexed arrays plus small integer
tables; `FlatMesh` is that flattened form. Layout conventions are documented in include/hexed_b200.h
(and mirrored in oracle/flat_mesh.h for the CPU oracle).

This module also synthesises the benchmark / test meshes: structured boxes of
Cartesian or smoothly deformed elements whose metric terms follow the reference's
`Deformed_element::set_jacobian` (src/Deformed_element.cpp:60-136) and `Solver::calc_jacobian`
(src/Solver.cpp:273-381). Geometry is computed with torch so the 1M-element benchmark mesh can be built on
the GPU; everything else is numpy.
"""
import numpy as np
import torch

from .tables import Connection_direction

N_FORCING = 4  # reference include/Storage_params.hpp:21

BC_FREESTREAM, BC_COPY, BC_NONPENETRATION, BC_OUTFLOW, BC_PRESSURE_OUTFLOW, BC_NO_SLIP, BC_RIEMANN_INVARIANTS = 0, 1, 2, 3, 4, 5, 6  # include/hexed_b200.h
BC_HOST = 99  # applied by the host (anything the device does not implement)
THERMAL_HEAT_FLUX, THERMAL_ENERGY, THERMAL_EQUILIBRIUM = 0, 1, 2  # Prescribed_heat_flux / Prescribed_energy / Thermal_equilibrium


def _mpow(x, n):
    """math::pow (reference include/math.hpp:29-36): repeated multiplication starting from 1"""
    r = 1.
    for _ in range(n):
        r = r*x
    return r


# reference include/constants.hpp:52 with the same operation order
STEFAN_BOLTZMANN = 2*_mpow(np.pi, 5)*_mpow(1.380649e-23, 4)/(15*_mpow(299792458., 2)*_mpow(6.62607015e-34, 3))


def no_slip_params(thermal_kind=THERMAL_HEAT_FLUX, a=0., b=0., c=0., coercion=2.):
    """parameter block of the device No_slip boundary condition: [thermal kind, a, b, c, coercion, stefan_boltzmann] with
    a = heat flux | energy per mass | emissivity, b = heat transfer coefficient, c = temperature (reference include/Boundary_condition.hpp:140-200)"""
    return np.array([thermal_kind, a, b, c, coercion, STEFAN_BOLTZMANN], dtype=np.float64)


def n_slot(n_dim, row_size):
    """slots of `nq` doubles per element (reference src/Storage_params.cpp:32-35 with n_stage = 2)"""
    nv = n_dim + 2
    return nv + 3 + N_FORCING + row_size + max(nv, row_size)


def cache_slot(n_dim, row_size):
    """first slot of the residual cache (reference src/Element.cpp:188)"""
    return n_dim + 2 + 3 + N_FORCING + row_size


TSS_SLOT = lambda n_dim: n_dim + 2          # noqa: E731  (reference include/pde.hpp:17)
BULK_AV_SLOT = lambda n_dim: n_dim + 3      # noqa: E731
LAPLACIAN_AV_SLOT = lambda n_dim: n_dim + 4  # noqa: E731
FORCING_SLOT = lambda n_dim: n_dim + 5      # noqa: E731
ADVECTION_SLOT = lambda n_dim: n_dim + 9    # noqa: E731


class FlatMesh:
    """slot-indexed arrays + integer tables; see module docstring"""

    def __init__(self, n_dim, row_size, n_car, n_def, n_ghost=0, n_extra_normal=0, alloc_elem_data=True,
                 with_ldg=False, with_wide=False, alloc_faces=True):
        self.n_dim, self.row_size = n_dim, row_size
        self.n_car, self.n_def = n_car, n_def
        self.nq = row_size**n_dim
        self.nfq = row_size**(n_dim - 1)
        self.nv = n_dim + 2
        self.n_slot = n_slot(n_dim, row_size)
        ne = n_car + n_def
        self.n_face_slot = 2*n_dim*ne + n_ghost
        self.n_normal_slot = 2*n_dim*n_def + n_extra_normal
        self.elem_data = np.zeros((ne, self.n_slot, self.nq)) if alloc_elem_data else None
        if alloc_elem_data:
            self.elem_data[:, TSS_SLOT(n_dim)] = 1.  # reference src/Element.cpp:24
        self.nom_size = np.ones(ne)
        self.vertex_tss = np.ones((ne, 2**n_dim))
        self.uncert = np.zeros(ne)
        self.ref_normals = np.zeros((n_def, n_dim*n_dim, self.nq))
        self.det = np.ones((n_def, self.nq))
        self.face_state = np.zeros((self.n_face_slot, self.nv*self.nfq)) if alloc_faces else None
        self.face_ldg = np.zeros((self.n_face_slot, self.nv*self.nfq)) if with_ldg else None
        self.face_wide = np.zeros((self.n_face_slot, (n_dim + row_size)*self.nfq)) if with_wide else None
        self.normals = np.zeros((self.n_normal_slot, n_dim, self.nfq))
        # unit-normal fallback for deformed faces without a deformed connection (reference include/Spatial.hpp:331-339,366)
        if n_def:
            n = self.normals[:2*n_dim*n_def].reshape(n_def, n_dim, 2, n_dim, self.nfq)
            for d in range(n_dim):
                n[:, d, :, d, :] = 1.
        self.car_con = np.zeros((0, 3), np.int32)
        self.def_con = np.zeros((0, 7), np.int32)
        self.ref_face = np.zeros((0, 7), np.int32)
        self.bcs = []  # list of dicts: kind, inside_slot, ghost_slot, normal_slot, con_index, params

    # ---- slot arithmetic ----
    @property
    def n_elem(self):
        return self.n_car + self.n_def

    def face_slot(self, elem, i_face):
        return elem*2*self.n_dim + i_face

    def elem_normal_slot(self, elem, i_face):
        return (elem - self.n_car)*2*self.n_dim + i_face

    # ---- views into elem_data ----
    def state(self):
        return self.elem_data[:, :self.nv]

    def tss(self):
        return self.elem_data[:, TSS_SLOT(self.n_dim)]

    def cache(self):
        c = cache_slot(self.n_dim, self.row_size)
        return self.elem_data[:, c:c + self.nv]

    def copy(self):
        import copy
        return copy.deepcopy(self)


# --------------------------------------------------------------------------------------
# metric terms (torch; batch over elements)
# --------------------------------------------------------------------------------------

def _apply_along(mat, arr, axis):
    """arr: (..., rs, rs, rs)-like tensor; contracts `mat[i][j]` with index j on `axis` (axis counted among the trailing n_dim axes)"""
    a = torch.movedim(arr, axis, -1)
    r = torch.matmul(a, mat.T)
    return torch.movedim(r, -1, axis)


def _cofactor_rows(cols, n_dim):
    """cols[k]: (..., n_dim) = d pos / d ref_k. Returns rows[i]: (..., n_dim) with rows[i][j] = det(J with column i := e_j)"""
    if n_dim == 1:
        return [torch.ones_like(cols[0])]
    if n_dim == 2:
        c0, c1 = cols
        return [torch.stack([c1[..., 1], -c1[..., 0]], -1), torch.stack([-c0[..., 1], c0[..., 0]], -1)]
    c0, c1, c2 = cols
    return [torch.linalg.cross(c1, c2), torch.linalg.cross(c2, c0), torch.linalg.cross(c0, c1)]


def element_metrics(vert_pos, nom_size, basis, device=None):
    """Metric terms of trilinear elements, following `Deformed_element::set_jacobian`.

    vert_pos: (E, 2^nd, nd) vertex positions (vertex index row-major, last dimension fastest)
    nom_size: (E,)
    returns dict of torch tensors: ref_normals (E, nd*nd, nq) [i_dim][j_dim], det (E, nq),
    face_normals (E, 2*nd, nd, nfq), vertex_tss (E, 2^nd)
    """
    vert_pos = torch.as_tensor(vert_pos, dtype=torch.float64, device=device)
    nom = torch.as_tensor(nom_size, dtype=torch.float64, device=vert_pos.device)
    E, n_vert, nd = vert_pos.shape
    rs = basis.row_size
    node = torch.as_tensor(basis.node, device=vert_pos.device)
    diff = torch.as_tensor(basis.diff_mat, device=vert_pos.device)
    bnd = torch.as_tensor(basis.boundary, device=vert_pos.device)
    interp = torch.stack([1. - node, node], -1)  # (rs, 2)
    # position at quadrature points: multilinear interpolation of the vertices (src/Deformed_element.cpp:15-58 with zero node adjustments)
    pos = vert_pos.movedim(-1, 1).reshape((E, nd) + (2,)*nd)  # (E, nd, 2, 2, 2)
    for d in range(nd):
        pos = _apply_along(interp, pos, 2 + d)
    # jacobian entries jac[i][j] = d pos_i / d ref_j / nom_size
    scale = nom.reshape((E,) + (1,)*(nd + 1))
    jac = [_apply_along(diff, pos, 2 + j)/scale for j in range(nd)]  # jac[j]: (E, nd(i), rs, rs, rs)
    cols = [jac[j].movedim(1, -1) for j in range(nd)]                # cols[j][..., i]
    rows = _cofactor_rows(cols, nd)                                   # rows[i][..., j]
    ref_normals = torch.stack([rows[i][..., j] for i in range(nd) for j in range(nd)], 1).reshape(E, nd*nd, rs**nd)
    if nd == 1:
        det = cols[0][..., 0]
    elif nd == 2:
        det = cols[0][..., 0]*cols[1][..., 1] - cols[0][..., 1]*cols[1][..., 0]
    else:
        det = (cols[0]*torch.linalg.cross(cols[1], cols[2])).sum(-1)
    det = det.reshape(E, rs**nd)
    # face normals: extrapolate the jacobian to the face, then the same replaced-column determinants (src/Deformed_element.cpp:91-113)
    face_normals = torch.empty((E, 2*nd, nd, rs**(nd - 1)), dtype=torch.float64, device=vert_pos.device)
    for d in range(nd):
        for sign in range(2):
            fcols = []
            for j in range(nd):
                fj = _apply_along(bnd[sign:sign + 1], jac[j], 2 + d)  # (E, nd, ..1..)
                fcols.append(fj.movedim(1, -1))
            frows = _cofactor_rows(fcols, nd)
            face_normals[:, 2*d + sign] = frows[d].movedim(-1, 1).reshape(E, nd, -1)
    # vertex time step scale (src/Deformed_element.cpp:114-135)
    vn = ref_normals.reshape((E, nd*nd) + (rs,)*nd)
    vd = det.reshape((E, 1) + (rs,)*nd)
    for d in range(nd):
        vn = _apply_along(bnd, vn, 2 + d)
        vd = _apply_along(bnd, vd, 2 + d)
    vn = vn.reshape(E, nd, nd, n_vert)
    vd = vd.reshape(E, n_vert)
    norm_sum = torch.sqrt((vn*vn).sum(2)).sum(1)
    vertex_tss = nom[:, None]*vd/norm_sum
    return dict(ref_normals=ref_normals, det=det, face_normals=face_normals, vertex_tss=vertex_tss, pos=pos.reshape(E, nd, rs**nd))


def qpoint_positions_cartesian(origin_index, nom_size, basis, n_dim):
    """positions of quadrature points of Cartesian elements (reference src/Element.cpp:62-70); returns (E, nd, nq)"""
    rs = basis.row_size
    E = origin_index.shape[0]
    out = np.empty((E, n_dim, rs**n_dim))
    q = np.arange(rs**n_dim)
    for d in range(n_dim):
        stride = rs**(n_dim - d - 1)
        out[:, d, :] = (basis.node[(q//stride) % rs][None, :] + origin_index[:, d:d + 1])*np.asarray(nom_size)[:, None]
    return out


# --------------------------------------------------------------------------------------
# structured box
# --------------------------------------------------------------------------------------

def default_warp(x, amplitude):
    """smooth displacement used for the synthetic deformed box: x += a*h*prod(sin(2 pi x_i))"""
    s = torch.ones_like(x[..., 0])
    for d in range(x.shape[-1]):
        s = s*torch.sin(2*np.pi*x[..., d])
    return x + amplitude*s[..., None]


def proc_grid(world, n_dim=3):
    """blocks per dimension for `world` ranks, filled in Z-order (x halves first): 1 -> (1,1,1), 2 -> (2,1,1), 4 -> (2,2,1), 8 -> (2,2,2)"""
    grid = [1]*n_dim
    d = 0
    w = world
    while w > 1:
        if w % 2:
            raise ValueError("world size must be a power of two")
        grid[d % n_dim] *= 2
        w //= 2
        d += 1
    return tuple(grid)


def block_rank(coords, grid):
    """rank of the block at integer block coordinates: Z-order (Morton) over the block grid, consistent with partition.morton_keys"""
    nd = len(grid)
    bits = max(int(g - 1).bit_length() for g in grid)
    key = 0
    for b in range(bits):
        for d in range(nd):
            key |= ((coords[d] >> b) & 1) << (b*nd + (nd - 1 - d))
    # compact the key to 0..world-1 (grid sides are powers of two, possibly unequal)
    order = sorted(_all_block_keys(grid))
    return order.index(key)


def _all_block_keys(grid):
    nd = len(grid)
    bits = max(int(g - 1).bit_length() for g in grid)
    keys = []
    for c in np.ndindex(*grid):
        key = 0
        for b in range(bits):
            for d in range(nd):
                key |= ((c[d] >> b) & 1) << (b*nd + (nd - 1 - d))
        keys.append(key)
    return keys


def block_coords(rank, grid):
    for c in np.ndindex(*grid):
        if block_rank(c, grid) == rank:
            return tuple(int(x) for x in c)
    raise ValueError("rank outside the block grid")


def box_mesh(n_dim, row_size, n, basis, deformed=False, warp_amplitude=0.1, bc_kind=BC_FREESTREAM, bc_params=None,
             device=None, geometry_chunk=32768, with_ldg=False, keep_geometry_on_device=False, lean=False,
             blocks=None, block=None):
    """`n^n_dim` elements, all Cartesian or all deformed, with a boundary connection on every outer face.

    Interior connections run along each dimension between neighbours (direction {d, d}, {1, 0});
    boundary connections are deformed-type connections {d, d}, {sign, !sign} against a ghost face, as in the
    reference (src/Accessible_mesh.cpp:136-147, include/connection.hpp:346-366).
    If `keep_geometry_on_device`, the large metric arrays are returned as torch tensors on `device`.
    `lean` skips the host allocation of element data and face storage (benchmark-sized meshes live on the device only).

    Domain decomposition: with `blocks` = blocks per dimension and `block` = this rank's block coordinates the function builds
    ONE block of a global box of `blocks[d]*n` elements per dimension (unit spacing 1/(blocks[0]*n)): faces shared with a
    neighbouring block become cut connections against halo face slots, exactly as `partition.partition_mesh` would produce them
    from the undivided mesh (same table layout: interior, boundary, cut), and `mesh.halo` lists what to exchange with whom.
    Metric terms are evaluated on the block plus one layer of neighbours so that shared normals and vertex spacings are the
    same numbers on both sides of a cut.
    """
    from .partition import Halo
    nd, rs = n_dim, row_size
    blocks = tuple(blocks) if blocks is not None else (1,)*nd
    block = tuple(block) if block is not None else (0,)*nd
    gdim = [blocks[d]*n for d in range(nd)]
    origin = np.array([block[d]*n for d in range(nd)])
    E = n**nd
    h = 1./gdim[0]
    idx = np.stack(np.meshgrid(*[np.arange(n)]*nd, indexing="ij"), -1).reshape(E, nd)  # row-major, last fastest
    strides = np.array([n**(nd - 1 - d) for d in range(nd)])
    # outer faces of the block: physical boundary (ghost + boundary condition) or cut (halo + peer)
    b_elem, b_dim, b_sign = [], [], []
    c_elem, c_dim, c_sign, c_peer = [], [], [], []
    for d in range(nd):
        for sign in range(2):
            sel = np.nonzero(idx[:, d] == (n - 1 if sign else 0))[0]
            nb = list(block); nb[d] += 1 if sign else -1
            if 0 <= nb[d] < blocks[d]:
                c_elem.append(sel); c_dim.append(np.full(sel.size, d)); c_sign.append(np.full(sel.size, sign))
                c_peer.append(np.full(sel.size, block_rank(nb, blocks)))
            else:
                b_elem.append(sel); b_dim.append(np.full(sel.size, d)); b_sign.append(np.full(sel.size, sign))
    cat = lambda parts: np.concatenate(parts) if parts else np.zeros(0, np.int64)  # noqa: E731
    b_elem, b_dim, b_sign = cat(b_elem), cat(b_dim), cat(b_sign)
    c_elem, c_dim, c_sign, c_peer = cat(c_elem), cat(c_dim), cat(c_sign), cat(c_peer)
    n_bc, n_cut = b_elem.size, c_elem.size
    n_car, n_def = (0, E) if deformed else (E, 0)
    mesh = FlatMesh(nd, rs, n_car, n_def, n_ghost=n_bc + n_cut, n_extra_normal=0 if deformed else n_bc, with_ldg=with_ldg,
                    alloc_elem_data=not lean, alloc_faces=not lean)
    mesh.nom_size[:] = h
    mesh.box_n = n
    mesh.elem_index = idx + origin
    nfq = mesh.nfq
    # interior connections
    cons = []
    for d in range(nd):
        lo = np.nonzero(idx[:, d] < n - 1)[0]
        hi = lo + strides[d]
        if deformed:
            c = np.zeros((lo.size, 7), np.int32)
            c[:, 0] = lo*2*nd + 2*d + 1
            c[:, 1] = hi*2*nd + 2*d
            c[:, 2] = d; c[:, 3] = d; c[:, 4] = 1; c[:, 5] = 0
            c[:, 6] = lo*2*nd + 2*d + 1
        else:
            c = np.zeros((lo.size, 3), np.int32)
            c[:, 0] = lo*2*nd + 2*d + 1
            c[:, 1] = hi*2*nd + 2*d
            c[:, 2] = d
        cons.append(c)
    interior = np.concatenate(cons) if cons else None
    # boundary connections (always deformed-type)
    ghost = 2*nd*E + np.arange(n_bc)
    inside = b_elem*2*nd + 2*b_dim + b_sign
    bc = np.zeros((n_bc, 7), np.int32)
    bc[:, 0] = inside; bc[:, 1] = ghost
    bc[:, 2] = b_dim; bc[:, 3] = b_dim; bc[:, 4] = b_sign; bc[:, 5] = 1 - b_sign
    # cut connections: the element on the low side of the shared face is side 0, as in the undivided mesh
    halo_slot = 2*nd*E + n_bc + np.arange(n_cut)
    own = c_elem*2*nd + 2*c_dim + c_sign
    if deformed:
        cut = np.zeros((n_cut, 7), np.int32)
        cut[:, 0] = np.where(c_sign == 1, own, halo_slot); cut[:, 1] = np.where(c_sign == 1, halo_slot, own)
        cut[:, 2] = c_dim; cut[:, 3] = c_dim; cut[:, 4] = 1; cut[:, 5] = 0
        cut[:, 6] = own  # the shared (averaged) normal is stored on both elements' faces
        bc[:, 6] = inside  # element face normal slot == face slot numbering for an all-deformed mesh
        mesh.def_con = np.concatenate([interior, bc, cut]).astype(np.int32)
        mesh.n_cut_car, mesh.n_cut_def = 0, n_cut
        con_index = interior.shape[0] + np.arange(n_bc)
    else:
        cut = np.zeros((n_cut, 3), np.int32)
        cut[:, 0] = np.where(c_sign == 1, own, halo_slot); cut[:, 1] = np.where(c_sign == 1, halo_slot, own)
        cut[:, 2] = c_dim
        bc[:, 6] = np.arange(n_bc)  # connection-owned unit normals
        mesh.normals[np.arange(n_bc), b_dim, :] = 1.
        mesh.car_con = np.concatenate([interior, cut]).astype(np.int32)
        mesh.def_con = bc
        mesh.n_cut_car, mesh.n_cut_def = n_cut, 0
        con_index = np.arange(n_bc)
    mesh.pre_prolong = np.zeros(0, np.int32)
    mesh.bcs.append(dict(kind=bc_kind, inside_slot=inside.astype(np.int32), ghost_slot=ghost.astype(np.int32),
                         normal_slot=bc[:, 6].copy(), con_index=con_index.astype(np.int32), params=bc_params))
    if any(b > 1 for b in blocks):
        mesh.halo = Halo()
        for peer in np.unique(c_peer):
            sel = c_peer == peer  # both sides enumerate the shared plane in row-major order of the tangential indices
            mesh.halo.send[int(peer)] = own[sel].astype(np.int32)
            mesh.halo.recv[int(peer)] = halo_slot[sel].astype(np.int32)
    # geometry
    if deformed:
        dev = torch.device(device) if device is not None else torch.device("cpu")
        # extended index range: the block plus one layer of neighbours, clipped to the global box
        lo_ext = [max(int(origin[d]) - 1, 0) for d in range(nd)]
        hi_ext = [min(int(origin[d]) + n + 1, gdim[d]) for d in range(nd)]
        ext = [hi_ext[d] - lo_ext[d] for d in range(nd)]
        Ee = int(np.prod(ext))
        eidx = np.stack(np.meshgrid(*[np.arange(lo_ext[d], hi_ext[d]) for d in range(nd)], indexing="ij"), -1).reshape(Ee, nd)
        estr = np.array([int(np.prod(ext[d + 1:])) for d in range(nd)])
        # vertices of the extended range (global vertex coordinates -> positions through the analytic warp)
        vgrid = torch.stack(torch.meshgrid(*[torch.arange(lo_ext[d], hi_ext[d] + 1, dtype=torch.float64, device=dev)*h for d in range(nd)], indexing="ij"), -1)
        vgrid = default_warp(vgrid, warp_amplitude*h)
        vflat = vgrid.reshape(-1, nd)
        vstr = torch.tensor([int(np.prod([ext[k] + 1 for k in range(d + 1, nd)])) for d in range(nd)], device=dev)
        corner = torch.stack(torch.meshgrid(*[torch.arange(2, device=dev)]*nd, indexing="ij"), -1).reshape(-1, nd)
        eloc_t = torch.as_tensor(eidx - np.array(lo_ext), device=dev)
        vid = ((eloc_t[:, None, :] + corner[None, :, :])*vstr).sum(-1)  # (Ee, 2^nd) vertex ids within the extended range
        # which extended elements are the block's own, in the block's row-major order
        own_ext = ((idx + origin - np.array(lo_ext))*estr).sum(-1)
        is_own = np.zeros(Ee, bool); is_own[own_ext] = True
        out_dev = dev if keep_geometry_on_device else torch.device("cpu")
        refn = torch.empty((E, nd*nd, mesh.nq), dtype=torch.float64, device=out_dev)
        det = torch.empty((E, mesh.nq), dtype=torch.float64, device=out_dev)
        pos = torch.empty((E, nd, mesh.nq), dtype=torch.float64, device=out_dev)
        fn_ext = torch.empty((Ee, 2*nd, nd, nfq), dtype=torch.float64, device=out_dev)  # face normals of the extended range
        vtss = torch.empty((Ee, 2**nd), dtype=torch.float64, device=dev)
        nom = torch.full((Ee,), h, dtype=torch.float64, device=dev)
        local_of_ext = np.full(Ee, -1); local_of_ext[own_ext] = np.arange(E)
        for s in range(0, Ee, geometry_chunk):
            sl = slice(s, min(Ee, s + geometry_chunk))
            g = element_metrics(vflat[vid[sl]], nom[sl], basis)
            fn_ext[sl] = g["face_normals"].to(out_dev); vtss[sl] = g["vertex_tss"]
            lo_ids = local_of_ext[sl]
            keep = np.nonzero(lo_ids >= 0)[0]
            if keep.size:
                kt = torch.as_tensor(keep, device=dev)
                dst = torch.as_tensor(lo_ids[keep], device=out_dev)
                refn[dst] = g["ref_normals"][kt].to(out_dev); det[dst] = g["det"][kt].to(out_dev); pos[dst] = g["pos"][kt].to(out_dev)
        # shared face normal = sign-aware average of both sides (src/Solver.cpp:327-352); here both signs are +1
        for d in range(nd):
            lo = np.nonzero(eidx[:, d] < hi_ext[d] - 1)[0]
            hi = lo + int(estr[d])
            sel = is_own[lo] | is_own[hi]
            lo_t = torch.as_tensor(lo[sel], device=out_dev); hi_t = torch.as_tensor(hi[sel], device=out_dev)
            avg = 0.5*fn_ext[lo_t, 2*d + 1] + 0.5*fn_ext[hi_t, 2*d]
            fn_ext[lo_t, 2*d + 1] = avg
            fn_ext[hi_t, 2*d] = avg
        fn = fn_ext[torch.as_tensor(own_ext, device=out_dev)]
        del fn_ext
        # vertex time step scale: minimum over the elements sharing each vertex (src/Solver.cpp:380, src/Vertex.cpp vector_min)
        vmin = torch.full((int(np.prod([e + 1 for e in ext])),), float("inf"), dtype=torch.float64, device=dev)
        vmin.scatter_reduce_(0, vid.reshape(-1), vtss.reshape(-1), reduce="amin")
        vtss = vmin[vid[torch.as_tensor(own_ext, device=dev)]]
        mesh.vertex_tss = vtss.cpu().numpy()
        if keep_geometry_on_device:
            mesh.ref_normals, mesh.det, mesh.qpoint_pos = refn, det, pos
            mesh.normals = fn.reshape(E*2*nd, nd, nfq)
        else:
            mesh.ref_normals = refn.numpy(); mesh.det = det.numpy(); mesh.qpoint_pos = pos.numpy()
            mesh.normals = fn.reshape(E*2*nd, nd, nfq).numpy().copy()
    else:
        mesh.vertex_tss[:] = h/nd  # reference src/Element.cpp:17
        mesh.qpoint_pos = qpoint_positions_cartesian(idx + origin, mesh.nom_size, basis, nd)
    return mesh


# --------------------------------------------------------------------------------------
# synthetic "soup" for kernel-level parity: arbitrary orientations, hanging faces, random metrics
# --------------------------------------------------------------------------------------

def all_directions(n_dim):
    """every (i_dim0, i_dim1, sign0, sign1) a deformed connection can take"""
    return [Connection_direction([d0, d1], [s0, s1]) for d0 in range(n_dim) for d1 in range(n_dim)
            for s0 in range(2) for s1 in range(2)]


def soup_mesh(n_dim, row_size, rng, n_car=6, n_def=10, n_ref=4, with_ldg=True, with_wide=False):
    """A geometrically meaningless but structurally complete mesh: every connection direction appears, faces are paired
    at random, metrics are random (positive determinant). Kernels do not care about geometric validity, so this
    exercises all orientation / hanging-face / fallback paths in a few dozen elements."""
    nd, rs = n_dim, row_size
    ne = n_car + n_def
    free = {(e, f) for e in range(ne) for f in range(2*nd)}
    car_cons, def_cons, refs = [], [], []
    ghost_extra = 0
    normal_extra = 0

    def take(e, f):
        free.remove((e, f))
        return e*2*nd + f

    # cartesian connections among cartesian elements (and one car-def pair to exercise the unit-normal fallback)
    pairs = [(e, e + 1) for e in range(0, n_car - 1, 2)]
    if n_car and n_def:
        pairs.append((n_car - 1, n_car))
    for k, (e0, e1) in enumerate(pairs):
        d = k % nd
        if (e0, 2*d + 1) in free and (e1, 2*d) in free:
            car_cons.append([take(e0, 2*d + 1), take(e1, 2*d), d])
    # deformed connections covering every direction
    dirs = all_directions(nd)
    def_elems = list(range(n_car, ne))
    k = 0
    for direction in dirs:
        for _ in range(50):
            e0, e1 = rng.choice(def_elems, 2, replace=False) if n_def > 1 else (def_elems[0], def_elems[0])
            f0, f1 = direction.i_face(0), direction.i_face(1)
            if (e0, f0) in free and (e1, f1) in free and e0 != e1:
                s0, s1 = take(e0, f0), take(e1, f1)
                def_cons.append([s0, s1] + direction.as_list() + [(e0 - n_car)*2*nd + f0])
                k += 1
                break
    # hanging-node faces: coarse face on a free element face, fine mortar faces are extra slots, fine element faces free faces
    ghost_base = 2*nd*ne
    if nd >= 2:
        for r in range(n_ref):
            options = [(False, False), (True, False), (False, True), (True, True)] if nd == 3 else [(False, False), (True, False)]
            stretch = list(options[r % len(options)])
            nf = 2**(nd - 1)
            for d in range(nd - 1):
                nf //= 1 + stretch[d]
            cand = sorted(free)
            if len(cand) < nf + 1:
                break
            ce, cf = cand[rng.integers(len(cand))]
            coarse = take(ce, cf)
            fine = []
            for i in range(nf):
                mortar = ghost_base + ghost_extra; ghost_extra += 1
                cand = [c for c in sorted(free) if c[0] >= n_car] or sorted(free)
                fe, ff = cand[rng.integers(len(cand))]
                fslot = take(fe, ff)
                direction = Connection_direction([cf//2, ff//2], [cf % 2, ff % 2])
                if fe >= n_car:
                    nslot = (fe - n_car)*2*nd + ff
                else:
                    nslot = 2*nd*n_def + normal_extra; normal_extra += 1
                # mortar is side 0, fine element face side 1 (non-reversed refined connection)
                def_cons.append([mortar, fslot] + direction.as_list() + [nslot])
                fine.append(mortar)
            refs.append([coarse] + fine + [-1]*(4 - nf) + [int(stretch[0]), int(stretch[1])])
    # boundary connections on a few remaining faces
    bc_inside, bc_ghost, bc_normal, bc_con = [], [], [], []
    for (e, f) in sorted(free)[:max(4, len(free)//3)]:
        free.remove((e, f))
        g = ghost_base + ghost_extra; ghost_extra += 1
        d, sgn = f//2, f % 2
        if e >= n_car:
            nslot = (e - n_car)*2*nd + f
        else:
            nslot = 2*nd*n_def + normal_extra; normal_extra += 1
        bc_con.append(len(def_cons))
        def_cons.append([e*2*nd + f, g, d, d, sgn, 1 - sgn, nslot])
        bc_inside.append(e*2*nd + f); bc_ghost.append(g); bc_normal.append(nslot)
    mesh = FlatMesh(nd, rs, n_car, n_def, n_ghost=ghost_extra, n_extra_normal=normal_extra, with_ldg=with_ldg, with_wide=with_wide)
    mesh.car_con = np.array(car_cons, np.int32).reshape(-1, 3)
    mesh.def_con = np.array(def_cons, np.int32).reshape(-1, 7)
    mesh.ref_face = np.array(refs, np.int32).reshape(-1, 7)
    mesh.bcs.append(dict(kind=BC_COPY, inside_slot=np.array(bc_inside, np.int32), ghost_slot=np.array(bc_ghost, np.int32),
                         normal_slot=np.array(bc_normal, np.int32), con_index=np.array(bc_con, np.int32), params=None))
    # random but benign metrics: identity + perturbation
    nq, nfq = mesh.nq, mesh.nfq
    eye = np.eye(nd).reshape(1, nd*nd, 1)
    mesh.ref_normals = eye + 0.1*rng.standard_normal((n_def, nd*nd, nq))
    mesh.det = 1. + 0.1*rng.random((n_def, nq))
    # element face normals: unit fallback unless the face takes part in a deformed connection
    used = set(mesh.def_con[:, 6].tolist())
    for slot in range(mesh.n_normal_slot):
        if slot in used or slot >= 2*nd*n_def:
            base = np.zeros((nd, nfq))
            if slot < 2*nd*n_def:
                base[(slot % (2*nd))//2] = 1.
            else:
                base[rng.integers(nd)] = 1.
            mesh.normals[slot] = base + 0.1*rng.standard_normal((nd, nfq))
    mesh.nom_size = 0.5 + rng.random(ne)
    mesh.vertex_tss = 0.2 + rng.random((ne, 2**nd))
    return mesh


def random_flow_state(mesh, rng, mach=0.3):
    """thermodynamically admissible random Euler state + consistent random face data"""
    nd, nv, nq = mesh.n_dim, mesh.nv, mesh.nq
    ne = mesh.n_elem
    st = mesh.elem_data
    rho = 1. + 0.2*rng.random((ne, nq))
    vel = mach*340.*(rng.random((ne, nd, nq)) - 0.5)
    p = 1e5*(1. + 0.2*rng.random((ne, nq)))
    st[:, :nd] = rho[:, None, :]*vel
    st[:, nd] = rho
    st[:, nd + 1] = p/0.4 + 0.5*rho*(vel**2).sum(1)
    st[:, TSS_SLOT(nd)] = 0.5 + rng.random((ne, nq))
    c = cache_slot(nd, mesh.row_size)
    st[:, c:c + nv] = rng.standard_normal((ne, nv, nq))
    nfq = mesh.nfq
    ns = mesh.n_face_slot
    rho = 1. + 0.2*rng.random((ns, nfq))
    vel = mach*340.*(rng.random((ns, nd, nfq)) - 0.5)
    p = 1e5*(1. + 0.2*rng.random((ns, nfq)))
    f = mesh.face_state.reshape(ns, nv, nfq)
    f[:, :nd] = rho[:, None, :]*vel
    f[:, nd] = rho
    f[:, nd + 1] = p/0.4 + 0.5*rho*(vel**2).sum(1)
    if mesh.face_ldg is not None:
        mesh.face_ldg[:] = rng.standard_normal(mesh.face_ldg.shape)


# --------------------------------------------------------------------------------------
# locally refined Cartesian box: real 2:1 hanging-node faces (the topology class of the adaptive BASELINE configs)
# --------------------------------------------------------------------------------------

def refined_box_mesh(n_dim, row_size, n, basis, refine, bc_kind=BC_NONPENETRATION, bc_params=None, with_ldg=False):
    """`n^n_dim` Cartesian cells of size 1/n on the unit box; the cells with `refine[cell index tuple]` true are replaced by their
    2^n_dim children. Faces between a cell and the children of a refined neighbour are hanging-node faces built the way the
    reference's `Refined_connection<Element>` builds them (include/connection.hpp:195-262): the coarse element's face is the coarse
    face of a `Refined_face`, each child meets its own mortar face through a Cartesian connection (mortar on the coarse element's
    side: side 0 when the coarse element is below the face, side 1 -- "reversed" -- when it is above), children listed in row-major
    order of the face's tangential dimensions. Boundary faces get ghost faces and deformed-type boundary connections as in
    `box_mesh`. Adds `mesh.cell_volume` (per element) for conservation checks."""
    nd, rs = n_dim, row_size
    refine = np.asarray(refine, bool)
    assert refine.shape == (n,)*nd
    h = 1./n
    elems = []      # (size level: 0 coarse / 1 fine, integer origin in units of its own size)
    index_of = {}   # (level, tuple origin) -> element id
    for cell in np.ndindex(*(n,)*nd):
        if refine[cell]:
            for child in np.ndindex(*(2,)*nd):
                o = tuple(2*c + k for c, k in zip(cell, child))
                index_of[(1, o)] = len(elems); elems.append((1, o))
        else:
            index_of[(0, cell)] = len(elems); elems.append((0, cell))
    E = len(elems)
    nf = 2*nd
    car, refs, bc_rows = [], [], []
    extra = 0  # mortar + ghost slots after the element faces

    def slot(e, d, sign):
        return e*nf + 2*d + sign

    def tangential_children(d):
        dims = [k for k in range(nd) if k != d]
        return dims, list(np.ndindex(*(2,)*(nd - 1)))

    for e, (lvl, o) in enumerate(elems):
        size_cells = n*(2 if lvl else 1)  # number of cells of this level per dimension
        for d in range(nd):
            # neighbour across the positive face
            if o[d] + 1 < size_cells:
                nb = list(o); nb[d] += 1; nb = tuple(nb)
                if (lvl, nb) in index_of:  # same level
                    car.append([slot(e, d, 1), slot(index_of[(lvl, nb)], d, 0), d])
                elif lvl == 0:  # coarse below, children of the refined neighbour above: not reversed, mortar faces are side 0
                    dims, kids = tangential_children(d)
                    fine = []
                    for kid in kids:
                        fo = [0]*nd
                        fo[d] = 2*nb[d]
                        for t, k in zip(dims, kid):
                            fo[t] = 2*nb[t] + k
                        fe = index_of[(1, tuple(fo))]
                        mortar = E*nf + extra; extra += 1
                        car.append([mortar, slot(fe, d, 0), d])
                        fine.append(mortar)
                    refs.append([slot(e, d, 1)] + fine + [-1]*(4 - len(fine)) + [0, 0])
                else:  # fine below, coarse above: reversed, mortar faces are side 1; handled once per coarse face below
                    pass
            else:
                bc_rows.append((e, d, 1))
            if o[d] == 0:
                bc_rows.append((e, d, 0))
            elif lvl == 0:
                # neighbour across the negative face of a coarse element: if it is refined, this coarse face is a reversed hanging face
                nb = list(o); nb[d] -= 1; nb = tuple(nb)
                if (0, nb) not in index_of:
                    dims, kids = tangential_children(d)
                    fine = []
                    for kid in kids:
                        fo = [0]*nd
                        fo[d] = 2*nb[d] + 1
                        for t, k in zip(dims, kid):
                            fo[t] = 2*nb[t] + k
                        fe = index_of[(1, tuple(fo))]
                        mortar = E*nf + extra; extra += 1
                        car.append([slot(fe, d, 1), mortar, d])
                        fine.append(mortar)
                    refs.append([slot(e, d, 0)] + fine + [-1]*(4 - len(fine)) + [0, 0])
    n_bc = len(bc_rows)
    mesh = FlatMesh(nd, rs, E, 0, n_ghost=extra + n_bc, n_extra_normal=n_bc, with_ldg=with_ldg)
    size = np.array([h/2 if lvl else h for lvl, _ in elems])
    mesh.nom_size[:] = size
    mesh.vertex_tss[:] = (size/nd)[:, None]  # reference src/Element.cpp:17
    mesh.cell_volume = size**nd
    origin = np.array([o for _, o in elems])
    mesh.qpoint_pos = qpoint_positions_cartesian(origin, size, basis, nd)
    mesh.car_con = np.array(car, np.int32).reshape(-1, 3)
    mesh.ref_face = np.array(refs, np.int32).reshape(-1, 7)
    bc = np.zeros((n_bc, 7), np.int32)
    for i, (e, d, sign) in enumerate(bc_rows):
        bc[i] = [slot(e, d, sign), E*nf + extra + i, d, d, sign, 1 - sign, i]
        mesh.normals[i, d, :] = 1.
    mesh.def_con = bc
    mesh.bcs.append(dict(kind=bc_kind, inside_slot=bc[:, 0].copy(), ghost_slot=bc[:, 1].copy(), normal_slot=bc[:, 6].copy(),
                         con_index=np.arange(n_bc, dtype=np.int32), params=bc_params))
    return mesh
