"""Integer orientation / ordering tables consumed when the reference's pointer graph is flattened
to device index tables. All of these must agree bit-for-bit with the reference (north-star parity bar);
they are pinned against the golden vectors of the reference's own tests in tests/test_tables.py.
"""
import numpy as np


class Connection_direction:
    """reference include/Kernel_connection.hpp:7-37"""

    def __init__(self, i_dim, face_sign):
        self.i_dim = [int(i_dim[0]), int(i_dim[1])]
        self.face_sign = [int(bool(face_sign[0])), int(bool(face_sign[1]))]

    def i_face(self, i_side):
        return 2*self.i_dim[i_side] + self.face_sign[i_side]

    def flip_normal(self, i_side):
        return self.face_sign[i_side] == i_side

    def flip_tangential(self):
        return (self.i_dim[0] != self.i_dim[1]) and (self.flip_normal(0) == self.flip_normal(1))

    def transpose(self):
        return sorted(self.i_dim) == [0, 2] and self.i_dim[0] != self.i_dim[1]

    def packed(self):
        """bit-packed form stored in the device connection table: i_dim0 | i_dim1 << 2 | sign0 << 4 | sign1 << 5"""
        return self.i_dim[0] | (self.i_dim[1] << 2) | (self.face_sign[0] << 4) | (self.face_sign[1] << 5)

    def as_list(self):
        return [self.i_dim[0], self.i_dim[1], self.face_sign[0], self.face_sign[1]]


def cartesian_direction(i_dim):
    """`Con_dir<Element>` converted to a deformed direction (reference include/connection.hpp:30-38)"""
    return Connection_direction([i_dim, i_dim], [1, 0])


def face_vertex_inds(n_dim, direction):
    """permutation of the vertices of face 1 to match face 0 (reference src/connection.cpp:6-28)"""
    n_vert = 2**(n_dim - 1)
    inds = list(range(n_vert))
    if direction.flip_tangential():
        stride = 1
        if n_dim == 3:
            unused_dim = 3 - direction.i_dim[0] - direction.i_dim[1]
            if unused_dim > direction.i_dim[0]:
                stride = 2
        for i_vert in range(n_vert):
            inds[i_vert] += stride*(1 - 2*((i_vert//stride) % 2))
    if direction.transpose():
        inds[1], inds[2] = inds[2], inds[1]
    return inds


def vertex_inds(n_dim, direction):
    """vertices taking part in a deformed connection, physically aligned pairwise (reference src/connection.cpp:30-50)"""
    n_vert = 2**(n_dim - 1)
    inds = [[], []]
    for i_side in range(2):
        stride = 2**(n_dim - direction.i_dim[i_side] - 1)
        for i_vertex in range(2*n_vert):
            if (i_vertex//stride) % 2 == direction.face_sign[i_side]:
                inds[i_side].append(i_vertex)
    perm = face_vertex_inds(n_dim, direction)
    inds[1] = [inds[1][perm[i]] for i in range(n_vert)]
    return inds


def stretched_ind(n_dim, ind, stretch):
    """reference include/math.hpp:188-199"""
    stride = 1
    stretched = 0
    for i_dim in range(n_dim - 2, -1, -1):
        if not stretch[i_dim]:
            stretched += ((ind//2**(n_dim - 2 - i_dim)) % 2)*stride
            stride *= 2
    return stretched


def face_permutation(n_dim, row_size, direction):
    """index table p -> source index such that `matched[p] = original[table[p]]` reproduces
    `Face_permutation::match_faces` (transpose, then flip; reference include/Spatial.hpp:85-129).
    `restore` is the inverse permutation. For n_dim == 1 the face is a single point."""
    nfq = row_size**(n_dim - 1)
    idx = np.arange(nfq)
    if n_dim == 3:
        a = idx.reshape(row_size, row_size)
        if direction.transpose():
            a = a.T
        if direction.flip_tangential():
            # reversal along the fastest face index <=> Eigen colwise().reverse() of the column-major map
            fast = (direction.i_dim[0] > 3 - direction.i_dim[0] - direction.i_dim[1]) != direction.transpose()
            a = a[:, ::-1] if fast else a[::-1, :]
        idx = np.ascontiguousarray(a).reshape(-1)
    elif n_dim == 2:
        if direction.flip_tangential():
            idx = idx[::-1].copy()
    return idx.astype(np.int32)


def refined_fine_order(n_dim, direction, reverse, stretch):
    """For a `Refined_connection`, tell which fine element backs each fine connection / mortar face.

    Returns (elem_of_con, mortar_of_con, coarse_stretch) where for the i-th fine connection created
    (the order the connection views enumerate them) `elem_of_con[i]` indexes the caller's row-major list of
    fine elements and `mortar_of_con[i]` is the index into `Refined_face::fine`
    (reference include/connection.hpp:218-269)."""
    n_fine = 2**(n_dim - 1)
    any_str = False
    for i_dim in range(n_dim - 1):
        if stretch[i_dim]:
            n_fine //= 2
            any_str = True
    perm = face_vertex_inds(n_dim, direction)
    rev = int(bool(reverse))
    elem_of_con, mortar_of_con = [], []
    for i_face in range(n_fine):
        inds = [i_face, perm[i_face]]
        if any_str:
            d = direction
            inds[1] = int(i_face != (d.flip_tangential() and not stretch[int(2*d.i_dim[rev] > 3 - d.i_dim[1 - rev])]))
        elem_of_con.append(inds[1 - rev])
        mortar_of_con.append(inds[rev])
    trans = direction.transpose()
    coarse_stretch = [bool(stretch[int(trans)]), bool(stretch[int(not trans)])]
    return elem_of_con, mortar_of_con, coarse_stretch


def hanging_vertex_face_inds(n_dim, i_dim, is_positive):
    """vertex indices of an element lying on face (i_dim, is_positive) (reference src/Hanging_vertex_matcher.cpp:18-20)"""
    n_vert = 2**(n_dim - 1)
    stride = 2**(n_dim - 1 - i_dim)
    return [i_vert//stride*stride*2 + i_vert % stride + int(is_positive)*stride for i_vert in range(n_vert)]


def hanging_vertex_interp_inds(n_dim, n_elem, stretch):
    """for each fine element and face vertex, the index into the 3[x3] interpolated array
    (reference src/Hanging_vertex_matcher.cpp:31-39)"""
    n_vert = 2**(n_dim - 1)
    out = []
    for i_elem in range(n_elem):
        row = []
        for i_vert in range(n_vert):
            interp_ind = (i_elem*(not stretch[n_dim - 2])) % 2 + (i_vert % 2)*(1 + stretch[n_dim - 2])
            if n_dim == 3:
                interp_ind += ((i_elem*(not stretch[0]))//(1 + (not stretch[1])) + i_vert//2*(1 + stretch[0]))*3
            row.append(int(interp_ind))
        out.append(row)
    return out


def row_qpoint(n_dim, row_size, i_dim, i_fq, i_node):
    """`Row_index::i_qpoint` (reference include/Row_index.hpp:32-63)"""
    stride = row_size**(n_dim - 1 - i_dim)
    i_outer, i_inner = divmod(i_fq, stride)
    return i_outer*stride*row_size + i_inner + i_node*stride


def hanging_vertex_match(n_dim, vertex_vals, i_dim, is_positive, stretch=(False, False)):
    """`Hanging_vertex_matcher::match` on plain arrays (reference src/Hanging_vertex_matcher.cpp:13-41).

    vertex_vals: (n_fine_elem, 2^n_dim) array of per-vertex values of the fine elements, modified in place: the values on
    face (i_dim, is_positive) are replaced by the multilinear interpolant of the coarse-face corner values."""
    vals = np.asarray(vertex_vals)
    n_vert = 2**(n_dim - 1)
    inds = hanging_vertex_face_inds(n_dim, i_dim, is_positive)
    corner = np.array([vals[stretched_ind(n_dim, i, stretch), inds[i]] for i in range(n_vert)], dtype=np.float64)
    interp = np.array([[1., 0.], [.5, .5], [0., 1.]])
    cube = corner.reshape((2,)*(n_dim - 1))
    for axis in range(n_dim - 1):
        cube = np.moveaxis(np.tensordot(interp, cube, axes=([1], [axis])), 0, axis)
    flat = cube.reshape(-1)
    table = hanging_vertex_interp_inds(n_dim, vals.shape[0], stretch)
    for i_elem in range(vals.shape[0]):
        for i_vert in range(n_vert):
            vals[i_elem, inds[i_vert]] = flat[table[i_elem][i_vert]]
    return vals
