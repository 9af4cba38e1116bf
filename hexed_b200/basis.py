"""Basis tables for the B200 residual path.

Host-side mirror of the part of the reference's `Basis` interface the kernels consume
(reference include/Basis.hpp:16-66, src/Basis.cpp:6-14). The numbers are the output of the
reference's own generator (script/auto_generate.py), stored as decimal strings in
`data/basis_tables.json` by `oracle/gen_basis.py`; matrices are indexed M[i][j] exactly like the
`Eigen::MatrixXd` the reference returns.
"""
import json
import os
from dataclasses import dataclass

import numpy as np

_TABLES = None
MAX_ROW_SIZE = 8  # reference hexed_config.hpp: config::max_row_size


def _tables():
    global _TABLES
    if _TABLES is None:
        with open(os.path.join(os.path.dirname(__file__), "data", "basis_tables.json")) as f:
            _TABLES = json.load(f)
    return _TABLES


def _arr(x):
    return np.array(x, dtype=str).astype(np.float64)


@dataclass
class Basis:
    row_size: int
    node: np.ndarray
    weight: np.ndarray
    diff_mat: np.ndarray
    boundary: np.ndarray
    orthogonal: np.ndarray
    filter: np.ndarray
    prolong: np.ndarray   # [i_half][i][j]
    restrict: np.ndarray  # [i_half][i][j]
    min_eig_convection: float
    min_eig_diffusion: float
    quadratic_safety: float
    legendre_node: np.ndarray = None  # nodes of Gauss_legendre(row_size): pde::Advection uses them whatever the basis (include/pde.hpp:281)

    def max_cfl(self):
        """reference src/Basis.cpp:6-9"""
        return -2*self.quadratic_safety/self.min_eig_convection

    def step_ratio(self):
        """reference src/Basis.cpp:11-14"""
        return .5/self.quadratic_safety

    def packed(self):
        """flat double array handed across the C ABI (layout documented in include/hexed_b200.h)"""
        rs = self.row_size
        return np.concatenate([
            self.node, self.weight, self.diff_mat.ravel(), self.boundary.ravel(), self.orthogonal.ravel(),
            self.filter.ravel(), self.prolong.ravel(), self.restrict.ravel(),
            [self.min_eig_convection, self.min_eig_diffusion, self.quadratic_safety],
            self.legendre_node if self.legendre_node is not None else self.node,
        ]).astype(np.float64)


def _make(name, row_size):
    if not 2 <= row_size <= MAX_ROW_SIZE:
        raise RuntimeError("Not implemented for required row_size.")
    t = _tables()[name][str(row_size)]
    rs = row_size
    zero = np.zeros((2, rs, rs))
    return Basis(
        row_size=rs, node=_arr(t["node"]), weight=_arr(t["weight"]), diff_mat=_arr(t["diff_mat"]),
        boundary=_arr(t["boundary"]), orthogonal=_arr(t["orthogonal"]), filter=_arr(t["filter"]),
        prolong=_arr(t["prolong"]) if "prolong" in t else zero,
        restrict=_arr(t["restrict"]) if "restrict" in t else zero,
        min_eig_convection=float(t["min_eig_convection"]), min_eig_diffusion=float(t["min_eig_diffusion"]),
        quadratic_safety=float(_tables()["quadratic_safety"][name]),
        legendre_node=_arr(_tables()["Gauss_legendre"][str(rs)]["node"]),
    )


def gauss_legendre(row_size):
    """reference include/Gauss_legendre.hpp:15-34"""
    return _make("Gauss_legendre", row_size)


def gauss_lobatto(row_size):
    """reference include/Gauss_lobatto.hpp (prolong/restrict are not implemented there either)"""
    return _make("Gauss_lobatto", row_size)
