#!/usr/bin/env python3
"""Golden vectors from the REFERENCE'S OWN kernels (oracle/_ref/libhexed_ref.so: src/kernels_*.cpp + include/Spatial.hpp / pde.hpp compiled
unmodified, oracle/Makefile.ref) for machines that have neither /root/reference nor the prebuilt oracle/_ref: for each of the five PDEs one
small structurally complete soup mesh (every connection direction, hanging faces, Cartesian + deformed) is driven through the call sequence
Solver makes (tests/test_ref_oracle.py::drive) and the results are stored next to the seed that regenerates the inputs.

    python tests/golden/gen_ref_kernel_vectors.py        # needs /root/reference (builds oracle/_ref) -> tests/golden/ref_kernel_vectors.npz

tests/test_oracle_kat.py::test_reference_kernel_vectors checks the restated oracle against the file at 1e-13."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import hexed_b200 as hb  # noqa: E402
import pyoracle  # noqa: E402
from pyoracle import EULER, NAVIER_STOKES, ADVECTION, SMOOTH_AV, FIX_THERM_ADMIS  # noqa: E402
from test_ref_oracle import soup, drive  # noqa: E402

CASES = [(EULER, 2, 4, 11), (EULER, 3, 3, 12), (NAVIER_STOKES, 2, 4, 13), (NAVIER_STOKES, 3, 2, 14), (ADVECTION, 2, 3, 15),
         (SMOOTH_AV, 2, 3, 16), (FIX_THERM_ADMIS, 2, 3, 17), (FIX_THERM_ADMIS, 3, 2, 18)]
SOUP = dict(n_car=3, n_def=4, n_ref=2)


def run(oracle, pde, nd, rs, seed):
    basis = hb.gauss_legendre(rs)
    m = soup(nd, rs, seed, pde, **SOUP)
    oracle.compute_write_face(basis, m)
    dts = drive(oracle, m, basis, pde, n_steps=2)
    return m, dts


def main():
    ref = pyoracle.RefOracle()
    out = {"cases": np.array(CASES, dtype=np.int64)}
    for i, (pde, nd, rs, seed) in enumerate(CASES):
        m, dts = run(ref, pde, nd, rs, seed)
        out["elem_%d" % i] = m.elem_data
        out["face_state_%d" % i] = m.face_state
        out["face_ldg_%d" % i] = m.face_ldg
        out["face_wide_%d" % i] = m.face_wide
        out["dt_%d" % i] = np.array(dts)
    path = os.path.join(HERE, "ref_kernel_vectors.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
