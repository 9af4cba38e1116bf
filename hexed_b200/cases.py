"""Synthetic initial / boundary states shared by the tests, smoke() and bench.py (no oracle dependency)."""
import numpy as np


def freestream_state(nd, mach=0.3):
    """uniform flow used for the freestream ghosts of the synthetic boxes"""
    rho, p = 1.2, 101325.
    vel = [mach*340.*(0.6 + 0.2*d) for d in range(nd)]
    return np.array([rho*v for v in vel] + [rho, p/0.4 + 0.5*rho*sum(v*v for v in vel)])


def density_wave(mesh, basis=None, mach=0.3, amplitude=0.1):
    """smooth admissible initial condition on the box meshes: travelling density wave
    (cf. reference test/test_Solver.cpp:103-124)"""
    nd = mesh.n_dim
    x = np.asarray(mesh.qpoint_pos)
    phase = sum(np.sin(2*np.pi*x[:, d] + 0.3*d) for d in range(nd))/nd
    rho = 1.2*(1 + amplitude*phase)
    vel = [mach*340.*(0.6 + 0.2*d) for d in range(nd)]
    p = 101325.*(1 + 0.05*np.cos(2*np.pi*x[:, 0]))
    st = mesh.state()
    ke = 0
    for d in range(nd):
        st[:, d] = rho*vel[d]
        ke = ke + 0.5*rho*vel[d]**2
    st[:, nd] = rho
    st[:, nd + 1] = p/0.4 + ke
    return mesh
