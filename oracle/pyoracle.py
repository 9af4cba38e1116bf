"""ctypes binding of the CPU oracle (oracle/liboracle*.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline leg,
never by the hexed_b200 package. Works on `hexed_b200.mesh.FlatMesh` objects in place.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
EULER, NAVIER_STOKES, ADVECTION, SMOOTH_AV, FIX_THERM_ADMIS = range(5)


class ho_basis(C.Structure):
    _fields_ = [("row_size", C.c_int), ("node", C.c_double*8), ("weight", C.c_double*8),
                ("diff_mat", C.c_double*64), ("boundary", C.c_double*16), ("orthogonal", C.c_double*64),
                ("filter", C.c_double*64), ("prolong", C.c_double*128), ("restrict_", C.c_double*128),
                ("min_eig_convection", C.c_double), ("min_eig_diffusion", C.c_double), ("quadratic_safety", C.c_double)]


class ho_transport(C.Structure):
    _fields_ = [("const_val", C.c_double), ("ref_val", C.c_double), ("ref_temp", C.c_double),
                ("sqrt_ref_temp", C.c_double), ("temp_offset", C.c_double), ("is_viscous", C.c_int)]


class ho_options(C.Structure):
    _fields_ = [("dt", C.c_double), ("i_stage", C.c_int), ("compute_residual", C.c_int), ("use_filter", C.c_int)]


dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int)


class ho_mesh(C.Structure):
    _fields_ = [("n_dim", C.c_int), ("row_size", C.c_int), ("n_car", C.c_int), ("n_def", C.c_int), ("n_slot", C.c_int),
                ("elem_data", dp), ("nom_size", dp), ("vertex_tss", dp), ("uncert", dp), ("ref_normals", dp), ("det", dp),
                ("n_face_slot", C.c_int), ("face_state", dp), ("face_ldg", dp), ("face_wide", dp),
                ("n_normal_slot", C.c_int), ("normals", dp),
                ("n_car_con", C.c_int), ("car_con", ip), ("n_def_con", C.c_int), ("def_con", ip),
                ("n_ref", C.c_int), ("ref_face", ip)]


CALLBACK = C.CFUNCTYPE(None, C.c_void_p)


def inviscid():
    """Transport_model::inviscid() (reference include/Transport_model.hpp:40)"""
    return ho_transport(0., 0., 1., 1., 1., 0)


def constant(value):
    return ho_transport(value, 0., 1., 1., 1., 1)


def sutherland(ref_val, ref_temp, temp_offset):
    return ho_transport(0., ref_val, ref_temp, float(np.sqrt(ref_temp)), temp_offset, 1)


def build(target="liboracle.so", jobs=8):
    subprocess.run(["make", "-C", HERE, "-j%d" % jobs, target], check=True, stdout=subprocess.DEVNULL)


def _fill(dst, src):
    flat = np.ascontiguousarray(src, dtype=np.float64).ravel()
    for i, v in enumerate(flat):
        dst[i] = v


def pack_basis(b):
    rs = b.row_size
    o = ho_basis()
    o.row_size = rs
    _fill(o.node, b.node); _fill(o.weight, b.weight)

    def pad(m):
        p = np.zeros(m.shape[:-2] + (8, 8)); p[..., :m.shape[-2], :m.shape[-1]] = m
        return p
    _fill(o.diff_mat, pad(b.diff_mat)); _fill(o.orthogonal, pad(b.orthogonal)); _fill(o.filter, pad(b.filter))
    _fill(o.prolong, pad(b.prolong)); _fill(o.restrict_, pad(b.restrict))
    bd = np.zeros((2, 8)); bd[:, :rs] = b.boundary
    _fill(o.boundary, bd)
    o.min_eig_convection = b.min_eig_convection; o.min_eig_diffusion = b.min_eig_diffusion
    o.quadratic_safety = b.quadratic_safety
    return o


def _ptr(a, typ):
    if a is None:
        return C.cast(None, typ)
    assert a.flags["C_CONTIGUOUS"], "oracle needs contiguous arrays"
    return a.ctypes.data_as(typ)


class Oracle:
    def __init__(self, lib="liboracle.so", autobuild=True):
        path = os.path.join(HERE, lib)
        if not os.path.exists(path) and autobuild:
            build(lib)
        self.lib = C.CDLL(path)
        L = self.lib
        L.ho_compute_euler.argtypes = [C.POINTER(ho_basis), C.POINTER(ho_mesh), ho_options]
        L.ho_compute_advection.argtypes = [C.POINTER(ho_basis), C.POINTER(ho_mesh), ho_options, C.c_double]
        L.ho_compute_navier_stokes.argtypes = [C.POINTER(ho_basis), C.POINTER(ho_mesh), ho_options, CALLBACK, C.c_void_p, ho_transport, ho_transport]
        L.ho_compute_smooth_av.argtypes = [C.POINTER(ho_basis), C.POINTER(ho_mesh), ho_options, CALLBACK, C.c_void_p, C.c_double, C.c_double]
        L.ho_compute_fix_therm_admis.argtypes = [C.POINTER(ho_basis), C.POINTER(ho_mesh), ho_options, CALLBACK, C.c_void_p]
        L.ho_max_dt.argtypes = [C.c_int, C.POINTER(ho_basis), C.POINTER(ho_mesh), C.c_double, C.c_double, C.c_int, ho_transport, ho_transport, C.c_double, dp]
        L.ho_compute_write_face.argtypes = [C.c_int, C.POINTER(ho_basis), C.POINTER(ho_mesh)]
        L.ho_compute_prolong.argtypes = [C.c_int, C.POINTER(ho_basis), C.POINTER(ho_mesh), C.c_int, C.c_int]
        L.ho_compute_restrict.argtypes = [C.c_int, C.POINTER(ho_basis), C.POINTER(ho_mesh), C.c_int, C.c_int]
        L.ho_face_permutation.argtypes = [C.c_int, C.c_int, C.c_int, ip, C.c_int, dp]
        L.ho_stabilizing_art_visc.argtypes = [C.POINTER(ho_basis), C.POINTER(ho_mesh), C.c_double]
        L.ho_neighbor.argtypes = [C.c_int, C.c_int, C.POINTER(ho_mesh), C.c_int, ho_transport, ho_transport, C.c_double, C.c_double]
        L.ho_local.argtypes = [C.c_int, C.c_int, C.POINTER(ho_basis), C.POINTER(ho_mesh), ho_options, ho_transport, ho_transport, C.c_double, C.c_double]
        L.ho_neighbor_reconcile.argtypes = [C.c_int, C.c_int, C.POINTER(ho_mesh)]
        L.ho_reconcile_ldg_flux.argtypes = [C.c_int, C.c_int, C.POINTER(ho_basis), C.POINTER(ho_mesh), ho_options, ho_transport, ho_transport, C.c_double, C.c_double]
        L.ho_derivative.argtypes = [C.POINTER(ho_basis), C.c_int, dp, dp, dp]
        L.ho_bc_freestream.argtypes = [C.POINTER(ho_mesh), C.c_int, ip, dp]
        L.ho_bc_copy.argtypes = [C.POINTER(ho_mesh), C.c_int, ip, ip]
        L.ho_bc_nonpenetration.argtypes = [C.POINTER(ho_mesh), C.c_int, ip, ip, ip]
        L.ho_characteristics.argtypes = [C.c_int, dp, dp, dp, dp, dp]
        L.ho_bc_riemann_state.argtypes = [C.POINTER(ho_mesh), C.c_int, ip, ip, ip, dp, dp]
        L.ho_bc_riemann_flux.argtypes = [C.POINTER(ho_mesh), C.c_int, ip, ip, ip, dp]

    @staticmethod
    def _check(rc):
        if rc == 1:
            raise RuntimeError("demand for invalid kernel")
        if rc:
            raise RuntimeError("oracle error %d" % rc)

    def num_threads(self):
        return self.lib.ho_num_threads()

    @staticmethod
    def pack_mesh(m):
        """m: FlatMesh with numpy arrays. The returned struct borrows the arrays (keep `m` alive)."""
        o = ho_mesh()
        o.n_dim, o.row_size, o.n_car, o.n_def, o.n_slot = m.n_dim, m.row_size, m.n_car, m.n_def, m.n_slot
        o.elem_data = _ptr(m.elem_data, dp); o.nom_size = _ptr(m.nom_size, dp); o.vertex_tss = _ptr(m.vertex_tss, dp)
        o.uncert = _ptr(m.uncert, dp); o.ref_normals = _ptr(m.ref_normals, dp); o.det = _ptr(m.det, dp)
        o.n_face_slot = m.n_face_slot
        o.face_state = _ptr(m.face_state, dp); o.face_ldg = _ptr(m.face_ldg, dp); o.face_wide = _ptr(m.face_wide, dp)
        o.n_normal_slot = m.n_normal_slot; o.normals = _ptr(m.normals, dp)
        o.n_car_con = m.car_con.shape[0]; o.car_con = _ptr(m.car_con, ip)
        o.n_def_con = m.def_con.shape[0]; o.def_con = _ptr(m.def_con, ip)
        o.n_ref = m.ref_face.shape[0]; o.ref_face = _ptr(m.ref_face, ip)
        return o

    @staticmethod
    def opts(dt=1., i_stage=0, compute_residual=False, use_filter=False):
        return ho_options(dt, int(i_stage), int(compute_residual), int(use_filter))

    # ---- mirrors of the reference's kernels.hpp entry points ----
    def compute_euler(self, basis, m, **kw):
        self._check(self.lib.ho_compute_euler(pack_basis(basis), self.pack_mesh(m), self.opts(**kw)))

    def compute_advection(self, basis, m, advect_length, **kw):
        self._check(self.lib.ho_compute_advection(pack_basis(basis), self.pack_mesh(m), self.opts(**kw), advect_length))

    def _cb(self, flux_bc):
        return CALLBACK((lambda _: flux_bc()) if flux_bc else (lambda _: None))

    def compute_navier_stokes(self, basis, m, flux_bc, visc, therm_cond, **kw):
        self._check(self.lib.ho_compute_navier_stokes(pack_basis(basis), self.pack_mesh(m), self.opts(**kw), self._cb(flux_bc), None, visc, therm_cond))

    def compute_smooth_av(self, basis, m, flux_bc, diff_time, cheby_step, **kw):
        self._check(self.lib.ho_compute_smooth_av(pack_basis(basis), self.pack_mesh(m), self.opts(**kw), self._cb(flux_bc), None, diff_time, cheby_step))

    def compute_fix_therm_admis(self, basis, m, flux_bc, **kw):
        self._check(self.lib.ho_compute_fix_therm_admis(pack_basis(basis), self.pack_mesh(m), self.opts(**kw), self._cb(flux_bc), None))

    def max_dt(self, pde, basis, m, safety_conv, safety_diff, local_time, visc=None, therm_cond=None, advect_length=1.):
        out = C.c_double(0.)
        self._check(self.lib.ho_max_dt(pde, pack_basis(basis), self.pack_mesh(m), safety_conv, safety_diff, int(local_time),
                                      visc or inviscid(), therm_cond or inviscid(), advect_length, C.byref(out)))
        return out.value

    def compute_write_face(self, basis, m, pde=EULER):
        self._check(self.lib.ho_compute_write_face(pde, pack_basis(basis), self.pack_mesh(m)))

    def compute_prolong(self, basis, m, scale=False, offset=False, pde=EULER):
        self._check(self.lib.ho_compute_prolong(pde, pack_basis(basis), self.pack_mesh(m), int(scale), int(offset)))

    def compute_restrict(self, basis, m, scale=True, offset=False, pde=EULER):
        self._check(self.lib.ho_compute_restrict(pde, pack_basis(basis), self.pack_mesh(m), int(scale), int(offset)))

    def face_permutation(self, n_dim, row_size, n_var, direction, data, restore=False):
        d = (C.c_int*4)(*direction.as_list())
        self._check(self.lib.ho_face_permutation(n_dim, row_size, n_var, d, int(restore), _ptr(data, dp)))

    def stabilizing_art_visc(self, basis, m, char_speed):
        self._check(self.lib.ho_stabilizing_art_visc(pack_basis(basis), self.pack_mesh(m), char_speed))

    def neighbor(self, pde, deformed, m, i_stage=0, visc=None, therm_cond=None, p0=1., p1=1.):
        self._check(self.lib.ho_neighbor(pde, int(deformed), self.pack_mesh(m), i_stage, visc or inviscid(), therm_cond or inviscid(), p0, p1))

    def local(self, pde, deformed, basis, m, visc=None, therm_cond=None, p0=1., p1=1., **kw):
        self._check(self.lib.ho_local(pde, int(deformed), pack_basis(basis), self.pack_mesh(m), self.opts(**kw), visc or inviscid(), therm_cond or inviscid(), p0, p1))

    def neighbor_reconcile(self, pde, deformed, m):
        self._check(self.lib.ho_neighbor_reconcile(pde, int(deformed), self.pack_mesh(m)))

    def reconcile_ldg_flux(self, pde, deformed, basis, m, visc=None, therm_cond=None, p0=1., p1=1., **kw):
        self._check(self.lib.ho_reconcile_ldg_flux(pde, int(deformed), pack_basis(basis), self.pack_mesh(m), self.opts(**kw), visc or inviscid(), therm_cond or inviscid(), p0, p1))

    def derivative(self, basis, qpoint_vals, boundary_vals):
        q = np.ascontiguousarray(qpoint_vals, dtype=np.float64)
        bv = np.ascontiguousarray(boundary_vals, dtype=np.float64)
        out = np.zeros_like(q)
        self._check(self.lib.ho_derivative(pack_basis(basis), q.shape[0], _ptr(q, dp), _ptr(bv, dp), _ptr(out, dp)))
        return out

    def characteristics(self, state, direction, state1):
        """(eigvals[3], decomp[nv][3]) of reference include/pde.hpp:181-256 Characteristics(state, direction).decomp(state1)"""
        state = np.ascontiguousarray(state, dtype=np.float64)
        direction = np.ascontiguousarray(direction, dtype=np.float64)
        state1 = np.ascontiguousarray(state1, dtype=np.float64)
        nd = direction.size
        vals = np.zeros(3); dec = np.zeros((nd + 2, 3))
        self._check(self.lib.ho_characteristics(nd, _ptr(state, dp), _ptr(direction, dp), _ptr(state1, dp), _ptr(vals, dp), _ptr(dec, dp)))
        return vals, dec

    def is_admissible(self, m):
        """Solver::is_admissible (reference src/Solver.cpp:921-958) with thermo::admissible (src/thermo.cpp:6-18): returns
        (admissible, record[n_elem]); raises RuntimeError("state is not finite") where the reference's HEXED_ASSERT fires. numpy."""
        nd, nq, nfq, nv = m.n_dim, m.nq, m.nfq, m.n_dim + 2
        st = m.state()
        faces = m.face_state[:2*nd*m.n_elem].reshape(m.n_elem, 2*nd, nv, nfq)
        fine = []
        for row in np.asarray(m.ref_face).reshape(-1, 7):
            n_fine = 2**(nd - 1)
            for i in range(nd - 1):
                n_fine //= 1 + int(row[5 + i])
            fine += [int(s) for s in row[1:1 + n_fine]]
        fine_faces = m.face_state[fine].reshape(len(fine), nv, nfq)
        if not (np.isfinite(st).all() and np.isfinite(faces).all() and np.isfinite(fine_faces).all()):
            raise RuntimeError("state is not finite")
        ok_elem = (st[:, nd:] > 0.).all(axis=(1, 2)) & (faces[:, :, nd:] > 0.).all(axis=(1, 2, 3))
        ok_fine = bool((fine_faces[:, nd:] > 0.).all())
        return bool(ok_elem.all() and ok_fine), (~ok_elem).astype(np.int32)

    def apply_state_bcs(self, m):
        """ghost-state fill for the device-capable boundary conditions (reference src/Solver.cpp:56-67). Freestream, Copy and
        Nonpenetration run in the C oracle; Outflow, Pressure_outflow and No_slip are numpy restatements of
        src/Boundary_condition.cpp:184-211,367-385,465-468 (operation order kept; small meshes only)."""
        from hexed_b200.mesh import BC_FREESTREAM, BC_COPY, BC_NONPENETRATION, BC_OUTFLOW, BC_PRESSURE_OUTFLOW, BC_NO_SLIP, BC_RIEMANN_INVARIANTS
        pm = self.pack_mesh(m)
        nd, nfq, nv = m.n_dim, m.nfq, m.n_dim + 2
        for bc in m.bcs:
            n = bc["ghost_slot"].size
            ins, gh = bc["inside_slot"], bc["ghost_slot"]
            if n == 0:
                continue
            if bc["kind"] == BC_FREESTREAM:
                fs = np.ascontiguousarray(bc["params"], dtype=np.float64)
                self.lib.ho_bc_freestream(pm, n, _ptr(bc["ghost_slot"], ip), _ptr(fs, dp))
            elif bc["kind"] == BC_COPY:
                self.lib.ho_bc_copy(pm, n, _ptr(bc["inside_slot"], ip), _ptr(bc["ghost_slot"], ip))
            elif bc["kind"] == BC_NONPENETRATION:
                self.lib.ho_bc_nonpenetration(pm, n, _ptr(bc["inside_slot"], ip), _ptr(bc["ghost_slot"], ip), _ptr(bc["normal_slot"], ip))
            elif bc["kind"] == BC_RIEMANN_INVARIANTS:  # C oracle, oracle/characteristics.hpp
                fs = np.ascontiguousarray(bc["params"], dtype=np.float64)
                bc["cache"] = np.zeros((n, nv*nfq))
                self._check(self.lib.ho_bc_riemann_state(pm, n, _ptr(ins, ip), _ptr(gh, ip), _ptr(bc["normal_slot"], ip), _ptr(fs, dp),
                                                         _ptr(bc["cache"], dp)))
            elif bc["kind"] == BC_OUTFLOW:  # copy_state: both halves
                m.face_state[gh] = m.face_state[ins]
                if m.face_ldg is not None:
                    m.face_ldg[gh] = m.face_ldg[ins]
            elif bc["kind"] == BC_PRESSURE_OUTFLOW:
                f = m.face_state[ins].reshape(n, nv, nfq)
                nr = m.normals[bc["normal_slot"]]
                sign = 2*((ins % (2*nd)) % 2) - 1
                dot = np.zeros((n, nfq)); nsq = np.zeros((n, nfq)); msq = np.zeros((n, nfq))
                for d in range(nd):
                    dot = dot + f[:, d]*nr[:, d]; nsq = nsq + nr[:, d]*nr[:, d]; msq = msq + f[:, d]*f[:, d]
                mass = f[:, nd]
                nrml_veloc = dot/mass/np.sqrt(nsq)
                kin = .5*msq/mass
                pres = np.maximum(.4*(f[:, nd + 1] - kin), 0.)
                sound = np.sqrt(1.4*pres/mass)
                g = f.copy()
                g[:, nd + 1] = np.where(nrml_veloc*sign[:, None] < sound, bc["params"][0]/.4 + kin, f[:, nd + 1])
                m.face_state[gh] = g.reshape(n, -1)
            elif bc["kind"] == BC_NO_SLIP:
                f = m.face_state[ins].reshape(n, nv, nfq)
                g = f.copy()
                g[:, :nd] = -f[:, :nd]
                ge = bc["params"][1]*f[:, nd] if int(bc["params"][0]) == 1 else f[:, nd + 1]
                g[:, nd + 1] = ge*ge/f[:, nd + 1]
                m.face_state[gh] = g.reshape(n, -1)
                bc["cache"] = ((g + f)/2).reshape(n, -1)

    def apply_flux_bcs(self, m):
        """flux boundary conditions of the same kinds (reference src/Solver.cpp:69-81): Freestream/Copy::apply_flux = copy_state
        (src/Boundary_condition.cpp:12-23,304-305,455-458), Nonpenetration::apply_flux (:329-341), Outflow / Pressure_outflow
        (:213-221,470-477), No_slip (:387-418) with its three Thermal_bc kinds (include/Boundary_condition.hpp:140-186);
        numpy, small meshes only"""
        from hexed_b200.mesh import BC_FREESTREAM, BC_COPY, BC_NONPENETRATION, BC_OUTFLOW, BC_PRESSURE_OUTFLOW, BC_NO_SLIP, BC_RIEMANN_INVARIANTS
        nd, nfq = m.n_dim, m.nfq
        for bc in m.bcs:
            ins, gh = bc["inside_slot"], bc["ghost_slot"]
            if ins.size == 0:
                continue
            if bc["kind"] in (BC_FREESTREAM, BC_COPY):
                m.face_ldg[gh] = m.face_ldg[ins]
                m.face_state[gh] = m.face_state[ins]
            elif bc["kind"] == BC_NONPENETRATION:
                g = -m.face_ldg[ins].reshape(-1, nd + 2, nfq)
                n = m.normals[bc["normal_slot"]]
                dot = (g[:, :nd]*n).sum(1)
                nsq = (n*n).sum(1)
                g[:, :nd] -= 2*dot[:, None, :]*n/nsq[:, None, :]
                m.face_ldg[gh] = g.reshape(len(gh), -1)
            elif bc["kind"] == BC_RIEMANN_INVARIANTS:
                self._check(self.lib.ho_bc_riemann_flux(self.pack_mesh(m), ins.size, _ptr(ins, ip), _ptr(gh, ip), _ptr(bc["normal_slot"], ip),
                                                        _ptr(bc["cache"], dp)))
            elif bc["kind"] in (BC_OUTFLOW, BC_PRESSURE_OUTFLOW):
                m.face_ldg[gh] = -m.face_ldg[ins]
            elif bc["kind"] == BC_NO_SLIP:
                p = bc["params"]
                f = m.face_ldg[ins].reshape(-1, nd + 2, nfq)
                g = f.copy()
                g[:, nd] = -f[:, nd]
                nr = m.normals[bc["normal_slot"]]
                nsq = np.zeros((len(ins), nfq))
                for d in range(nd):
                    nsq = nsq + nr[:, d]*nr[:, d]
                nrm = np.sqrt(nsq)
                flux_sign = (2*((ins % (2*nd)) % 2) - 1)[:, None]
                in_e = f[:, nd + 1]
                kind = int(p[0])
                if kind == 0:
                    ghf = p[1]
                elif kind == 1:
                    ghf = in_e*flux_sign/nrm
                else:
                    sc = bc["cache"].reshape(-1, nd + 2, nfq)
                    temp = sc[:, nd + 1]*.4/sc[:, nd]/287.05287
                    ghf = p[1]*p[5]*(temp*temp*temp*temp) + p[2]*(temp - p[3])
                g[:, nd + 1] = p[4]*(nrm*flux_sign*ghf - in_e) + in_e
                m.face_ldg[gh] = g.reshape(len(gh), -1)


REF_LIB = os.path.join(HERE, "_ref", "libhexed_ref.so")


def build_ref(jobs=8):
    """compile the reference's own kernel sources (oracle/Makefile.ref); needs /root/reference, i.e. only works in the build container"""
    subprocess.run(["make", "-C", HERE, "-f", "Makefile.ref", "-j%d" % jobs], check=True, stdout=subprocess.DEVNULL)


def ref_available():
    return os.path.exists(REF_LIB)


class _RefLib:
    """attribute proxy: `ho_x` resolves to the reference build's `hr_x` where it exists, else to the restated oracle's `ho_x`
    (boundary-condition fills, which oracle/_ref does not contain)"""

    def __init__(self, ref, port):
        self._ref, self._port = ref, port

    def __getattr__(self, name):
        if name.startswith("ho_"):
            try:
                return getattr(self._ref, "hr_" + name[3:])
            except AttributeError:
                pass
        return getattr(self._port, name)


class RefOracle(Oracle):
    """Same interface as `Oracle`, but every kernel call lands in oracle/_ref/libhexed_ref.so: the REFERENCE's own
    src/kernels_*.cpp + include/Spatial.hpp / pde.hpp compiled unmodified (oracle/Makefile.ref, oracle/ref_harness.cpp) on the
    reference's own generated Gauss_legendre basis. The `basis` argument of the methods is accepted and ignored."""

    def __init__(self, lib="liboracle.so"):
        if not ref_available():
            if os.path.isdir("/root/reference"):
                build_ref()
            else:
                raise FileNotFoundError(REF_LIB)
        Oracle.__init__(self, lib)
        port = self.lib
        self.ref = C.CDLL(REF_LIB)
        self.lib = _RefLib(self.ref, port)
        R = self.ref
        for name in ("compute_euler", "compute_advection", "compute_navier_stokes", "compute_smooth_av", "compute_fix_therm_admis", "max_dt",
                     "compute_write_face", "compute_prolong", "compute_restrict", "face_permutation", "stabilizing_art_visc", "derivative",
                     "characteristics"):
            getattr(R, "hr_" + name).argtypes = getattr(port, "ho_" + name).argtypes
        R.hr_basis.argtypes = [C.c_int, C.POINTER(ho_basis)]
        R.hr_basis_max_cfl.argtypes = [C.c_int]; R.hr_basis_max_cfl.restype = C.c_double
        R.hr_basis_step_ratio.argtypes = [C.c_int]; R.hr_basis_step_ratio.restype = C.c_double
        R.hr_chebyshev_step.argtypes = [C.c_int, C.c_int]; R.hr_chebyshev_step.restype = C.c_double
        R.hr_view_create.argtypes = [C.POINTER(ho_mesh)]; R.hr_view_create.restype = C.c_void_p
        R.hr_view_destroy.argtypes = [C.c_void_p]; R.hr_view_destroy.restype = None
        R.hr_view_store.argtypes = [C.c_void_p]; R.hr_view_store.restype = None
        R.hr_view_compute_euler.argtypes = [C.c_void_p, ho_options]
        R.hr_view_compute_navier_stokes.argtypes = [C.c_void_p, ho_options, ho_transport, ho_transport]
        R.hr_view_max_dt_euler.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_int, dp]
        R.hr_view_bc_freestream.argtypes = [C.c_void_p, C.c_int, ip, dp]

    def num_threads(self):
        return self.ref.hr_num_threads()

    def basis_tables(self, row_size):
        """the reference's generated Gauss-Legendre tables as a dict of numpy arrays (M[i][j] mathematical indexing)"""
        b = ho_basis()
        self._check(self.ref.hr_basis(row_size, C.byref(b)))
        rs = row_size

        def arr(field, *shape):
            return np.array(list(field), dtype=np.float64).reshape(shape)
        return dict(node=arr(b.node, 8)[:rs], weight=arr(b.weight, 8)[:rs], diff_mat=arr(b.diff_mat, 8, 8)[:rs, :rs],
                    boundary=arr(b.boundary, 2, 8)[:, :rs], orthogonal=arr(b.orthogonal, 8, 8)[:rs, :rs], filter=arr(b.filter, 8, 8)[:rs, :rs],
                    prolong=arr(b.prolong, 2, 8, 8)[:, :rs, :rs], restrict=arr(b.restrict_, 2, 8, 8)[:, :rs, :rs],
                    min_eig_convection=b.min_eig_convection, min_eig_diffusion=b.min_eig_diffusion, quadratic_safety=b.quadratic_safety,
                    max_cfl=self.ref.hr_basis_max_cfl(rs), step_ratio=self.ref.hr_basis_step_ratio(rs))


# ---------------------------------------------------------------------------------------------------------------------
# metric terms: numpy restatement of Deformed_element::position / set_jacobian, one element at a time with
# plain loops (small cases only). TEST INFRASTRUCTURE like everything else in this directory.
# ---------------------------------------------------------------------------------------------------------------------
def _dimension_matvec(mat, vec, i_dim):
    """math::dimension_matvec (reference src/math.cpp:58-81): vec viewed as [cols^i_dim][cols][rest], contract the middle index"""
    mat = np.atleast_2d(mat)
    rows, cols = mat.shape
    n_rows = cols**(i_dim + 1)
    n_cols = vec.size//n_rows
    v = vec.reshape(n_rows//cols, cols, n_cols)
    return np.einsum("rc,ocn->orn", mat, v).reshape(-1)


def element_position(vert, node_adj, basis, q):
    """Deformed_element::position (reference src/Deformed_element.cpp:15-58): vert (2^nd, nd), node_adj (2*nd*nfq,)"""
    n_vert, nd = vert.shape
    rs = basis.row_size
    nq = rs**nd
    nfq = nq//rs
    pos = np.zeros(nd)
    for i_dim in range(nd):
        vert_pos = vert[:, i_dim].copy()
        adjustments = [vert_pos.copy() for _ in range(nd)]
        for j_dim in range(nd - 1, -1, -1):
            stride = rs**(nd - j_dim - 1)
            node = basis.node[(q//stride) % rs]
            interp = np.array([[1. - node, node]])
            vert_pos = _dimension_matvec(interp, vert_pos, j_dim)
            fq, face_stride = 0, nq//rs//rs if nd > 1 else 0
            for k_dim in range(nd):
                if k_dim != j_dim:
                    fq += ((q//rs**(nd - k_dim - 1)) % rs)*face_stride
                    face_stride //= rs
            for k_dim in range(nd):
                i_adjust = 2*j_dim*nfq + fq
                adjust = np.array([[node_adj[i_adjust], node_adj[i_adjust + nfq]]])
                difference = np.array([[-1., 1.], [-1., 1.]])
                contract = (adjust*interp) @ difference if j_dim == k_dim else interp
                adjustments[k_dim] = _dimension_matvec(contract, adjustments[k_dim], j_dim)
        pos[i_dim] = vert_pos[0]
        for j_dim in range(nd):
            pos[i_dim] += adjustments[j_dim][0]
    return pos


def set_jacobian(vert, node_adj, nom, basis):
    """Deformed_element::set_jacobian (reference src/Deformed_element.cpp:60-136) for one element.
    returns dict: jac (nd, nd, nq) [i][j] = d pos_i/d ref_j/nom, ref_normals (nd*nd, nq), det (nq,), face_normals (2*nd, nd, nfq),
    vertex_tss (2^nd,). Determinants by numpy (LU with partial pivoting, like Eigen's dynamic-size determinant())."""
    n_vert, nd = vert.shape
    rs = basis.row_size
    nq = rs**nd
    nfq = nq//rs
    diff, bnd = np.asarray(basis.diff_mat), np.asarray(basis.boundary)
    pos = np.array([element_position(vert, node_adj, basis, q) for q in range(nq)])  # (nq, nd)
    jac = np.zeros((nd, nd, nq))
    for i in range(nd):
        for j in range(nd):
            jac[i, j] = _dimension_matvec(diff, pos[:, i].copy(), j)/nom
    refn = np.zeros((nd*nd, nq))
    det = np.zeros(nq)
    for q in range(nq):
        J = jac[:, :, q]
        det[q] = np.linalg.det(J)
        for i in range(nd):
            for j in range(nd):
                c = J.copy()
                c[:, i] = 0.
                c[j, i] = 1.
                refn[i*nd + j, q] = np.linalg.det(c)
    fn = np.zeros((2*nd, nd, nfq))
    for i in range(nd):
        for sign in range(2):
            fj = np.array([[_dimension_matvec(bnd[sign:sign + 1], jac[a, b].copy(), i) for b in range(nd)] for a in range(nd)])  # (nd, nd, nfq)
            for fq in range(nfq):
                for j in range(nd):
                    c = fj[:, :, fq].copy()
                    c[:, i] = 0.
                    c[j, i] = 1.
                    fn[2*i + sign, j, fq] = np.linalg.det(c)

    def to_vertices(field):
        v = field.copy()
        for d in range(nd - 1, -1, -1):  # hypercube_matvec(boundary, .): contract every dimension with the 2 x rs boundary matrix
            v = _dimension_matvec(bnd, v, d)   # (last dimension first, so the leading dimensions still have row_size entries)
        return v
    vn = np.array([to_vertices(refn[k]) for k in range(nd*nd)])  # (nd*nd, n_vert)
    vd = to_vertices(det)
    vtss = np.zeros(n_vert)
    for iv in range(n_vert):
        norm_sum = 0.
        for i in range(nd):
            norm_sum += np.sqrt(sum(vn[i*nd + j, iv]**2 for j in range(nd)))
        vtss[iv] = nom*vd[iv]/norm_sum
    return dict(jac=jac, ref_normals=refn, det=det, face_normals=fn, vertex_tss=vtss, pos=pos)


# ---------------------------------------------------------------------------------------------------------------------
# pointwise loops of the artificial-viscosity pipelines, numpy on a FlatMesh. TEST INFRASTRUCTURE.
# ---------------------------------------------------------------------------------------------------------------------
def av_scale_velocity(m, restore=False):
    """reference src/Solver.cpp:467-478 (restore: :567-571)"""
    nd = m.n_dim
    st = m.state()
    scale = np.sqrt(2*st[:, nd]*st[:, nd + 1])
    for d in range(nd):
        st[:, d] = st[:, d]*scale if restore else st[:, d]/scale


def av_project_forcing(m, weights, orth):
    """reference src/Solver.cpp:527-541"""
    nd, rs = m.n_dim, m.row_size
    adv = m.elem_data[:, nd + 9:nd + 9 + rs]
    proj = np.zeros_like(adv[:, 0])
    for i in range(rs):
        proj = proj + adv[:, i]*weights[i]*orth[i]
    st = m.state()
    m.elem_data[:, nd + 5] = proj*proj*2*st[:, nd + 1]/st[:, nd]


def av_finish(m, mult, us_max, n_real, weights):
    """reference src/Solver.cpp:551-573; returns art_visc_residual"""
    nd, rs, nq = m.n_dim, m.row_size, m.nq
    wq = np.ones(nq)
    q = np.arange(nq)
    for d in range(nd):
        wq = wq*np.asarray(weights)[(q//rs**(nd - 1 - d)) % rs]
    f = mult*m.elem_data[:, nd + 5 + n_real]
    new_av = us_max*f/(us_max + f)
    vol = np.asarray(m.nom_size)**nd
    resid = (((m.elem_data[:, nd + 3] - new_av)**2)*wq[None, :]*vol[:, None]).sum()
    m.elem_data[:, nd + 3] = new_av
    av_scale_velocity(m, restore=True)
    return float(np.sqrt(resid))


def interp_vertices(m, target, vertex_values, interp):
    """math::hypercube_matvec(interp, vertex values) -> bulk (0) or laplacian (1) AV coefficient (reference src/Solver.cpp:652-656,1021-1031)"""
    nd, rs, nq = m.n_dim, m.row_size, m.nq
    interp = np.asarray(interp).reshape(rs, 2)
    for e in range(m.n_elem):
        v = np.asarray(vertex_values[e], dtype=np.float64).copy()
        for d in range(nd - 1, -1, -1):
            v = _dimension_matvec(interp, v, d)
        m.elem_data[e, nd + 3 + target] = v


def av_swap(m):
    nd = m.n_dim
    tmp = m.elem_data[:, nd + 3].copy()
    m.elem_data[:, nd + 3] = m.elem_data[:, nd + 4]
    m.elem_data[:, nd + 4] = tmp


def av_elwise_ramp(m, scale):
    """Solver::update_art_visc_elwise, reference src/Solver.cpp:590-601, on m.uncert (one value per element)"""
    rs = m.row_size
    ramp_center = -4 - 4.25*np.log(rs - 1)/np.log(10)
    half_width = 0.5
    with np.errstate(divide="ignore", invalid="ignore"):
        u = 2*np.log(m.uncert)/np.log(10)
        out = np.where(~(u > ramp_center - half_width), 0., np.where(u >= ramp_center + half_width, 1.,
                                                                     .5*(1 + np.sin(np.pi*(u - ramp_center)/2/half_width))))
    m.uncert[:] = out*scale


def av_elwise_forcing(m, restore):
    """the point loops of the PDE-based branch, reference src/Solver.cpp:603-612 (restore false) and :614-619 (restore true)"""
    nd = m.n_dim
    lap, f0 = nd + 4, nd + 5  # laplacian_av_coef, first art_visc_forcing slot (reference src/Element.cpp:114-142)
    if restore:
        m.elem_data[:, lap] = m.elem_data[:, f0 + 1]
    else:
        m.elem_data[:, f0] = m.uncert[:, None]
        m.elem_data[:, f0 + 1] = m.elem_data[:, lap]


def av_elwise_vertices(m, elem_vertex, n_vertex, matchers, interp):
    """the vertex-based branch, reference src/Solver.cpp:620-632: element value on its vertices, share_vertex_data(vector_max),
    laplacian_av_coef = hypercube_matvec(interp, vertex values)"""
    n_vert = 2**m.n_dim
    vals = np.repeat(m.uncert[:, None], n_vert, axis=1).astype(np.float64)
    vals = share_vertex_data(vals, elem_vertex, n_vertex, matchers, m.n_dim, True)
    interp_vertices(m, 1, vals, interp)
    return vals


def apply_aux_bcs(m, mode):
    """boundary loops of the AV / admissibility pipelines, numpy (TEST INFRASTRUCTURE). mode 0: Flow_bc::apply_advection and overrides
    (reference src/Boundary_condition.cpp:24-41,346-369,429-448,460-463) on the wide faces; 1: Flow_bc::apply_diffusion (:43-52) /
    the ghost copy of Solver.cpp:1063-1068; 2: Flow_bc::flux_diffusion (:54-60) / Solver::apply_fta_flux_bcs (Solver.cpp:103-115)"""
    from hexed_b200.mesh import BC_COPY, BC_NONPENETRATION, BC_NO_SLIP
    nd, rs, nfq, nv = m.n_dim, m.row_size, m.nfq, m.n_dim + 2
    for bc in m.bcs:
        ins, gh = bc["inside_slot"], bc["ghost_slot"]
        if ins.size == 0:
            continue
        if mode == 1:
            m.face_state[gh] = m.face_state[ins]
        elif mode == 2:
            m.face_ldg[gh] = -m.face_ldg[ins]
        else:
            f = m.face_wide[ins].reshape(-1, nd + rs, nfq)
            g = m.face_wide[gh].reshape(-1, nd + rs, nfq).copy()
            if bc["kind"] == BC_COPY:
                n_copy = min(2*nv, nd + rs)
                g[:, :n_copy] = f[:, :n_copy]
            elif bc["kind"] == BC_NONPENETRATION:
                n = m.normals[bc["normal_slot"]]
                dot = np.zeros((ins.size, nfq)); nsq = np.zeros((ins.size, nfq))
                for d in range(nd):
                    dot = dot + f[:, d]*n[:, d]; nsq = nsq + n[:, d]*n[:, d]
                for d in range(nd):
                    g[:, d] = f[:, d] - 2*dot*n[:, d]/nsq
                g[:, nd:] = f[:, nd:]
            elif bc["kind"] == BC_NO_SLIP:
                g[:, :nd] = -f[:, :nd]
                g[:, nd:] = f[:, nd:]
            else:
                g[:, :nd] = f[:, :nd]
                g[:, nd:] = 2. - f[:, nd:]
            m.face_wide[gh] = g.reshape(ins.size, -1)


def share_vertex_data(elem_vals, elem_vertex, n_vertex, matchers, n_dim, is_max):
    """Solver::share_vertex_data (reference src/Solver.cpp:35-54) on an (n_elem, 2^nd) array, in place: vertex-wise min/max over the
    sharing elements, then every Hanging_vertex_matcher (rows {i_dim, is_positive, stretch0, stretch1, fine elements x4, -1 padded})
    through the golden-vector-pinned hexed_b200.tables.hanging_vertex_match. numpy, TEST INFRASTRUCTURE."""
    from hexed_b200.tables import hanging_vertex_match
    ev = np.asarray(elem_vertex)
    red = np.full(n_vertex, -np.inf if is_max else np.inf)
    (np.maximum if is_max else np.minimum).at(red, ev.reshape(-1), elem_vals.reshape(-1))
    elem_vals[:] = red[ev]
    for row in np.asarray(matchers).reshape(-1, 8):
        fine = [int(e) for e in row[4:] if e >= 0]
        sub = elem_vals[fine].copy()
        hanging_vertex_match(n_dim, sub, int(row[0]), bool(row[1]), (bool(row[2]), bool(row[3])))
        elem_vals[fine] = sub
    return elem_vals


def fix_admis_spread(m, record, elem_vertex, n_vertex, matchers, interp):
    """the spreading step of Solver::fix_admissibility (reference src/Solver.cpp:1000-1038)"""
    nd = m.n_dim
    v = np.repeat(np.asarray(record, dtype=np.float64)[:, None], 2**nd, axis=1)
    share_vertex_data(v, elem_vertex, n_vertex, matchers, nd, True)
    v[:] = v.max(axis=1, keepdims=True)
    share_vertex_data(v, elem_vertex, n_vertex, matchers, nd, True)
    interp_vertices(m, 1, v, interp)
    av_swap(m)
    return v


def calc_shared_normals(oracle, basis, m):
    """the connection passes of Solver::calc_jacobian (reference src/Solver.cpp:287-369) on a FlatMesh whose element faces hold their
    element's normal in the first n_dim*nfq doubles (as Deformed_element::set_jacobian leaves them). numpy, TEST INFRASTRUCTURE; the
    face permutation is hexed_b200.tables.face_permutation (pinned by the reference's Face_permutation tests)."""
    from hexed_b200.tables import Connection_direction, face_permutation
    nd, rs, nfq, ne = m.n_dim, m.row_size, m.nfq, m.n_elem
    n_elem_face, def_first = 2*nd*ne, 2*nd*m.n_car
    mortar = set()
    for row in np.asarray(m.ref_face).reshape(-1, 7):
        n_fine = 2**(nd - 1)
        for k in range(nd - 1):
            n_fine //= 1 + int(row[5 + k])
        mortar.update(int(s) for s in row[1:1 + n_fine])
    if len(mortar):
        oracle.compute_prolong(basis, m, scale=True)
    face = lambda s: m.face_state[s].reshape(nd + 2, nfq)[:nd]  # noqa: E731  (a view)
    rows = np.asarray(m.def_con).reshape(-1, 7)
    dirs = [Connection_direction(r[2:4], r[4:6]) for r in rows]
    tables = [face_permutation(nd, rs, d) for d in dirs]
    for r, d, t in zip(rows, dirs, tables):  # fine connections (:301-317)
        m0, m1 = int(r[0]) in mortar, int(r[1]) in mortar
        if m0 == m1:
            continue
        mort, other = (face(r[1]), face(r[0])) if m1 else (face(r[0]), face(r[1]))
        sign = 1 - 2*(d.flip_normal(0) != d.flip_normal(1))
        other[:, t] = sign*mort
    for r in rows:  # boundary ghosts (:319-326)
        if r[1] >= n_elem_face and int(r[1]) not in mortar and r[0] < n_elem_face:
            face(r[1])[:] = face(r[0])
    for r, d, t in zip(rows, dirs, tables):  # shared normal (:328-355)
        s = [1 - 2*d.flip_normal(0), 1 - 2*d.flip_normal(1)]
        f0, f1 = face(r[0]), face(r[1])
        n = np.zeros((nd, nfq))
        n = n + 0.5*s[0]*f0
        n = n + 0.5*s[1]*f1[:, t]
        f0[:] = s[0]*n
        f1[:, t] = s[1]*n
        if not (def_first <= r[1] < n_elem_face and r[6] == r[1] - def_first):  # normal() aliasing side 1's copy: the element's wins
            m.normals[r[6]] = f0
        if def_first <= r[0] < n_elem_face:
            m.normals[r[0] - def_first] = f0
        if def_first <= r[1] < n_elem_face:
            m.normals[r[1] - def_first] = f1
    for row in np.asarray(m.ref_face).reshape(-1, 7):  # coarse hanging faces (:357-369)
        if def_first <= row[0] < n_elem_face:
            m.normals[row[0] - def_first] = face(row[0])
