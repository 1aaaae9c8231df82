"""Test-side helpers: ctypes binding of the plain-C oracle (oracle/libm3b_oracle.so),
a runner for the compiled reference (oracle/_ref/mara_ref) and a reader for the
"M3BD" dump files both write.  TEST INFRASTRUCTURE ONLY -- never imported by the
product package mara3_b200/."""
import ctypes as C
import os
import struct
import subprocess
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "libm3b_oracle.so")
REF_BIN = os.path.join(ORACLE_DIR, "_ref", "mara_ref")
REF_BIN_FAST = os.path.join(ORACLE_DIR, "_ref", "mara_ref_fast")

SCALAR_NAMES = (
    ["time", "iter_num", "iter_den"]
    + [f"mass_accreted_on[{k}]" for k in range(2)]
    + [f"angular_momentum_accreted_on[{k}]" for k in range(2)]
    + [f"integrated_torque_on[{k}]" for k in range(2)]
    + [f"work_done_on[{k}]" for k in range(2)]
    + ["mass_ejected", "angular_momentum_ejected"]
    + [f"orbital_elements_acc[{k}]" for k in range(10)]
    + [f"orbital_elements_grav[{k}]" for k in range(10)]
    + [f"orbital_elements[{k}]" for k in range(10)]
)


def read_dump(path):
    """Read an M3BD dump file -> dict name -> numpy array."""
    out = {}
    with open(path, "rb") as f:
        assert f.read(8) == b"M3BD0001", "not an M3BD dump"
        while True:
            head = f.read(4)
            if len(head) < 4:
                break
            (nl,) = struct.unpack("<I", head)
            name = f.read(nl).decode()
            (dtype,) = struct.unpack("<B", f.read(1))
            (nd,) = struct.unpack("<I", f.read(4))
            dims = struct.unpack("<%dQ" % nd, f.read(8 * nd))
            count = int(np.prod(dims)) if nd else 1
            data = np.frombuffer(f.read(8 * count), dtype=np.float64 if dtype == 0 else np.int64)
            out[name] = data.reshape(dims).copy()
    return out


def dump_scalars(d, prefix):
    """The 43-double scalar vector (oracle layout) from a reference dump."""
    g = lambda k: np.atleast_1d(d[prefix + k]).astype(np.float64)
    return np.concatenate([
        g("time"), g("iteration"), g("mass_accreted_on"), g("angular_momentum_accreted_on"),
        g("integrated_torque_on"), g("work_done_on"), g("mass_ejected"), g("angular_momentum_ejected"),
        g("orbital_elements_acc"), g("orbital_elements_grav"), g("orbital_elements")])


def have_reference():
    return os.path.exists(REF_BIN) and os.access(REF_BIN, os.X_OK)


def run_reference(config, steps=0, dump=None, dump_steps=(), stages=False, mesh_only=False, timing=False,
                  warmup=0, fast=False, timeout=3600):
    """Run the compiled reference (oracle/_ref).  config: dict key -> value."""
    cmd = [REF_BIN_FAST if fast else REF_BIN, "--steps", str(steps)]
    if warmup:
        cmd += ["--warmup", str(warmup)]
    if dump:
        cmd += ["--dump", dump]
    if dump_steps:
        cmd += ["--dump-steps", ",".join(str(s) for s in dump_steps)]
    if stages:
        cmd += ["--stages"]
    if mesh_only:
        cmd += ["--mesh-only"]
    if timing:
        cmd += ["--timing"]
    cmd += [f"{k}={v}" for k, v in config.items()]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)
    if r.returncode != 0:
        raise RuntimeError("reference failed: %s\n%s" % (" ".join(cmd), r.stderr))
    return r.stdout


class Config(C.Structure):
    _fields_ = [
        ("restart", C.c_char * 512), ("outdir", C.c_char * 512),
        ("cpi", C.c_double), ("dfi", C.c_double), ("tsi", C.c_double), ("tfinal", C.c_double), ("cfl_number", C.c_double),
        ("fixed_dt", C.c_int), ("depth", C.c_int), ("begin_live_binary", C.c_double),
        ("conserve_linear_p", C.c_int), ("block_size", C.c_int),
        ("focus_factor", C.c_double), ("focus_index", C.c_double),
        ("threaded", C.c_int), ("rk_order", C.c_int), ("reconstruct_method", C.c_char * 16),
        ("plm_theta", C.c_double), ("source_term_softening", C.c_double), ("softening_radius", C.c_double),
        ("sink_radius", C.c_double), ("sink_rate", C.c_double), ("buffer_damping_rate", C.c_double),
        ("domain_radius", C.c_double), ("disk_radius", C.c_double), ("disk_mass", C.c_double),
        ("ambient_density", C.c_double), ("density_floor", C.c_double), ("separation", C.c_double),
        ("mass_ratio", C.c_double), ("eccentricity", C.c_double), ("counter_rotate", C.c_int),
        ("mach_number", C.c_double), ("axisymmetric_cs2", C.c_int), ("no_accretion_force", C.c_int),
        ("alpha_cutoff_radius", C.c_double), ("alpha", C.c_double), ("nu", C.c_double), ("mdot", C.c_double),
    ]


_lib = None


def oracle_lib():
    global _lib
    if _lib is None:
        if not os.path.exists(ORACLE_SO):
            subprocess.check_call(["make", "-C", ORACLE_DIR, "port"])
        L = C.CDLL(ORACLE_SO)
        vp, dp, ip = C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int64)
        L.m3o_config_default.argtypes = [C.POINTER(Config)]
        L.m3o_config_set.argtypes = [C.POINTER(Config), C.c_char_p, C.c_char_p]
        L.m3o_mesh_create.restype = vp
        L.m3o_mesh_create.argtypes = [C.POINTER(Config)]
        L.m3o_mesh_destroy.argtypes = [vp]
        for name in ("num_blocks", "block_size"):
            getattr(L, "m3o_mesh_" + name).argtypes = [vp]
        L.m3o_mesh_tree_index.argtypes = [vp, ip]
        for name in ("vertices", "cell_centers", "cell_areas", "buffer_rate", "initial_conserved_u"):
            getattr(L, "m3o_mesh_" + name).argtypes = [vp, dp]
        for name in ("recommended_time_step", "gst_suppr_radius", "density_floor"):
            fn = getattr(L, "m3o_mesh_" + name)
            fn.argtypes = [vp]
            fn.restype = C.c_double
        L.m3o_solution_create.restype = vp
        L.m3o_solution_create.argtypes = [vp]
        L.m3o_solution_clone.restype = vp
        L.m3o_solution_clone.argtypes = [vp]
        L.m3o_solution_destroy.argtypes = [vp]
        for name in ("get_conserved", "get_scalars"):
            getattr(L, "m3o_solution_" + name).argtypes = [vp, dp]
        for name in ("set_conserved", "set_scalars"):
            getattr(L, "m3o_solution_" + name).argtypes = [vp, dp]
        L.m3o_advance.argtypes = [vp, vp, C.c_double, C.c_int, vp]
        L.m3o_maximum_timestep.argtypes = [vp, vp]
        L.m3o_maximum_timestep.restype = C.c_double
        L.m3o_next_solution.argtypes = [vp, vp, dp]
        L.m3o_two_body_state.argtypes = [dp, C.c_double, dp]
        L.m3o_orbital_elements.argtypes = [dp, C.c_double, dp]
        L.m3o_plm_gradient.argtypes = [C.c_double] * 4
        L.m3o_plm_gradient.restype = C.c_double
        L.m3o_riemann_hlle.argtypes = [dp, dp, C.c_double, C.c_int, dp]
        _lib = L
    return _lib


def _dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class OracleMesh:
    """create_solver_data of the plain-C oracle."""

    def __init__(self, config):
        L = oracle_lib()
        self.cfg = Config()
        L.m3o_config_default(C.byref(self.cfg))
        for k, v in config.items():
            rc = L.m3o_config_set(C.byref(self.cfg), k.encode(), str(v).encode())
            if rc:
                raise ValueError("config has no option " + k if rc == 1 else "bad value for " + k)
        self.h = L.m3o_mesh_create(C.byref(self.cfg))
        self.B = L.m3o_mesh_num_blocks(self.h)
        self.N = L.m3o_mesh_block_size(self.h)

    def __del__(self):
        if getattr(self, "h", None):
            oracle_lib().m3o_mesh_destroy(self.h)
            self.h = None

    def _get(self, name, shape):
        a = np.empty(shape, dtype=np.float64)
        getattr(oracle_lib(), "m3o_mesh_" + name)(self.h, _dptr(a))
        return a

    @property
    def tree_index(self):
        a = np.empty((self.B, 3), dtype=np.int64)
        oracle_lib().m3o_mesh_tree_index(self.h, a.ctypes.data_as(C.POINTER(C.c_int64)))
        return a

    vertices = property(lambda s: s._get("vertices", (s.B, 2, s.N + 1, s.N + 1)))
    cell_centers = property(lambda s: s._get("cell_centers", (s.B, 2, s.N, s.N)))
    cell_areas = property(lambda s: s._get("cell_areas", (s.B, s.N, s.N)))
    buffer_rate_field = property(lambda s: s._get("buffer_rate", (s.B, s.N, s.N)))
    initial_conserved_u = property(lambda s: s._get("initial_conserved_u", (s.B, 3, s.N, s.N)))
    recommended_time_step = property(lambda s: oracle_lib().m3o_mesh_recommended_time_step(s.h))
    gst_suppr_radius = property(lambda s: oracle_lib().m3o_mesh_gst_suppr_radius(s.h))
    density_floor = property(lambda s: oracle_lib().m3o_mesh_density_floor(s.h))


class OracleSolution:
    """solution_t of the plain-C oracle."""

    def __init__(self, mesh, handle=None):
        self.mesh = mesh
        self.h = handle if handle is not None else oracle_lib().m3o_solution_create(mesh.h)

    def __del__(self):
        if getattr(self, "h", None):
            oracle_lib().m3o_solution_destroy(self.h)
            self.h = None

    def clone(self):
        return OracleSolution(self.mesh, oracle_lib().m3o_solution_clone(self.h))

    @property
    def conserved_u(self):
        m = self.mesh
        a = np.empty((m.B, 3, m.N, m.N), dtype=np.float64)
        oracle_lib().m3o_solution_get_conserved(self.h, _dptr(a))
        return a

    @conserved_u.setter
    def conserved_u(self, a):
        a = np.ascontiguousarray(a, dtype=np.float64)
        oracle_lib().m3o_solution_set_conserved(self.h, _dptr(a))

    @property
    def scalars(self):
        a = np.empty(43, dtype=np.float64)
        oracle_lib().m3o_solution_get_scalars(self.h, _dptr(a))
        return a

    @scalars.setter
    def scalars(self, a):
        a = np.ascontiguousarray(a, dtype=np.float64)
        oracle_lib().m3o_solution_set_scalars(self.h, _dptr(a))

    @property
    def time(self):
        return self.scalars[0]

    def advance(self, dt, safe_mode=False):
        """binary::advance -> (new solution, status)."""
        out = self.clone()
        status = oracle_lib().m3o_advance(self.mesh.h, self.h, float(dt), int(safe_mode), out.h)
        return out, status

    def maximum_timestep(self):
        return oracle_lib().m3o_maximum_timestep(self.mesh.h, self.h)

    def next_solution(self):
        """binary::next_solution, in place -> (dt used, fell_back)."""
        dt = C.c_double(0.0)
        fb = oracle_lib().m3o_next_solution(self.mesh.h, self.h, C.byref(dt))
        return dt.value, bool(fb)
