"""mara3_b200 -- B200-native implementation of Mara3's `binary` isothermal-2D hot path.

This package is a thin ctypes view of the C ABI in ``include/mara3_b200.h``
(``mara3_b200/libmara3_b200.so``, built from ``mara3_b200/csrc`` for sm_100a).
The names mirror the reference's operator API for the path
(Mara3 ``src/subprog_binary.hpp:180-208``):

    solver   = mara3_b200.create_solver_data(depth=4, block_size=64, ...)   # create_run_config + create_solver_data
    solution = solver.create_solution()                                     # binary::create_solution
    dt       = solver.maximum_timestep(solution)                            # binary::maximum_timestep
    s1       = solver.advance(solution, dt)                                 # binary::advance
    solver.next_solution(solution)                                          # binary::next_solution (in place)

There is no CPU compute path: without the compiled CUDA library, or without a
GPU, every compute call raises.
"""
import ctypes as C
import os
import numpy as np

__all__ = ["create_solver_data", "Solver", "Solution", "NegativeDensity", "Mara3Error", "library_path", "load_library"]

NUM_SCALARS = 43
OK, NEGATIVE_DENSITY, UNBOUND_ORBIT, UNSUPPORTED = 0, 1, 2, 3
FLAG_GENERAL_ONLY = 1   # every block through the any-tree kernels (tests)
FLAG_HOST_ONLY = 2      # mesh / solver_data queries only; no device context
FLAG_TILED_KERNEL = 4   # stage_fused instead of stage_strip (tests)

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


class Mara3Error(RuntimeError):
    pass


class NegativeDensity(Mara3Error):
    """The reference's std::runtime_error("negative density in updated state")."""

    def __init__(self, message, lines):
        super().__init__(message)
        self.lines = lines


def library_path():
    """The in-tree library; M3B_LIBRARY names another build of it (development: two builds compared in one visit to the GPU box)."""
    return os.environ.get("M3B_LIBRARY") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libmara3_b200.so")


_lib = None


def load_library():
    """Load the C-ABI library; fail loudly if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise Mara3Error(
            f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C mara3_b200/csrc`; mara3_b200 has no CPU fallback")
    L = C.CDLL(path)
    vp, dp, ip = C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int)
    L.m3b_version.restype = C.c_char_p
    L.m3b_global_error.restype = C.c_char_p
    L.m3b_solver_create.restype = vp
    L.m3b_solver_create.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.c_int, C.c_int]
    L.m3b_solver_create_distributed.restype = vp
    L.m3b_solver_create_distributed.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.c_int, C.c_int, C.c_int, C.c_int, C.c_char_p]
    L.m3b_nccl_unique_id.argtypes = [C.c_char_p]
    for name in ("num_global_blocks", "first_block", "num_local_blocks"):
        getattr(L, "m3b_" + name).argtypes = [vp]
    L.m3b_num_owned_cells.argtypes = [vp]
    L.m3b_exchange_timing.argtypes = [vp, C.POINTER(C.c_double)]
    L.m3b_num_owned_cells.restype = C.c_int64
    L.m3b_local_to_global.argtypes = [vp, ip]
    L.m3b_halo_plan_size.argtypes = [vp, C.c_int, C.c_int]
    L.m3b_halo_plan.argtypes = [vp, C.c_int, C.c_int, ip]
    L.m3b_neighbor_table.argtypes = [vp, ip]
    L.m3b_face_neighbor_table.argtypes = [vp, ip]
    for name in ("m3b_write_checkpoint", "m3b_write_diagnostics", "m3b_read_checkpoint"):
        getattr(L, name).argtypes = [vp, vp, C.c_char_p]
    L.m3b_time_series_sample.argtypes = [vp, vp, dp]
    L.m3b_binary_main.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.c_int]
    L.m3b_binary_main_distributed.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.c_int, C.c_int, C.c_int, C.c_char_p]
    L.m3b_exchange_transport.argtypes = [vp]
    L.m3b_exchange_transport.restype = C.c_int
    L.m3b_halo_bytes_per_exchange.argtypes = [vp]
    L.m3b_halo_bytes_per_exchange.restype = C.c_uint64
    L.m3b_solver_destroy.argtypes = [vp]
    L.m3b_last_error.restype = C.c_char_p
    L.m3b_last_error.argtypes = [vp]
    for name in ("num_blocks", "block_size", "num_regular_blocks", "num_messages"):
        getattr(L, "m3b_" + name).argtypes = [vp]
    L.m3b_num_cells.argtypes = [vp]
    L.m3b_num_cells.restype = C.c_int64
    L.m3b_tree_index.argtypes = [vp, C.POINTER(C.c_int64)]
    for name in ("vertices", "cell_centers", "cell_areas", "buffer_rate_field", "initial_conserved_u"):
        getattr(L, "m3b_" + name).argtypes = [vp, dp]
    for name in ("recommended_time_step", "gst_suppr_radius", "density_floor"):
        fn = getattr(L, "m3b_" + name)
        fn.argtypes = [vp]
        fn.restype = C.c_double
    L.m3b_config_get.argtypes = [vp, C.c_char_p, C.c_char_p, C.c_int]
    L.m3b_solution_create.restype = vp
    L.m3b_solution_create.argtypes = [vp]
    L.m3b_solution_clone.restype = vp
    L.m3b_solution_clone.argtypes = [vp, vp]
    L.m3b_solution_destroy.argtypes = [vp]
    L.m3b_solution_set_conserved.argtypes = [vp, vp, dp]
    L.m3b_solution_get_conserved.argtypes = [vp, vp, dp]
    L.m3b_solution_set_scalars.argtypes = [vp, dp]
    L.m3b_solution_get_scalars.argtypes = [vp, dp]
    L.m3b_maximum_timestep.argtypes = [vp, vp, dp]
    L.m3b_advance.argtypes = [vp, vp, C.c_double, C.c_int, vp]
    L.m3b_solution_combine.argtypes = [vp, vp, vp, C.c_double, vp]
    L.m3b_next_solution.argtypes = [vp, vp, dp, ip]
    L.m3b_run_steps.argtypes = [vp, vp, C.c_int, ip]
    L.m3b_advance_host.argtypes = [vp, dp, dp, C.c_double, C.c_int, dp, dp]
    L.m3b_next_solution_host.argtypes = [vp, dp, dp, dp, dp, dp, ip]
    L.m3b_message.argtypes = [vp, C.c_int]
    L.m3b_message.restype = C.c_char_p
    L.m3b_set_quiet.argtypes = [vp, C.c_int]
    L.m3b_set_pipelining.argtypes = [vp, C.c_int]
    L.m3b_kernel_launches.argtypes = [vp]
    L.m3b_kernel_launches.restype = C.c_uint64
    L.m3b_stage_timing.argtypes = [vp, C.c_int]
    L.m3b_stage_timing_read.argtypes = [vp, dp, C.POINTER(C.c_uint64)]
    L.m3b_synchronize.argtypes = [vp]
    L.m3b_two_body_state.argtypes = [dp, C.c_double, dp]
    L.m3b_orbital_elements.argtypes = [dp, C.c_double, dp]
    L.m3b_set_stream.argtypes = [vp, vp]
    _lib = L
    return L


def _dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _host_array(a, shape):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if a.shape != tuple(shape):
        raise ValueError(f"expected an array of shape {tuple(shape)}, got {a.shape}")
    return a


class Solution:
    """binary::solution_t -- conserved field resident on the GPU plus host scalars."""

    def __init__(self, solver, handle):
        if not handle:
            raise Mara3Error(solver._error())
        self.solver = solver
        self._h = handle

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.m3b_solution_destroy(self._h)
            self._h = None

    def clone(self):
        return Solution(self.solver, _lib.m3b_solution_clone(self.solver._h, self._h))

    @property
    def conserved_u(self):
        """Download as [B][3][N][N] (sigma, px, py)."""
        s = self.solver
        a = np.empty((s.num_blocks, 3, s.block_size, s.block_size), dtype=np.float64)
        s._check(_lib.m3b_solution_get_conserved(s._h, self._h, _dptr(a)))
        return a

    @conserved_u.setter
    def conserved_u(self, a):
        s = self.solver
        a = _host_array(a, (s.num_blocks, 3, s.block_size, s.block_size))
        s._check(_lib.m3b_solution_set_conserved(s._h, self._h, _dptr(a)))

    @property
    def scalars(self):
        a = np.empty(NUM_SCALARS, dtype=np.float64)
        _lib.m3b_solution_get_scalars(self._h, _dptr(a))
        return a

    @scalars.setter
    def scalars(self, a):
        a = _host_array(a, (NUM_SCALARS,))
        _lib.m3b_solution_set_scalars(self._h, _dptr(a))

    @property
    def time(self):
        return float(self.scalars[0])

    @property
    def iteration(self):
        s = self.scalars
        return int(s[1]), int(s[2])


class Solver:
    """run_config + solver_data_t + the device context (one GPU)."""

    def __init__(self, config=None, device=0, general_only=False, host_only=False, tiled_kernel=False, quiet=True, argv=None,
                 rank=0, nranks=1, nccl_unique_id=None, **keys):
        """rank / nranks / nccl_unique_id: one process per GPU (see nccl_unique_id()); per-block arrays and
        conserved_u then cover the blocks this rank owns, global ids first_block .. first_block + num_blocks."""
        L = load_library()
        items = dict(config or {})
        items.update(keys)
        tokens = list(argv or []) + [f"{k}={_format(v)}" for k, v in items.items()]
        arr = (C.c_char_p * max(1, len(tokens)))(*[t.encode() for t in tokens])
        flags = (FLAG_GENERAL_ONLY if general_only else 0) | (FLAG_HOST_ONLY if host_only else 0) | (FLAG_TILED_KERNEL if tiled_kernel else 0)
        uid = bytes(nccl_unique_id) if nccl_unique_id is not None else None
        self._h = L.m3b_solver_create_distributed(len(tokens), arr, int(device), flags, int(rank), int(nranks), uid)
        if not self._h:
            raise Mara3Error(L.m3b_global_error().decode())
        self.host_only = host_only
        self.rank, self.nranks = int(rank), int(nranks)
        self.num_blocks = L.m3b_num_blocks(self._h)
        self.num_global_blocks = L.m3b_num_global_blocks(self._h)
        self.first_block = L.m3b_first_block(self._h)
        self.num_local_blocks = L.m3b_num_local_blocks(self._h)
        self.num_owned_cells = L.m3b_num_owned_cells(self._h)
        self.block_size = L.m3b_block_size(self._h)
        self.num_cells = L.m3b_num_cells(self._h)
        L.m3b_set_quiet(self._h, int(quiet))

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.m3b_solver_destroy(self._h)
            self._h = None

    # ---- errors -------------------------------------------------------------------------------
    def _error(self):
        return _lib.m3b_last_error(self._h).decode()

    def messages(self):
        return [_lib.m3b_message(self._h, n).decode() for n in range(_lib.m3b_num_messages(self._h))]

    def _check(self, status):
        if status == OK:
            return
        if status == NEGATIVE_DENSITY:
            raise NegativeDensity("negative density in updated state", self.messages())
        raise Mara3Error(self._error() or f"status {status}")

    # ---- solver_data --------------------------------------------------------------------------
    def _get(self, name, shape, dtype=np.float64):
        a = np.empty(shape, dtype=dtype)
        ptr = a.ctypes.data_as(C.POINTER(C.c_int64 if dtype == np.int64 else C.c_double))
        getattr(_lib, "m3b_" + name)(self._h, ptr)
        return a

    tree_index = property(lambda s: s._get("tree_index", (s.num_blocks, 3), np.int64))
    vertices = property(lambda s: s._get("vertices", (s.num_blocks, 2, s.block_size + 1, s.block_size + 1)))
    cell_centers = property(lambda s: s._get("cell_centers", (s.num_blocks, 2, s.block_size, s.block_size)))
    cell_areas = property(lambda s: s._get("cell_areas", (s.num_blocks, s.block_size, s.block_size)))
    buffer_rate_field = property(lambda s: s._get("buffer_rate_field", (s.num_blocks, s.block_size, s.block_size)))
    initial_conserved_u = property(lambda s: s._get("initial_conserved_u", (s.num_blocks, 3, s.block_size, s.block_size)))
    recommended_time_step = property(lambda s: _lib.m3b_recommended_time_step(s._h))
    gst_suppr_radius = property(lambda s: _lib.m3b_gst_suppr_radius(s._h))
    density_floor = property(lambda s: _lib.m3b_density_floor(s._h))
    num_regular_blocks = property(lambda s: _lib.m3b_num_regular_blocks(s._h))
    kernel_launches = property(lambda s: int(_lib.m3b_kernel_launches(s._h)))

    exchange_transport = property(lambda s: ("none", "nccl", "peer")[int(_lib.m3b_exchange_transport(s._h))])
    halo_bytes_per_exchange = property(lambda s: int(_lib.m3b_halo_bytes_per_exchange(s._h)))

    @property
    def local_to_global(self):
        a = np.empty(self.num_local_blocks, dtype=np.int32)
        _lib.m3b_local_to_global(self._h, a.ctypes.data_as(C.POINTER(C.c_int)))
        return a

    @property
    def neighbor_table(self):
        """[owned][3][3] local ids of the same-level neighbours (-1: none), the table the stage kernel reads its halo through."""
        a = np.empty((self.num_blocks, 3, 3), dtype=np.int32)
        _lib.m3b_neighbor_table(self._h, a.ctypes.data_as(C.POINTER(C.c_int)))
        return a

    @property
    def face_neighbor_table(self):
        """[local block][side -x, +x, -y, +y][kind, 4 leaf ids] in local numbering (-1: none or not stored on this rank)."""
        a = np.empty((self.num_local_blocks, 4, 5), dtype=np.int32)
        _lib.m3b_face_neighbor_table(self._h, a.ctypes.data_as(C.POINTER(C.c_int)))
        return a

    def halo_plan(self, peer, send):
        """Ordered (local block, di, dj) strips sent to / received from `peer` before each stage."""
        n = _lib.m3b_halo_plan_size(self._h, int(peer), int(send))
        a = np.empty((n, 3), dtype=np.int32)
        if n:
            _lib.m3b_halo_plan(self._h, int(peer), int(send), a.ctypes.data_as(C.POINTER(C.c_int)))
        return a

    def config(self, key):
        buf = C.create_string_buffer(1024)
        if _lib.m3b_config_get(self._h, key.encode(), buf, 1024):
            raise KeyError("config has no option " + key)
        return buf.value.decode()

    # ---- the operator API ---------------------------------------------------------------------
    def create_solution(self):
        return Solution(self, _lib.m3b_solution_create(self._h))

    def maximum_timestep(self, solution):
        dt = C.c_double(0.0)
        self._check(_lib.m3b_maximum_timestep(self._h, solution._h, C.byref(dt)))
        return dt.value

    def advance(self, solution, dt, safe_mode=False):
        """binary::advance: returns the new solution; raises NegativeDensity like the reference throws."""
        out = solution.clone()
        status = _lib.m3b_advance(self._h, solution._h, float(dt), int(safe_mode), out._h)
        if status == NEGATIVE_DENSITY:
            err = NegativeDensity("negative density in updated state", self.messages())
            err.solution = out
            raise err
        self._check(status)
        return out

    def combine(self, a, b, b0):
        """a * b0 + b * (1 - b0)."""
        out = a.clone()
        self._check(_lib.m3b_solution_combine(self._h, a._h, b._h, float(b0), out._h))
        return out

    def next_solution(self, solution):
        """binary::next_solution in place -> (dt used, fell back to safe mode)."""
        dt, fb = C.c_double(0.0), C.c_int(0)
        self._check(_lib.m3b_next_solution(self._h, solution._h, C.byref(dt), C.byref(fb)))
        return dt.value, bool(fb.value)

    def run_steps(self, solution, count):
        fb = C.c_int(0)
        self._check(_lib.m3b_run_steps(self._h, solution._h, int(count), C.byref(fb)))
        return fb.value

    def advance_host(self, u, scalars, dt, safe_mode=False, out=None):
        """binary::advance on host arrays (H2D + stage + D2H inside the call)."""
        shape = (self.num_blocks, 3, self.block_size, self.block_size)
        u = _host_array(u, shape)
        scalars = _host_array(scalars, (NUM_SCALARS,))
        u_out = np.empty(shape, dtype=np.float64) if out is None else out
        s_out = np.empty(NUM_SCALARS, dtype=np.float64)
        self._check(_lib.m3b_advance_host(self._h, _dptr(u), _dptr(scalars), float(dt), int(safe_mode), _dptr(u_out), _dptr(s_out)))
        return u_out, s_out

    def next_solution_host(self, u, scalars, out=None):
        """binary::next_solution on host arrays -> (u, scalars, dt, fell_back)."""
        shape = (self.num_blocks, 3, self.block_size, self.block_size)
        u = _host_array(u, shape)
        scalars = _host_array(scalars, (NUM_SCALARS,))
        u_out = np.empty(shape, dtype=np.float64) if out is None else out
        s_out = np.empty(NUM_SCALARS, dtype=np.float64)
        dt, fb = C.c_double(0.0), C.c_int(0)
        self._check(_lib.m3b_next_solution_host(self._h, _dptr(u), _dptr(scalars), _dptr(u_out), _dptr(s_out), C.byref(dt), C.byref(fb)))
        return u_out, s_out, dt.value, bool(fb.value)

    # ---- HDF5 products --------------------------------------------------------
    def write_checkpoint(self, solution, filename):
        """chkpt.NNNN.h5 layout (subprog_binary_io.cpp:131-158) for `solution`, initial schedule, empty time series."""
        self._check(_lib.m3b_write_checkpoint(self._h, solution._h, os.fsencode(filename)))

    def write_diagnostics(self, solution, filename):
        """diagnostics.NNNN.h5 layout (subprog_binary_io.cpp:160-172): vertices, sigma, radial and azimuthal velocity per leaf."""
        self._check(_lib.m3b_write_diagnostics(self._h, solution._h, os.fsencode(filename)))

    def read_checkpoint(self, solution, filename):
        """Load /solution of a checkpoint (written by this library or by the reference with libhdf5's defaults) into `solution`."""
        self._check(_lib.m3b_read_checkpoint(self._h, solution._h, os.fsencode(filename)))

    def time_series_sample(self, solution):
        """The 47 doubles of time_series_sample_t (subprog_binary.hpp:144-161) for `solution`."""
        a = np.empty(47, dtype=np.float64)
        self._check(_lib.m3b_time_series_sample(self._h, solution._h, _dptr(a)))
        return a

    # ---- measurement helpers ------------------------------------------------------------------
    def stage_timing(self, enable):
        _lib.m3b_stage_timing(self._h, int(enable))

    def stage_timing_read(self):
        ms, n = C.c_double(0.0), C.c_uint64(0)
        self._check(_lib.m3b_stage_timing_read(self._h, C.byref(ms), C.byref(n)))
        return ms.value, int(n.value)

    def exchange_timing(self):
        """Device-clock instrumentation of the guard-zone / result exchange since the last call (multi-GPU): microseconds per
        step that boundary tiles waited for ghost cells, that the result exchange waited for the slowest rank, and the achieved
        NVLink rate of the guard-zone pushes.  Empty on one rank."""
        a = np.zeros(8, dtype=np.float64)
        if not hasattr(_lib, "m3b_exchange_timing") or _lib.m3b_exchange_timing(self._h, _dptr(a)) != 0:
            return {}
        steps = max(1.0, a[0])
        out = {"steps_instrumented": int(a[0]), "result_exchange_us": a[2] / steps,
               "clock": "%globaltimer inside the kernels (result exchange, fused guard-zone exchange) and CUDA events (split launch), rank 0"}
        if a[7] > 0:
            # fused exchange: every CTA of the stage kernel passes exchange_unpack once per stage (flag wait + its share of the unpack)
            out["unpack_us_per_cta_visit"] = a[6] / a[7]
            out["exposed_wait_us_per_step"] = 2.0 * a[6] / a[7]
            out["how"] = "guard zones pushed and unpacked by the stage kernel itself (stage_tma: exchange_push / exchange_unpack)"
            if a[4] > 0 and a[3] > 0:
                # kernel start -> every strip stored in the neighbours' landing buffers and the flags raised there
                out["push_us_per_exchange"] = a[3] / a[4]
                out["nvlink_gbs_achieved"] = a[5] / (a[3] * 1e-6) * 1e-9
                out["nvlink_peak_gbs"] = 900.0
                out["nvlink_note"] = ("bytes this rank stores into its neighbours' memory per exchange / push time; the messages are a few "
                                      "hundred KB, so the figure is set by store + fence latency, not by the 900 GB/s per direction of NVLink 5")
        elif a[4] > 0:
            out["exposed_wait_us_per_step"] = a[1] / steps
            out["push_to_unpacked_us_per_exchange"] = a[3] / a[4]
            out["nvlink_gbs_achieved"] = a[5] / (a[3] * 1e-6) * 1e-9 if a[3] > 0 else None
            out["how"] = "halo_push / halo_wait_unpack on the exchange stream beside the interior launch, boundary launch behind them"
        return out

    def set_stream(self, cuda_stream):
        """Launch on a caller-owned CUDA stream (an integer cudaStream_t, e.g. torch.cuda.current_stream().cuda_stream)."""
        self._check(_lib.m3b_set_stream(self._h, C.c_void_p(cuda_stream)))

    def set_pipelining(self, on):
        """Queue the following step before waiting for the current one (default on)."""
        _lib.m3b_set_pipelining(self._h, int(on))

    def synchronize(self):
        _lib.m3b_synchronize(self._h)


def _format(v):
    if isinstance(v, bool):
        return str(int(v))
    if isinstance(v, float):
        return repr(v)
    return str(v)


def create_solver_data(config=None, **kwargs):
    """create_run_config + create_solver_data + set_scheme_globals."""
    return Solver(config, **kwargs)


def two_body_state(elements, t):
    """mara::compute_two_body_state: elements (10 doubles) -> [[mass, x, y, vx, vy], [..]]."""
    L = load_library()
    e = _host_array(elements, (10,))
    out = np.empty((2, 5), dtype=np.float64)
    L.m3b_two_body_state(_dptr(e), float(t), _dptr(out))
    return out


def orbital_elements(bodies, t):
    """mara::compute_orbital_elements; raises ValueError for an unbound pair (the reference throws)."""
    L = load_library()
    b = _host_array(bodies, (2, 5))
    out = np.empty(10, dtype=np.float64)
    if L.m3b_orbital_elements(_dptr(b), float(t), _dptr(out)):
        raise ValueError("mara::compute_orbital_elements (two_body_state does not correspond to a bound orbit)")
    return out


def binary_main(argv, device=0, rank=0, nranks=1, nccl_unique_id=None):
    """The `binary` subprogram (subprog_binary.cpp:414-436): argv = ["binary", "key=value", ...]; returns the exit code.
    With nranks > 1 every rank (one process per GPU) calls this; rank 0 prints and writes the gathered products."""
    load_library()
    args = (C.c_char_p * len(argv))(*[os.fsencode(a) for a in argv])
    if nranks == 1:
        return _lib.m3b_binary_main(len(argv), args, int(device))
    return _lib.m3b_binary_main_distributed(len(argv), args, int(device), int(rank), int(nranks), bytes(nccl_unique_id))


def nccl_unique_id():
    """A fresh 128-byte NCCL unique id (rank 0 creates it and sends it to the other ranks)."""
    L = load_library()
    buf = C.create_string_buffer(128)
    if L.m3b_nccl_unique_id(buf):
        raise Mara3Error(L.m3b_global_error().decode())
    return buf.raw
