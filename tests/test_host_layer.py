"""CPU: the product's host-side C++ (config, quadtree, solver_data, two-body) through the C ABI,
without a device context: bit-exact layouts against the reference's golden vectors and the oracle,
the reference's config error behaviour, and the exported symbol set of include/mara3_b200.h."""
import ctypes as C
import os
import re
import numpy as np
import pytest
import mara3_b200 as m3
from conftest import load_golden, ROOT
from oracle_util import OracleMesh, oracle_lib

MESH_FIELDS = ["vertices", "cell_centers", "cell_areas", "buffer_rate_field", "initial_conserved_u"]


def host_solver(cfg):
    return m3.Solver(cfg, host_only=True)


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "mara3_b200.h")).read()
    names = sorted(set(re.findall(r"\b(m3b_[a-z0-9_]+)\s*\(", header)))
    assert len(names) > 35
    lib = C.CDLL(m3.library_path())
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert m3.load_library().m3b_version().decode().startswith("mara3_b200")


def test_library_is_built_for_sm_100a():
    """The shipped .so carries sm_100a SASS (cuobjdump is in the image)."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", m3.library_path()], capture_output=True, text=True).stdout
    assert "sm_100a" in out


@pytest.mark.parametrize("name", ["nested_d3_n8", "uniform_d2_n16", "live_ecc_d3_n8", "rk1_axisym_d3_n8", "mesh_d5_n12", "mesh_default"])
def test_mesh_and_solver_data_match_reference_bit_for_bit(name):
    g = load_golden(name)
    s = host_solver(g["config"])
    assert np.array_equal(s.tree_index, g["tree_index"])
    for k in MESH_FIELDS:
        if k in g:
            assert np.array_equal(getattr(s, k), g[k]), k
    assert s.recommended_time_step == g["recommended_time_step"][0]
    assert s.gst_suppr_radius == g["gst_suppr_radius"][0]
    assert s.density_floor == g["density_floor"][0]


@pytest.mark.parametrize("cfg", [
    dict(), dict(depth=6, block_size=16), dict(depth=5, block_size=12), dict(depth=2, block_size=64),
    dict(depth=7, block_size=8, focus_factor=1.5, focus_index=1.2, domain_radius=7.5),
    dict(depth=4, block_size=10, focus_factor=1e3), dict(depth=1, block_size=6), dict(depth=0, block_size=8),
])
def test_mesh_matches_oracle_bit_for_bit(cfg):
    s, o = host_solver(cfg), OracleMesh(cfg)
    assert s.num_blocks == o.B and s.block_size == o.N
    assert np.array_equal(s.tree_index, o.tree_index)
    for k in MESH_FIELDS:
        assert np.array_equal(getattr(s, k), getattr(o, k)), k
    assert s.recommended_time_step == o.recommended_time_step
    assert s.gst_suppr_radius == o.gst_suppr_radius


def test_leaves_are_in_morton_order_and_tile_the_domain():
    s = host_solver(dict(depth=6, block_size=8))
    idx = s.tree_index
    top = int(idx[:, 0].max())

    def morton(level, i, j):
        x, y, key = int(i) << (top - level), int(j) << (top - level), 0
        for b in range(top):
            key |= ((x >> b) & 1) << (2 * b) | ((y >> b) & 1) << (2 * b + 1)
        return key
    keys = [morton(*r) for r in idx]
    assert keys == sorted(keys) and len(set(keys)) == len(keys)
    assert sum(4.0 ** -int(l) for l in idx[:, 0]) == 1.0                       # leaves tile the unit square
    v = s.vertices
    assert v.min() == -12.0 and v.max() == 12.0
    assert abs(s.cell_areas.sum() - 24.0 ** 2) < 1e-9


def test_config_behaves_like_the_reference():
    """app_config.hpp:103-136, 223-245: unknown key, duplicate key, wrong type; non key=value tokens ignored."""
    with pytest.raises(m3.Mara3Error, match="config has no option foo"):
        host_solver(dict(foo=1))
    with pytest.raises(m3.Mara3Error, match="duplicate parameter depth"):
        m3.Solver(argv=["depth=2", "depth=3"], host_only=True)
    with pytest.raises(m3.Mara3Error):
        host_solver(dict(depth="abc"))
    with pytest.raises(m3.Mara3Error, match="invalid reconstruct_method"):
        host_solver(dict(reconstruct_method="weno", depth=1, block_size=8))
    with pytest.raises(m3.Mara3Error, match="threaded"):
        host_solver(dict(threaded=0, depth=1, block_size=8))
    s = m3.Solver(argv=["binary", "--flag", "depth=1", "block_size=8"], host_only=True)
    assert s.num_blocks == 4
    defaults = {"cfl_number": "0.4", "depth": "1", "plm_theta": "1.8", "outdir": "data", "begin_live_binary": "1e+06",
                "domain_radius": "12", "alpha": "0.1", "rk_order": "2", "reconstruct_method": "plm"}
    for k, v in defaults.items():
        assert s.config(k) == v
    with pytest.raises(KeyError):
        s.config("nope")


def test_compute_calls_fail_loudly_without_a_device():
    s = host_solver(dict(depth=1, block_size=8))
    with pytest.raises(m3.Mara3Error, match="CUDA device"):
        s.create_solution()
    if m3.load_library().m3b_device_count() == 0:
        with pytest.raises(m3.Mara3Error, match="no CUDA device"):
            m3.Solver(dict(depth=1, block_size=8))


def test_two_body_model_matches_oracle():
    rng = np.random.default_rng(7)
    L = oracle_lib()
    dp = C.POINTER(C.c_double)
    for _ in range(50):
        e = np.array([rng.uniform(-3, 3), rng.uniform(-2, 2), *rng.uniform(-0.1, 0.1, 4), rng.uniform(0.5, 2.0),
                      rng.uniform(0.5, 2.0), rng.uniform(0.1, 1.0), rng.choice([0.0, rng.uniform(0.0, 0.8)])])
        t = rng.uniform(0.0, 20.0)
        want = np.zeros(10)
        L.m3o_two_body_state(e.ctypes.data_as(dp), t, want.ctypes.data_as(dp))
        got = m3.two_body_state(e, t)
        assert np.array_equal(got.reshape(-1), want)
        back_want = np.zeros(10)
        assert L.m3o_orbital_elements(want.ctypes.data_as(dp), t, back_want.ctypes.data_as(dp)) == 0
        assert np.array_equal(m3.orbital_elements(got, t), back_want)
    # reference vectors (src/physics_test.cpp:171-174): bodies start at (+-0.5, 0)
    b = m3.two_body_state([0, 0, 0, 0, 0, 0, 1, 1, 1, 0], 0.0)
    assert b[0, 1] == 0.5 and b[1, 1] == -0.5 and b[0, 2] == 0.0 and b[1, 2] == 0.0
    with pytest.raises(ValueError):
        fast = b.copy()
        fast[:, 3:] *= 10.0
        m3.orbital_elements(fast, 0.0)


def test_exchange_transport_is_none_on_one_rank():
    """m3b_exchange_transport: 0 / "none" without a device or with one rank (1 NCCL, 2 peer memory on several)."""
    assert m3.Solver(dict(depth=1, block_size=8), host_only=True).exchange_transport == "none"
