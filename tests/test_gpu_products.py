"""GPU: the HDF5 products and the `binary` subprogram itself.
Files are parsed with tests/h5_reader.py (an independent parser pinned against a libhdf5-written file)."""
import math
import os
import subprocess
import sys
import numpy as np
import pytest
import mara3_b200 as m3
from h5_reader import H5File

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CFG = dict(depth=4, block_size=16)          # the default nested tree: 64 leaves on levels 2-4
ELEMENTS = ("pomega", "tau", "cm_position_x", "cm_position_y", "cm_velocity_x", "cm_velocity_y", "elements")


def leaf_names(solver):
    idx = solver.tree_index
    return ["%d:%0*d-%0*d" % (l, int(1 + math.log10(1 << l)), i, int(1 + math.log10(1 << l)), j) for l, i, j in idx]


def test_checkpoint_layout_and_round_trip(tmp_path):
    s = m3.Solver(CFG)
    u = s.create_solution()
    for _ in range(3):
        s.next_solution(u)
    path = str(tmp_path / "chkpt.0000.h5")
    s.write_checkpoint(u, path)

    f = H5File(path)
    assert f.keys("/") == ["run_config", "schedule", "solution", "time_series"]
    assert f.read("/solution/time") == u.time and f.dtype("/solution/time") == np.dtype("<f8")
    assert list(f.read("/solution/iteration")) == [3, 1] and f.dtype("/solution/iteration") == np.dtype(("<i4", (2,)))

    names = leaf_names(s)
    assert sorted(names) == f.keys("/solution/conserved_u") and len(names) == s.num_blocks
    U = u.conserved_u
    N = s.block_size
    for b, name in enumerate(names):
        block = f.read("/solution/conserved_u/" + name)
        assert f.shape("/solution/conserved_u/" + name) == (N, N) and f.dtype("/solution/conserved_u/" + name) == np.dtype(("<f8", (3,)))
        # raw image of std::tuple<sigma, px, py> under libstdc++: (py, px, sigma)
        assert np.array_equal(block[..., 2], U[b, 0]) and np.array_equal(block[..., 1], U[b, 1]) and np.array_equal(block[..., 0], U[b, 2])
    assert f.keys("/solution/conserved_q") == ["0:0-0"] and f.shape("/solution/conserved_q/0:0-0") == (0, 0)

    scal = u.scalars
    assert np.array_equal(f.read("/solution/mass_accreted_on"), scal[3:5]) and f.dtype("/solution/mass_accreted_on") == np.dtype(("<f8", (2,)))
    el = f.read("/solution/orbital_elements")
    assert f.dtype("/solution/orbital_elements").names == ELEMENTS
    assert f.dtype("/solution/orbital_elements").fields["elements"][0].names == ("separation", "total_mass", "mass_ratio", "eccentricity")
    assert el["elements"]["separation"] == 1.0 and el["elements"]["mass_ratio"] == 1.0

    assert f.keys("/schedule") == ["record_time_series", "write_checkpoint", "write_diagnostics"]
    assert f.read("/schedule/write_checkpoint/name") == b"write_checkpoint" and f.read("/schedule/write_checkpoint/num_times_performed") == 0
    assert f.shape("/time_series") == (0,) and f.dtype("/time_series").itemsize == 376
    assert f.dtype("/time_series").names[:4] == ("time", "disk_mass", "disk_angular_momentum", "mass_accreted_on")
    assert f.dtype("/time_series").fields["position_of_mass2"][1] == 360
    assert len(f.keys("/run_config")) == 39
    assert f.read("/run_config/depth") == 4 and f.read("/run_config/cfl_number") == 0.4 and f.read("/run_config/outdir") == b"data"
    assert f.dtype("/run_config/restart") == np.dtype("S1")         # empty string: one NUL byte

    # restart: bit-identical state, and the following steps are bit-identical too
    v = s.create_solution()
    s.read_checkpoint(v, path)
    assert np.array_equal(v.conserved_u, U) and np.array_equal(v.scalars, scal) and v.time == u.time and v.iteration == (3, 1)
    for _ in range(2):
        s.next_solution(u)
        s.next_solution(v)
    assert np.array_equal(v.conserved_u, u.conserved_u) and np.array_equal(v.scalars, u.scalars)


def test_diagnostics_layout_and_values(tmp_path):
    s = m3.Solver(CFG)
    u = s.create_solution()
    s.next_solution(u)
    path = str(tmp_path / "diagnostics.0000.h5")
    s.write_diagnostics(u, path)
    f = H5File(path)
    assert f.keys("/") == ["phi_velocity", "position_of_mass1", "position_of_mass2", "radial_velocity", "run_config", "sigma", "time", "vertices"]
    names = leaf_names(s)
    N, U, X = s.block_size, u.conserved_u, s.cell_centers
    V = s.vertices
    for b in (0, len(names) // 2, len(names) - 1):
        verts = f.read("/vertices/" + names[b])
        assert verts.shape == (N + 1, N + 1, 2) and np.array_equal(verts[..., 0], V[b, 0]) and np.array_equal(verts[..., 1], V[b, 1])
        sigma, vx, vy = U[b, 0], U[b, 1] / U[b, 0], U[b, 2] / U[b, 0]
        r = np.sqrt(X[b, 0] ** 2 + X[b, 1] ** 2)
        assert np.array_equal(f.read("/sigma/" + names[b]), sigma)
        assert np.allclose(f.read("/radial_velocity/" + names[b]), vx * X[b, 0] / r + vy * X[b, 1] / r, rtol=1e-13, atol=1e-15)
        assert np.allclose(f.read("/phi_velocity/" + names[b]), -vx * X[b, 1] / r + vy * X[b, 0] / r, rtol=1e-13, atol=1e-15)
    bodies = m3.two_body_state(u.scalars[33:43], u.time)
    assert np.allclose(f.read("/position_of_mass1"), bodies[0, 1:3], rtol=0, atol=1e-15)

    sample = s.time_series_sample(u)
    dA = s.cell_areas
    assert sample[0] == u.time
    assert abs(sample[1] - float((U[:, 0] * dA).sum())) <= 1e-12 * abs(sample[1])                              # disk_mass
    lz = X[:, 0] * U[:, 2] - X[:, 1] * U[:, 1]
    assert abs(sample[2] - float((lz * dA).sum())) <= 1e-10 * abs(sample[2])                                   # disk_angular_momentum


def test_subprogram_run_loop_products_and_restart(tmp_path):
    exe = os.path.join(ROOT, "mara3_b200", "bin", "mara3b")
    out1, out2 = str(tmp_path / "run1"), str(tmp_path / "run2")
    common = ["binary", "depth=2", "block_size=32", "domain_radius=6", "cpi=0.01", "dfi=0.02", "tsi=0.004"]
    r = subprocess.run([exe] + common + ["tfinal=0.03", "outdir=" + out1], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    lines = r.stdout.splitlines()
    assert lines[0] == "=" * 52 and lines[1] == "config:"
    assert any(l.startswith("\tdepth...................") and l.endswith(" 2") for l in lines)
    assert f"write diagnostics: {out1}/diagnostics.0000.h5" in lines and f"write checkpoint: {out1}/chkpt.0000.h5" in lines
    steps = [l for l in lines if l.startswith("[")]
    assert steps[0].startswith("[0001] orbits=0.00") and "kzps=" in steps[0]
    assert lines[-1].startswith("total execution time: ")
    files = sorted(os.listdir(out1))
    assert "chkpt.0003.h5" in files and "diagnostics.0001.h5" in files and "chkpt.0004.h5" not in files

    last = H5File(os.path.join(out1, "chkpt.0003.h5"))
    t3 = last.read("/solution/time")
    assert 0.03 <= t3 / (2 * math.pi) < 0.036      # the first step past 3 x cpi
    assert last.read("/schedule/write_checkpoint/num_times_performed") == 4            # stored AFTER the task is marked complete
    series = last.read("/time_series")
    assert len(series) == last.read("/schedule/record_time_series/num_times_performed") >= 8
    assert np.all(np.diff(series["time"]) > 0) and series["time"][0] == 0.0             # oldest first
    assert abs(series["disk_mass"][0] - series["disk_mass"][-1]) < 1e-3 * series["disk_mass"][0]

    # restart from the second checkpoint with a different outdir: the run_config comes from the file, then the command line
    r2 = subprocess.run([exe, "binary", "restart=" + os.path.join(out1, "chkpt.0001.h5"), "outdir=" + out2, "tfinal=0.03"],
                        capture_output=True, text=True, timeout=600)
    assert r2.returncode == 0, r2.stdout[-2000:] + r2.stderr[-2000:]
    assert any(l.startswith("\tblock_size..............") and l.endswith(" 32") for l in r2.stdout.splitlines())
    again = H5File(os.path.join(out2, "chkpt.0003.h5"))
    assert again.read("/solution/time") == t3
    for name in last.keys("/solution/conserved_u"):
        assert np.array_equal(again.read("/solution/conserved_u/" + name), last.read("/solution/conserved_u/" + name))
    assert np.array_equal(again.read("/time_series")["disk_mass"], series["disk_mass"])

    bad = subprocess.run([exe, "binary", "no_such_key=1"], capture_output=True, text=True, timeout=120)
    assert bad.returncode == 1 and "config has no option no_such_key" in bad.stdout


def test_checkpoint_of_angular_momentum_variables(tmp_path):
    """conserve_linear_p=0: the state is written under /solution/conserved_q, conserved_u is the empty default tree."""
    s = m3.Solver(dict(depth=3, block_size=16, conserve_linear_p=0, fixed_dt=1))
    u = s.create_solution()
    s.next_solution(u)
    path = str(tmp_path / "chkpt.q.h5")
    s.write_checkpoint(u, path)
    f = H5File(path)
    assert f.keys("/solution/conserved_u") == ["0:0-0"] and f.shape("/solution/conserved_u/0:0-0") == (0, 0)
    names = leaf_names(s)
    assert sorted(names) == f.keys("/solution/conserved_q")
    Q = u.conserved_u                      # the API's state array holds conserved_q = (sigma, Sr, Lz) in this mode
    block = f.read("/solution/conserved_q/" + names[5])
    assert np.array_equal(block[..., 2], Q[5, 0]) and np.array_equal(block[..., 1], Q[5, 1]) and np.array_equal(block[..., 0], Q[5, 2])
    v = s.create_solution()
    s.read_checkpoint(v, path)
    assert np.array_equal(v.conserved_u, Q) and np.array_equal(v.scalars, u.scalars)
    sample = s.time_series_sample(u)
    assert abs(sample[2] - float((Q[:, 2] * s.cell_areas).sum())) <= 1e-12 * abs(sample[2])        # disk Lz = sum of the third component
