"""Development aid: one-step parity of the CUDA path against the plain-C oracle plus the stage kernel's
CUDA-event time on configs 2 and 3.  Usage: python tools/kernel_bench.py [tag]"""
import sys, os, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import mara3_b200 as m3
from oracle_util import OracleMesh, OracleSolution, SCALAR_NAMES


def block_rel_err(a, b):
    scale = np.abs(b).max(axis=(2, 3), keepdims=True)
    return float((np.abs(a - b) / scale).max())


def parity(cfg, steps=2):
    s = m3.Solver(cfg); om = OracleMesh(cfg)
    u = s.create_solution(); ou = OracleSolution(om)
    worst = 0.0
    for n in range(steps):
        dto, fbo = ou.next_solution()
        dtg, fbg = s.next_solution(u)
        den = np.maximum(np.abs(ou.scalars), 1e-300)
        se = np.abs(u.scalars - ou.scalars) / den; se[u.scalars == ou.scalars] = 0.0
        keep = [k for k, nm in enumerate(SCALAR_NAMES) if "work" not in nm and "pomega" not in nm and "tau" not in nm and "eccentricity" not in nm]
        worst = max(worst, block_rel_err(u.conserved_u, ou.conserved_u))
        print(f"  {cfg} step {n+1}: dt err {abs(dtg-dto)/dto:.1e} field err {block_rel_err(u.conserved_u, ou.conserved_u):.2e} scalars {se[keep].max():.1e}")
    return worst


def timing(cfg, n=40):
    s = m3.Solver(cfg); u = s.create_solution()
    s.run_steps(u, 5); s.synchronize()
    s.stage_timing(True)
    t0 = time.perf_counter(); s.run_steps(u, n); s.synchronize(); t1 = time.perf_counter()
    ms, k = s.stage_timing_read(); s.stage_timing(False)
    print(f"TIMING {cfg.get('depth')} cells {s.num_cells}: stage kernel {ms / k * 1e3:.1f} us ({s.num_cells / (ms / k) * 1e-6:.2f} Gcell-stage/s), step {(t1 - t0) / n * 1e3:.3f} ms, {s.num_cells * n / (t1 - t0) * 1e-9:.2f} Gzps")


if __name__ == "__main__":
    print("TAG", sys.argv[1] if len(sys.argv) > 1 else "")
    w = parity(dict(depth=2, block_size=64, domain_radius=6.0))
    w = max(w, parity(dict(depth=3, block_size=32, eccentricity=0.3, mass_ratio=0.5, density_floor=1e-2)))
    w = max(w, parity(dict(depth=4, block_size=64), steps=1))
    print("PARITY", "OK" if w < 1e-12 else "FAIL", w)
    timing(dict(depth=4, block_size=64, focus_factor=1e3, mach_number=10.0), 200)
    timing(dict(depth=6, block_size=64, focus_factor=1e3, mach_number=10.0), 40)
