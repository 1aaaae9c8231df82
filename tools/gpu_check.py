"""Quick GPU parity report (development aid): CUDA path vs the plain-C oracle."""
import sys, os, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import mara3_b200 as m3
from oracle_util import OracleMesh, OracleSolution, SCALAR_NAMES


def block_rel_err(a, b):
    """max over cells of |a - b| / (max |b| over the cell's block and field)"""
    scale = np.abs(b).max(axis=(2, 3), keepdims=True)
    return float((np.abs(a - b) / scale).max())


def scalar_rel_err(a, b):
    den = np.maximum(np.abs(b), 1e-300)
    e = np.abs(a - b) / den
    e[(a == b)] = 0.0
    return e


def check(cfg, general_only=False, steps=3):
    tag = f"{cfg} general_only={general_only}"
    s = m3.Solver(cfg, general_only=general_only)
    om = OracleMesh(cfg)
    u = s.create_solution()
    ou = OracleSolution(om)
    print(f"--- {tag}: B={s.num_blocks} N={s.block_size} regular={s.num_regular_blocks}")
    dt_o = om.cfg.cfl_number * ou.maximum_timestep()
    dt_g = om.cfg.cfl_number * s.maximum_timestep(u)
    print(f"  max_timestep rel err {abs(dt_g - dt_o) / dt_o:.2e}")
    o1, st = ou.advance(dt_o)
    try:
        g1 = s.advance(u, dt_o)
    except m3.NegativeDensity as e:
        print("  NEGATIVE", e.lines[:3]); g1 = e.solution
    print(f"  advance: field err {block_rel_err(g1.conserved_u, o1.conserved_u):.2e}")
    se = scalar_rel_err(g1.scalars, o1.scalars)
    k = int(se.argmax()); print(f"  advance: worst scalar {SCALAR_NAMES[k]} {se[k]:.2e} ({g1.scalars[k]} vs {o1.scalars[k]})")
    for n in range(steps):
        dto, fbo = ou.next_solution()
        dtg, fbg = s.next_solution(u)
        se = scalar_rel_err(u.scalars, ou.scalars); k = int(se.argmax())
        print(f"  step {n+1}: dt err {abs(dtg-dto)/dto:.2e} fb {fbo}/{fbg} field err {block_rel_err(u.conserved_u, ou.conserved_u):.2e} worst scalar {SCALAR_NAMES[k]} {se[k]:.2e}")


if __name__ == "__main__":
    check(dict(depth=2, block_size=64, domain_radius=6.0))
    check(dict(depth=2, block_size=64, domain_radius=6.0), general_only=True)
    check(dict(depth=4, block_size=24))
    check(dict(depth=6, block_size=16))
    check(dict(depth=3, block_size=8, eccentricity=0.3, mass_ratio=0.5, nu=0.01, alpha_cutoff_radius=1.0, begin_live_binary=0.0, density_floor=1e-2, mdot=1e-4))
    check(dict(depth=3, block_size=8, axisymmetric_cs2=1, counter_rotate=1, rk_order=1, fixed_dt=1, no_accretion_force=1))
    # throughput smoke
    for cfg in [dict(depth=4, block_size=64, focus_factor=1e3), dict(depth=6, block_size=64, focus_factor=1e3)]:
        s = m3.Solver(cfg); u = s.create_solution()
        s.run_steps(u, 3); s.synchronize()
        t0 = time.time(); n = 20; s.run_steps(u, n); s.synchronize(); t1 = time.time()
        print(cfg, "cells", s.num_cells, "Mzps", s.num_cells * n / (t1 - t0) * 1e-6)
