import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import mara3_b200 as m3
from oracle_util import OracleMesh, OracleSolution, SCALAR_NAMES

def show(cfg, nsteps, stage_only=False):
    s = m3.Solver(cfg); u = s.create_solution(); o = OracleSolution(OracleMesh(cfg))
    print("=====", cfg, "steps", nsteps)
    if stage_only:
        dt = 0.4 * o.maximum_timestep()
        o, _ = o.advance(dt); u = s.advance(u, dt)
    else:
        for _ in range(nsteps):
            s.next_solution(u); o.next_solution()
    for n, a, b in list(zip(SCALAR_NAMES, u.scalars, o.scalars))[:23]:
        print(f"{n:36s} {a: .17e} {b: .17e} abs {abs(a-b):.2e} rel {abs(a-b)/max(abs(b),1e-300):.2e}")

show(dict(depth=3, block_size=16, alpha=0.0, sink_rate=5.0, sink_radius=0.1, softening_radius=0.1, buffer_damping_rate=0.0, mach_number=5.0, plm_theta=1.0, focus_factor=1e3), 0, stage_only=True)
