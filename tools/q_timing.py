"""Development aid: throughput of advance_q (conserve_linear_p=0) on a uniform grid, strip kernel (QMODE) against the any-tree kernels."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mara3_b200 as m3
for cfg in [dict(depth=4, block_size=64, focus_factor=1e3, conserve_linear_p=0, fixed_dt=1), dict(depth=6, block_size=64, focus_factor=1e3, conserve_linear_p=0, fixed_dt=1)]:
    s = m3.Solver(cfg); u = s.create_solution()
    s.run_steps(u, 5); s.synchronize()
    t0 = time.perf_counter(); n = 20; s.run_steps(u, n); s.synchronize(); t1 = time.perf_counter()
    print(os.environ.get("M3B_Q_STRIP", "1"), cfg["depth"], "blocks", s.num_blocks, "fused", s.num_regular_blocks, "cells", s.num_cells, "Mzps %.1f" % (s.num_cells * n / (t1 - t0) * 1e-6), "ms/step %.3f" % ((t1 - t0) / n * 1e3))
