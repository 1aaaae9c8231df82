"""Development aid: the reference's default configuration (depth=4, block_size=24: 64 leaves on levels 2-4) against the same tree
in 32^2 and 64^2 blocks -- steps queued back to back.  usage: python tools/default_config_timing.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mara3_b200 as m3
for cfg in (dict(), dict(block_size=32), dict(block_size=64), dict(block_size=24, depth=6), dict(block_size=32, depth=6)):
    s = m3.Solver(cfg); u = s.create_solution()
    s.run_steps(u, 20); s.synchronize()
    n = 400
    t0 = time.perf_counter(); s.run_steps(u, n); s.synchronize(); t1 = time.perf_counter()
    print(f"DEFAULT {cfg or 'defaults (depth=4 block_size=24)'}: blocks {s.num_blocks} (regular {s.num_regular_blocks}) cells {s.num_cells}: "
          f"{(t1 - t0) / n * 1e6:.1f} us/step, {s.num_cells * n / (t1 - t0) * 1e-6:.1f} Mzps")
