"""Development aid: stage-kernel time on config 2 (and 3) for a few values of an environment knob.
usage: python tools/c2_sweep.py KNOB v1 v2 ..."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import sys, time; sys.path.insert(0, %r); import mara3_b200 as m3
for depth, n in ((4, 300), (6, 40)):
    s = m3.Solver(dict(depth=depth, block_size=64, focus_factor=1e3, mach_number=10.0)); u = s.create_solution()
    s.run_steps(u, 10); s.synchronize(); s.stage_timing(True)
    t0 = time.perf_counter(); s.run_steps(u, n); s.synchronize(); t1 = time.perf_counter()
    ms, k = s.stage_timing_read()
    print("  depth %%d: stage kernel %%.2f us, step %%.4f ms" %% (depth, ms / k * 1e3, (t1 - t0) / n * 1e3))
''' % ROOT
knob, values = sys.argv[1], sys.argv[2:]
for v in values:
    env = dict(os.environ); env[knob] = v
    print(knob, "=", v, flush=True)
    subprocess.run([sys.executable, "-c", code], env=env)
