import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mara3_b200 as m3
cfg = dict(depth=4, block_size=64, focus_factor=1e3, mach_number=10.0)
for pipe in (True, False):
    s = m3.Solver(cfg); s.set_pipelining(pipe); u = s.create_solution()
    s.run_steps(u, 20); s.synchronize()
    t0 = time.perf_counter(); s.run_steps(u, 500); s.synchronize(); t1 = time.perf_counter()
    print("pipelining", pipe, "run_steps: %.1f us/step" % ((t1 - t0) / 500 * 1e6), "launches/step", None)
    ts = []
    for k in range(10):
        a = time.perf_counter(); s.next_solution(u); ts.append((time.perf_counter() - a) * 1e6)
    print("   per-call wall us:", [round(x) for x in ts])
