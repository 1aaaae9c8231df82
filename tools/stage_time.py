"""Development aid: CUDA-event time of the stage kernel on one workload, steps queued back to back.
usage: python tools/stage_time.py [c2|c3|c4|c5] [steps]   (environment: M3B_STAGE, M3B_TMA_CTAS, ...)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mara3_b200 as m3
from bench import WORKLOADS

name = sys.argv[1] if len(sys.argv) > 1 else "c3"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 20
s = m3.Solver(WORKLOADS[name]["config"]); u = s.create_solution()
s.run_steps(u, 3); s.synchronize()
s.stage_timing(True)
t0 = time.perf_counter(); s.run_steps(u, n); s.synchronize(); t1 = time.perf_counter()
ms, k = s.stage_timing_read(); s.stage_timing(False)
tag = " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("M3B_"))
print(f"STAGE {name} [{tag}] cells {s.num_cells}: stage {ms / k * 1e3:.1f} us ({s.num_cells / (ms / k) * 1e-6:.2f} Gcell-stage/s, "
      f"{s.num_cells * 60 / (ms / k * 1e-3) / 6548.2e9:.3f} of HBM roofline), step {(t1 - t0) / n * 1e3:.3f} ms, {s.num_cells * n / (t1 - t0) * 1e-9:.2f} Gzps")
