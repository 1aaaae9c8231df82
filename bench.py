#!/usr/bin/env python
"""bench.py -- zone-updates/s of the `binary` iso2d step on N B200s of one node.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU path (oracle/_ref)

A "step" is one binary::next_solution: CFL dt + all RK stages + the RK combination
(reference kzps definition, Mara3 src/subprog_binary.cpp:394-404, I/O excluded).
Mzps = leaf cells x steps / seconds / 1e6.  Rank 0 prints ONE JSON line.

ONE workload and ONE procedure at every N (so that the 1 -> 8 GPU curve is a strong-scaling curve): BASELINE.json
configs[2], uniform 4096^2 cells; after W warm-up steps the K timed steps are queued back to back, as the subprogram's
run loop queues them (m3b_run_steps), between one pair of CUDA events on the stream the kernels are launched on; max over
ranks.  The working set per GPU (three state copies + initial state + buffer rate) is larger than the 126 MB L2 at every
N <= 8, so nothing is flushed.  At N = 1 the line also carries, as `secondary`, the same workload with one host round trip
and one L2 flush per step, and configs[1] (1024^2) under both procedures.  At N > 1 rank 0 also advances the same initial
data on its GPU alone and the gathered N-rank state must equal it (`parity_vs_1gpu`, exit code 1 beyond 1e-12).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# BASELINE.json configs[1] (N=1 workload) and configs[2] (strong scaling, N>1); keys of subprog_binary.cpp:57-99
WORKLOADS = {
    "c1": dict(name="binary iso2d disk 256^2 (depth=2, 16 blocks of 64^2), RK2", config=dict(depth=2, block_size=64)),
    "c2": dict(name="binary iso2d uniform 1024^2 (depth=4, 256 blocks of 64^2), mach 10, sinks+buffer on, PLM+HLLE, RK2",
               config=dict(depth=4, block_size=64, focus_factor=1e3, mach_number=10.0)),
    "c3": dict(name="binary iso2d uniform 4096^2 (depth=6, 4096 blocks of 64^2), PLM+HLLE, RK2",
               config=dict(depth=6, block_size=64, focus_factor=1e3, mach_number=10.0)),
    "c4": dict(name="binary iso2d nested quadtree (depth=8, 424 blocks of 64^2 on levels 3-8, prolong/restrict at the jumps), PLM+HLLE, RK2",
               config=dict(depth=8, block_size=64)),
    "c5": dict(name="binary iso2d uniform 16384^2 (depth=8, 65536 blocks of 64^2), PLM+HLLE, RK2",
               config=dict(depth=8, block_size=64, focus_factor=1e3, mach_number=10.0)),
    # a nested tree with real load for 8 GPUs ("raise block_size / focus_factor"): 4552 leaves of 64^2 =
    # 18.6 M cells on levels 4-8 (144 + 260 + 392 + 668 + 3088), finest spacing 24 / (64 * 256)
    "c4x": dict(name="binary iso2d nested quadtree (depth=8, focus_factor=12: 4552 blocks of 64^2 on levels 4-8, prolong/restrict at the jumps), PLM+HLLE, RK2",
                config=dict(depth=8, block_size=64, focus_factor=12.0)),
}
DEFAULT_WORKLOAD = "c3"        # BASELINE.json configs[2]: the configuration the metric ("at 1/2/4/8 B200") is quoted on
C4X_BLOCKS = 4552
ALGORITHMIC_BYTES_PER_CELL_STEP = 120.0     # RK2 = 2 x (24 read + 24 write) + 24 re-read of U^n
ALGORITHMIC_BYTES_PER_CELL_LAUNCH = 60.0    # mean over the two stage launches of a step (48 and 72)


def ncu_traffic(workload, local_cells):
    """DRAM bytes per stage launch from the committed ncu --set full capture (profiles/r02_traffic.json), if it is of this workload."""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_traffic.json")) as f:
            t = json.load(f)
        if t["workload"] == workload and t["algorithmic_bytes_per_launch"] == local_cells * ALGORITHMIC_BYTES_PER_CELL_LAUNCH:
            return t["traffic_bytes_per_launch"]
    except (OSError, KeyError, ValueError):
        pass
    return None


def measured_hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return None
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        for line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return None
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(smax), "reasons": sorted(reasons), "samples": len(sm)}


def reference_binary():
    for name in ("mara_ref_fast", "mara_ref"):
        path = os.path.join(ROOT, "oracle", "_ref", name)
        if os.path.exists(path) and os.access(path, os.X_OK):
            probe = subprocess.run([path, "--mesh-only", "depth=1", "block_size=8"], capture_output=True)
            if probe.returncode == 0:       # the -march=x86-64-v3 build needs AVX2/FMA on the host
                return path, name
    return None, None


def time_reference(config, steps, warmup, threads, max_seconds=None):
    """Mzps of the reference's own thread-pooled CPU implementation (oracle/_ref, built from
    /root/reference/src by oracle/Makefile); falls back to the plain-C port if it is not there."""
    path, name = reference_binary()
    if path:
        cmd = [path, "--steps", str(steps), "--warmup", str(warmup), "--timing"] + (["--max-seconds", str(max_seconds)] if max_seconds else []) \
              + [f"{k}={v}" for k, v in config.items()] + [f"threaded={threads}"]
        out = subprocess.run(cmd, capture_output=True, text=True, check=True).stdout
        line = [l for l in out.splitlines() if l.startswith("timing:")][0]
        mzps = float(line.split("mzps=")[1].split()[0])
        done = int(line.split("steps=")[1].split()[0]) if "steps=" in line else steps
        flags = "-O3 -march=x86-64-v3" if name == "mara_ref_fast" else "-O2"
        return mzps, "reference", threads, f"oracle/_ref/{name} (g++ {flags}), threaded={threads}", done
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_util import OracleMesh, OracleSolution
    mesh = OracleMesh(config)
    sol = OracleSolution(mesh)
    for _ in range(warmup):
        sol.next_solution()
    t0 = time.time()
    done = 0
    for _ in range(steps):
        sol.next_solution()
        done += 1
        if max_seconds and time.time() - t0 > max_seconds:
            break
    steps = done
    mzps = mesh.B * mesh.N ** 2 * steps / (time.time() - t0) * 1e-6
    return mzps, "port", 1, "oracle/libm3b_oracle.so (plain C, single thread)", steps


def workload_cells(wl_name):
    wl = WORKLOADS[wl_name]
    if wl_name == "c4":
        return 424 * 64 * 64, 424       # leaves of the default-focus depth-8 tree x block cells
    if wl_name == "c4x":
        return C4X_BLOCKS * 64 * 64, C4X_BLOCKS
    side = 2 ** wl["config"]["depth"] * wl["config"]["block_size"]
    return side * side, 4 ** wl["config"]["depth"]


def workload_config(wl_name, cells, blocks, world):
    """`config` of the JSON line: the workload and the procedure, the same dict in both arms."""
    wl = WORKLOADS[wl_name]
    working_set_mb = cells / world * 8 * (3 * 3 + 3 + 1) / 1e6
    return {"workload": wl["name"], "keys": wl["config"], "cells": cells, "blocks": blocks,
            "l2": f"not flushed: the working set per GPU at {world} GPU(s), {working_set_mb:.0f} MB "
                  f"(three state copies, initial state, buffer rate), is {'larger' if working_set_mb > 126 else 'SMALLER'} than the 126 MB L2",
            "timing": "K steps queued back to back after W warm-up steps, one pair of clock readings around them "
                      "(GPU arm: CUDA events on the launch stream, max over ranks; reference arm: the harness' own clock)"}


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    wl_name = args.workload or DEFAULT_WORKLOAD
    wl = WORKLOADS[wl_name]
    threads = os.cpu_count() or 1
    cells, blocks = workload_cells(wl_name)
    t0 = time.time()
    # --steps / --warmup are honoured; a wall-clock budget bounds the sample (the reference manages ~10 Mzps on 16 cores)
    mzps, kind, cores, how, done = time_reference(wl["config"], args.steps, args.warmup, threads, max_seconds=args.reference_budget)
    sample = f"{done} timed steps after {args.warmup} warm-up of the full workload" + \
             (f" (of {args.steps} asked for: stopped by the {args.reference_budget:.0f} s budget)" if done < args.steps else "") + f"; {how}"
    print(json.dumps({
        "impl": "reference", "metric": "iso2d zone-updates/sec", "value": mzps, "unit": "Mzps", "n_gpus": world,
        "steps": done, "steps_requested": args.steps, "warmup": args.warmup, "ms_per_step": cells / (mzps * 1e6) * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(wl_name, cells, blocks, world),
        "cpu_baseline": {"value": mzps, "unit": "Mzps", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": mzps, "unit": "Mzps", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.time() - t0,
    }))


def time_io(solver, solution, rank, world, dist, torch, mara3_b200):
    """One checkpoint, one diagnostics file and one time-series sample (the three tasks of subprog_binary.cpp:326-386),
    wall clock between barriers, apart from the step timing.  Files go to a scratch directory and are removed."""
    import shutil
    import tempfile
    box = [tempfile.mkdtemp(prefix="m3b_bench_io_") if rank == 0 else None]
    if world > 1:
        dist.broadcast_object_list(box, src=0)
    outdir = box[0]

    def timed(fn):
        solver.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        fn()
        solver.synchronize()
        if world > 1:
            dist.barrier()
        return time.perf_counter() - t0

    chk, diag = os.path.join(outdir, "chkpt.0000.h5"), os.path.join(outdir, "diagnostics.0000.h5")
    t_chk = timed(lambda: solver.write_checkpoint(solution, chk))
    t_diag = timed(lambda: solver.write_diagnostics(solution, diag))
    t_ts = timed(lambda: solver.time_series_sample(solution))
    out = None
    if rank == 0:
        nb_chk, nb_diag = os.path.getsize(chk), os.path.getsize(diag)
        out = {"checkpoint_s": t_chk, "checkpoint_bytes": nb_chk, "checkpoint_gbs": nb_chk / t_chk * 1e-9,
               "diagnostics_s": t_diag, "diagnostics_bytes": nb_diag, "diagnostics_gbs": nb_diag / t_diag * 1e-9,
               "time_series_sample_s": t_ts, "where": "local scratch directory (tempfile), files removed afterwards",
               "layout": "chkpt.NNNN.h5 / diagnostics.NNNN.h5 of subprog_binary_io.cpp:131-172, one file, every rank writes its own blocks' byte ranges"}
        shutil.rmtree(outdir, ignore_errors=True)
    return out


def timed_steps(solver, solution, steps, stream, torch):
    """K steps back to back between two CUDA events on the launch stream -> (ms, safe-mode retries)."""
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    fallbacks = solver.run_steps(solution, steps)
    e1.record(stream)
    solver.synchronize()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1), fallbacks


def sync_flush_steps(solver, solution, steps, stream, torch, flush):
    """The round-1 N=1 procedure: one host round trip (dt, validation flag) and one L2 flush per step, events per step."""
    total = 0.0
    for _ in range(steps):
        torch.cuda.synchronize()
        flush.fill_(1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        solver.next_solution(solution)
        e1.record(stream)
        solver.synchronize()
        torch.cuda.synchronize()
        total += e0.elapsed_time(e1)
    return total


def block_rel_err(a, b):
    scale = abs(b).max(axis=(-2, -1), keepdims=True)
    scale[scale == 0] = 1.0
    return float((abs(a - b) / scale).max())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="N=1: skip the sync+flush variant and the 1024^2 workload")
    ap.add_argument("--no-parity", action="store_true", help="N>1: skip the comparison with the same steps on one GPU")
    ap.add_argument("--parity-steps", type=int, default=4)
    ap.add_argument("--reference-budget", type=float, default=240.0, help="--impl reference: wall-clock bound of the timed steps, seconds")
    ap.add_argument("--io", action="store_true", help="also time one checkpoint + one diagnostics file + one time-series sample (reported separately)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return
    if args.warmup < 3:
        args.warmup = 3

    import numpy as np
    import torch
    import mara3_b200

    torch.cuda.set_device(local_rank)
    dist = None

    def fresh_id():
        # an NCCL unique id serves exactly one communicator
        if world == 1:
            return None
        box = [mara3_b200.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        return box[0]

    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    wl_name = args.workload or DEFAULT_WORKLOAD
    wl = WORKLOADS[wl_name]

    # ---- N > 1: the same workload on ONE GPU (rank 0's), in this run: strong-scaling anchor and parity reference
    scaling_reference, ref_state, ref_scalars = None, None, None
    if world > 1 and rank == 0 and not args.no_parity:
        ref = mara3_b200.Solver(wl["config"], device=local_rank)
        ref_u = ref.create_solution()
        ref.run_steps(ref_u, args.parity_steps)
        ref.synchronize()
        ref_state, ref_scalars = ref_u.conserved_u, ref_u.scalars
        ref.run_steps(ref_u, 3)
        ref.synchronize()
        t0 = time.perf_counter()
        ref.run_steps(ref_u, 10)
        ref.synchronize()
        scaling_reference = {"n_gpus": 1, "value": ref.num_cells * 10 / (time.perf_counter() - t0) * 1e-6, "unit": "Mzps",
                             "how": "10 steps of the same workload on rank 0's GPU alone, host clock"}
        del ref_u, ref
    if world > 1:
        dist.barrier()

    solver = mara3_b200.Solver(wl["config"], device=local_rank, rank=rank, nranks=world, nccl_unique_id=fresh_id())
    cells = solver.num_cells
    blocks_total = solver.num_global_blocks if world > 1 else solver.num_blocks

    # ---- N > 1: the N-rank state after P steps must be the single-GPU state (the exchange moves bits, nothing else differs)
    parity = None
    if world > 1 and not args.no_parity:
        pu = solver.create_solution()
        solver.run_steps(pu, args.parity_steps)
        solver.synchronize()
        mine = torch.from_numpy(pu.conserved_u).cuda()
        counts = [blocks_total * (r + 1) // world - blocks_total * r // world for r in range(world)]
        pad = torch.zeros((max(counts),) + tuple(mine.shape[1:]), dtype=torch.float64, device="cuda")
        pad[:mine.shape[0]] = mine
        parts = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(parts, pad)
        scal = torch.from_numpy(pu.scalars).cuda()
        allscal = [torch.empty_like(scal) for _ in range(world)]
        dist.all_gather(allscal, scal)
        if rank == 0:
            got = torch.cat([p[:c] for p, c in zip(parts, counts)]).cpu().numpy()
            err = block_rel_err(got, ref_state)
            parity = {"max_rel": err, "steps": args.parity_steps, "bit_identical": bool(np.array_equal(got, ref_state)),
                      "scalars_equal_on_all_ranks": all(bool(torch.equal(allscal[0], x)) for x in allscal),
                      "scalars_max_abs_diff": float(np.abs(allscal[0].cpu().numpy() - ref_scalars).max()),
                      "norm": "max over cells of |N-rank - 1-GPU| / max|1-GPU| over the cell's block and field"}
            del got
        del pu, mine, pad, parts
        ref_state = None
        dist.barrier()

    solution = solver.create_solution()
    stream = torch.cuda.Stream()                # a real stream handle: the legacy default stream is 0
    solver.set_stream(stream.cuda_stream)       # so torch.cuda.Event sees the stream the kernels run on

    # ---- device-resident throughput (`value`)
    solver.run_steps(solution, args.warmup)
    solver.synchronize()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()

    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = solver.kernel_launches
    solver.stage_timing(True)
    wall0 = time.time()
    ms, fallbacks = timed_steps(solver, solution, args.steps, stream, torch)
    wall = time.time() - wall0
    stage_ms, stage_launches = solver.stage_timing_read()
    solver.stage_timing(False)
    launches = solver.kernel_launches - launches0
    step_ms = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.barrier()
        dist.all_reduce(step_ms, op=dist.ReduceOp.MAX)      # the run is as slow as its slowest rank
    total_ms = float(step_ms.sum())

    # the timed region of a small workload is shorter than nvidia-smi's sampling period: keep the same load
    # running (untimed, same count on every rank) until the sampler has seen about a second of it
    extra_steps = min(20000, max(0, int((1200.0 - total_ms) / (total_ms / args.steps))))
    try:
        solver.run_steps(solution, extra_steps)
    except mara3_b200.Mara3Error:
        pass
    solver.synchronize()
    clocks = sampler.stop()
    if clocks is not None:
        clocks["window"] = f"timed region + {extra_steps} further identical steps (untimed)"
    ms_per_step = total_ms / args.steps
    value = cells * args.steps / (total_ms * 1e-3) * 1e-6

    peak, peak_how = measured_hbm_peak()
    kernel_ms = stage_ms / max(1, stage_launches)
    local_cells = solver.num_owned_cells
    nested = solver.num_regular_blocks < solver.num_blocks
    achieved = local_cells * ALGORITHMIC_BYTES_PER_CELL_LAUNCH / (kernel_ms * 1e-3) * 1e-9 if stage_launches else None
    stage_kernel = "stage_strip" if os.environ.get("M3B_STAGE") == "strip" else "stage_tma"
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
                "traffic": ncu_traffic(wl_name, local_cells),
                "kernel": (stage_kernel + " + stage_strip<JUMP> + general_gradients_ring (one RK stage)") if nested else stage_kernel,
                "kernel_ms": kernel_ms, "launches_timed": stage_launches,
                "algorithmic_bytes_per_launch": local_cells * ALGORITHMIC_BYTES_PER_CELL_LAUNCH, "cells_per_launch": local_cells,
                "per": "GPU (rank 0)", "peak_source": peak_how,
                "step_frac_of_hbm_roofline": value / world * 1e6 * ALGORITHMIC_BYTES_PER_CELL_STEP / (peak * 1e9)}

    # ---- N = 1: the same workload with a host round trip and an L2 flush per step, and configs[1] under both procedures
    secondary = None
    if world == 1 and not args.no_secondary:
        flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device="cuda")
        n2 = max(3, min(args.steps, 20))
        ms2 = sync_flush_steps(solver, solution, n2, stream, torch, flush)
        secondary = {f"{wl_name}_sync_flush": {"value": cells * n2 / (ms2 * 1e-3) * 1e-6, "unit": "Mzps", "ms_per_step": ms2 / n2, "steps": n2,
                                                "procedure": "CUDA events around each step; one host round trip and one 512 MiB L2 flush (untimed) per step"}}
        if wl_name != "c2":
            small = mara3_b200.Solver(WORKLOADS["c2"]["config"], device=local_rank)
            small.set_stream(stream.cuda_stream)
            su = small.create_solution()
            small.run_steps(su, args.warmup)
            small.synchronize()
            small.stage_timing(True)
            msb, _ = timed_steps(small, su, args.steps, stream, torch)
            sms, sk = small.stage_timing_read()
            small.stage_timing(False)
            mss = sync_flush_steps(small, su, n2, stream, torch, flush)
            c2 = small.num_cells
            secondary["c2"] = {"workload": WORKLOADS["c2"]["name"], "value": c2 * args.steps / (msb * 1e-3) * 1e-6, "unit": "Mzps",
                               "ms_per_step": msb / args.steps, "steps": args.steps, "kernel_ms": sms / max(1, sk),
                               "roofline_frac": c2 * ALGORITHMIC_BYTES_PER_CELL_LAUNCH / (sms / max(1, sk) * 1e-3) / (peak * 1e9),
                               "procedure": "steps queued back to back (its 220 MB working set is about the size of the L2: partly L2-resident)"}
            secondary["c2_sync_flush"] = {"value": c2 * n2 / (mss * 1e-3) * 1e-6, "unit": "Mzps", "ms_per_step": mss / n2, "steps": n2,
                                          "procedure": "as c3_sync_flush (the round-1 bench line)"}
            del su, small
        del flush

    # ---- checkpoint / diagnostics / time-series sample, timed apart from the steps (BASELINE.json configs[4])
    io = None
    if args.io:
        io = time_io(solver, solution, rank, world, dist, torch, mara3_b200)

    # ---- end to end through the C ABI with host buffers (H2D + step + D2H inside the timed region)
    e2e = None
    if not args.no_e2e:
        # every rank moves the blocks it owns: host -> device, one full step, device -> host; wall clock between barriers
        shape = (solver.num_blocks, 3, solver.block_size, solver.block_size)
        u_in = torch.empty(shape, dtype=torch.float64).pin_memory()
        u_out = torch.empty(shape, dtype=torch.float64).pin_memory()
        u_in.numpy()[...] = solution.conserved_u
        scalars = solution.scalars
        n_e2e = max(3, min(args.steps, 20))
        for _ in range(2):
            solver.next_solution_host(u_in.numpy(), scalars, out=u_out.numpy())
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            _, scalars_out, _, _ = solver.next_solution_host(u_in.numpy(), scalars, out=u_out.numpy())
            u_in, u_out = u_out, u_in
            scalars = scalars_out
        torch.cuda.synchronize()
        dt_e2e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        nbytes = torch.tensor([int(np.prod(shape)) * 8 + 43 * 8], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(dt_e2e, op=dist.ReduceOp.MAX)
            dist.all_reduce(nbytes, op=dist.ReduceOp.SUM)
        per_gpu_gbs = 2 * float(nbytes) / world * n_e2e / float(dt_e2e) * 1e-9
        e2e = {"value": cells * n_e2e / float(dt_e2e) * 1e-6, "unit": "Mzps", "h2d_bytes_per_step": int(nbytes), "d2h_bytes_per_step": int(nbytes),
               "steps": n_e2e, "host_link_gbs_per_gpu": per_gpu_gbs,
               "bound": "PCIe: at 55 GB/s per direction per GPU the two copies alone allow "
                        f"{cells / (2 * float(nbytes) / world / 55e9) * 1e-6:.0f} Mzps",
               "api": "m3b_next_solution_host (pinned host buffers; every rank moves the blocks it owns, bytes summed over ranks)"}

    exchange = None
    if world > 1:
        # NVLink side of the step: two guard-zone exchanges (one per RK stage) + one result exchange
        exchange = {"halo_bytes_sent_per_exchange_rank0": solver.halo_bytes_per_exchange, "exchanges_per_step": 2,
                    "transport": {"peer": "peer memory over NVLink (CUDA IPC mailboxes): halo_push / halo_wait_unpack kernels, no NCCL call in the step loop",
                                  "nccl": "NCCL send/recv, grouped per stage, on the compute stream"}.get(solver.exchange_transport, solver.exchange_transport)}
        exchange.update(solver.exchange_timing())
        dist.barrier()
        if rank != 0:
            dist.destroy_process_group()
            return
    cpu_baseline = None
    if not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        sample_steps = max(1, int(100e6 / cells))       # ~10-15 s of host work at the reference's ~8 Mzps on 16 cores
        mzps, kind, cores, how, done = time_reference(wl["config"], sample_steps, 1, threads, max_seconds=40.0)
        cpu_baseline = {"value": mzps, "unit": "Mzps", "cores": cores, "kind": kind,
                        "sample": f"{done} timed steps after 1 warm-up of the same workload; {how}"}

    print(json.dumps({
        "metric": "iso2d zone-updates/sec", "value": value, "unit": "Mzps", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": workload_config(wl_name, cells, blocks_total, world),
        "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
        "safe_mode_retries": fallbacks, "wall_s": wall, "exchange": exchange, "scaling_reference": scaling_reference,
        "parity_vs_1gpu": parity, "secondary": secondary, "io": io,
    }))
    if world > 1:
        dist.destroy_process_group()
    if parity is not None and not parity["max_rel"] <= 1e-12:
        sys.exit(1)


if __name__ == "__main__":
    main()
