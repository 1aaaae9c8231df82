#!/usr/bin/env python
"""bench.py -- zone-updates/s of the `binary` iso2d step on N B200s of one node.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU path (oracle/_ref)

A "step" is one binary::next_solution: CFL dt + all RK stages + the RK combination
(reference kzps definition, Mara3 src/subprog_binary.cpp:394-404, I/O excluded).
Mzps = leaf cells x steps / seconds / 1e6.  Rank 0 prints ONE JSON line.

Timing, N = 1: CUDA events around every step on the stream the kernels are launched
on, summed over the K timed steps; L2 is flushed between steps by overwriting a
512 MiB buffer outside the event pairs, because the N=1 workload (1024^2 cells, 25 MB
per state copy) would otherwise live in the 126 MB L2.  N > 1 (4096^2 cells, per-rank
working set larger than L2): one event pair around the K steps, queued back to back as
the subprogram's run loop queues them; max over ranks (--sync-steps: the N = 1 procedure).
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
}
ALGORITHMIC_BYTES_PER_CELL_STEP = 120.0     # RK2 = 2 x (24 read + 24 write) + 24 re-read of U^n
ALGORITHMIC_BYTES_PER_CELL_LAUNCH = 60.0    # mean over the two stage launches of a step (48 and 72)


def ncu_traffic(workload, local_cells):
    """DRAM bytes per stage launch from the committed ncu --set full capture (profiles/r01_traffic.json), if it is of this workload."""
    try:
        with open(os.path.join(ROOT, "profiles", "r01_traffic.json")) as f:
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


def time_reference(config, steps, warmup, threads):
    """Mzps of the reference's own thread-pooled CPU implementation (oracle/_ref, built from
    /root/reference/src by oracle/Makefile); falls back to the plain-C port if it is not there."""
    path, name = reference_binary()
    if path:
        cmd = [path, "--steps", str(steps), "--warmup", str(warmup), "--timing"] + [f"{k}={v}" for k, v in config.items()] + [f"threaded={threads}"]
        out = subprocess.run(cmd, capture_output=True, text=True, check=True).stdout
        line = [l for l in out.splitlines() if l.startswith("timing:")][0]
        mzps = float(line.split("mzps=")[1])
        flags = "-O3 -march=x86-64-v3" if name == "mara_ref_fast" else "-O2"
        return mzps, "reference", threads, f"oracle/_ref/{name} (g++ {flags}), threaded={threads}"
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_util import OracleMesh, OracleSolution
    mesh = OracleMesh(config)
    sol = OracleSolution(mesh)
    for _ in range(warmup):
        sol.next_solution()
    t0 = time.time()
    for _ in range(steps):
        sol.next_solution()
    mzps = mesh.B * mesh.N ** 2 * steps / (time.time() - t0) * 1e-6
    return mzps, "port", 1, "oracle/libm3b_oracle.so (plain C, single thread)"


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    wl_name = args.workload or ("c2" if world == 1 else "c3")
    wl = WORKLOADS[wl_name]
    threads = os.cpu_count() or 1
    steps, warmup = args.steps, args.warmup
    cells = (2 ** wl["config"]["depth"] * wl["config"]["block_size"]) ** 2
    if wl_name == "c4":
        cells = 424 * 64 * 64       # leaves of the default-focus depth-8 tree x block cells
    # bounded sample: the reference manages ~1 Mzps, keep the run to a few minutes
    max_steps = max(1, int(120e6 / cells))
    sample_steps = min(steps, max_steps)
    sample_warmup = min(warmup, 1)
    t0 = time.time()
    mzps, kind, cores, how = time_reference(wl["config"], sample_steps, sample_warmup, threads)
    sample = f"{sample_steps} timed steps after {sample_warmup} warm-up of the full workload; {how}"
    print(json.dumps({
        "impl": "reference", "metric": "iso2d zone-updates/sec", "value": mzps, "unit": "Mzps", "n_gpus": world,
        "steps": sample_steps, "warmup": sample_warmup, "ms_per_step": cells / (mzps * 1e6) * 1e3, "higher_is_better": True,
        "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["name"], "keys": wl["config"], "cells": cells},
        "cpu_baseline": {"value": mzps, "unit": "Mzps", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": mzps, "unit": "Mzps", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.time() - t0,
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-flush", action="store_true", help="do not flush L2 between timed steps")
    ap.add_argument("--sync-steps", action="store_true", help="N>1: one host round trip and one L2 flush per timed step, as at N=1")
    ap.add_argument("--no-scaling-reference", action="store_true", help="N>1: skip the single-GPU run of the same workload on rank 0")
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
    uid = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        box = [mara3_b200.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        uid = box[0]

    wl_name = args.workload or ("c2" if world == 1 else "c3")
    wl = WORKLOADS[wl_name]
    scaling_reference = None
    if world > 1 and rank == 0 and not args.no_scaling_reference:
        # the same workload on ONE GPU, measured in this run, so that strong-scaling efficiency can be
        # computed against the same problem (the N=1 default of this script is the smaller config 2)
        ref = mara3_b200.Solver(wl["config"], device=local_rank)
        ref_u = ref.create_solution()
        ref.run_steps(ref_u, 3)
        ref.synchronize()
        t0 = time.perf_counter()
        ref.run_steps(ref_u, 10)
        ref.synchronize()
        scaling_reference = {"n_gpus": 1, "value": ref.num_cells * 10 / (time.perf_counter() - t0) * 1e-6, "unit": "Mzps",
                             "how": "10 steps of the same workload on rank 0's GPU alone, host clock, no L2 flush"}
        del ref_u, ref
    if world > 1:
        dist.barrier()

    solver = mara3_b200.Solver(wl["config"], device=local_rank, rank=rank, nranks=world, nccl_unique_id=uid)
    solution = solver.create_solution()
    cells = solver.num_cells
    flush = None if args.no_flush else torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device="cuda")

    def flush_l2():
        # the step queued ahead by next_solution must have drained first, or the fill would run beside it
        torch.cuda.synchronize()
        if flush is not None:
            flush.fill_(1)
            torch.cuda.synchronize()

    # ---- device-resident throughput (`value`): K steps, CUDA events per step on the solver's stream
    for _ in range(args.warmup):
        solver.next_solution(solution)
    solver.synchronize()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()

    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = solver.kernel_launches
    solver.stage_timing(True)
    fallbacks = 0
    wall0 = time.time()
    stream = torch.cuda.Stream()                # a real stream handle: the legacy default stream is 0
    solver.set_stream(stream.cuda_stream)       # so torch.cuda.Event sees the stream the kernels run on
    events = []
    # per rank: three state buffers, initial state and buffer rate of the owned cells (the same number on every rank: the ranks must agree)
    working_set_mb = cells / world * 8 * (3 * 3 + 3 + 1) / 1e6
    back_to_back = world > 1 and not args.sync_steps and working_set_mb > 126
    if back_to_back:
        # several ranks: the K steps are queued back to back, as the subprogram's run loop does (m3b_run_steps: the next step is
        # launched before the host reads the last one's dt); the per-rank working set is larger than the 126 MB L2
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        fallbacks += solver.run_steps(solution, args.steps)
        e1.record(stream)
        events.append((e0, e1))
    else:
        for _ in range(args.steps):
            flush_l2()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            _, fb = solver.next_solution(solution)
            e1.record(stream)
            events.append((e0, e1))
            fallbacks += fb
    solver.synchronize()
    torch.cuda.synchronize()
    wall = time.time() - wall0
    stage_ms, stage_launches = solver.stage_timing_read()
    solver.stage_timing(False)
    launches = solver.kernel_launches - launches0
    step_ms = torch.tensor([a.elapsed_time(b) for a, b in events], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.barrier()
        dist.all_reduce(step_ms, op=dist.ReduceOp.MAX)      # a step is as slow as its slowest rank
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
    # what the timed stage kernels update on this rank: every owned cell (on a nested tree the timed region spans the regular
    # blocks' launch, the jump blocks' launch beside it and the ring gradients they read)
    fused_cells = local_cells
    nested = solver.num_regular_blocks < solver.num_blocks
    achieved = fused_cells * ALGORITHMIC_BYTES_PER_CELL_LAUNCH / (kernel_ms * 1e-3) * 1e-9 if stage_launches else None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
                "traffic": ncu_traffic(wl_name, local_cells), "kernel": "stage_strip + stage_strip<JUMP> + general_gradients_ring (one RK stage)" if nested else "stage_strip", "kernel_ms": kernel_ms, "launches_timed": stage_launches,
                "algorithmic_bytes_per_launch": fused_cells * ALGORITHMIC_BYTES_PER_CELL_LAUNCH, "cells_per_launch": fused_cells, "per": "GPU (rank 0)", "peak_source": peak_how,
                "step_frac_of_hbm_roofline": value / world * 1e6 * ALGORITHMIC_BYTES_PER_CELL_STEP / (peak * 1e9)}

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
        e2e = {"value": cells * n_e2e / float(dt_e2e) * 1e-6, "unit": "Mzps", "h2d_bytes_per_step": int(nbytes), "d2h_bytes_per_step": int(nbytes),
               "steps": n_e2e, "api": "m3b_next_solution_host (pinned host buffers; every rank moves the blocks it owns, bytes summed over ranks)"}

    exchange = None
    if world > 1:
        # NVLink side of the step: two guard-zone exchanges (one per RK stage) + one result all-gather
        exchange = {"halo_bytes_sent_per_exchange_rank0": solver.halo_bytes_per_exchange, "exchanges_per_step": 2,
                    "transport": {"peer": "peer memory over NVLink (CUDA IPC mailboxes): halo_push / halo_wait_unpack kernels, no NCCL call in the step loop",
                                  "nccl": "NCCL send/recv, grouped per stage, on the compute stream"}.get(solver.exchange_transport, solver.exchange_transport)}
        dist.barrier()
        if rank != 0:
            dist.destroy_process_group()
            return
    cpu_baseline = None
    if not args.no_cpu_baseline and world == 1:
        threads = os.cpu_count() or 1
        sample_steps = max(1, int(12e6 / cells))
        mzps, kind, cores, how = time_reference(wl["config"], sample_steps, 1, threads)
        cpu_baseline = {"value": mzps, "unit": "Mzps", "cores": cores, "kind": kind,
                        "sample": f"{sample_steps} timed steps after 1 warm-up of the same workload; {how}"}

    print(json.dumps({
        "metric": "iso2d zone-updates/sec", "value": value, "unit": "Mzps", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if world > 1 else "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["name"], "keys": wl["config"], "cells": cells, "blocks": solver.num_blocks,
                   "l2": (f"not flushed: the per-rank working set ({working_set_mb:.0f} MB) is larger than the 126 MB L2" if back_to_back else
                          "not flushed" if args.no_flush else "flushed between steps (512 MiB fill, outside the per-step timing)"),
                   "timing": ("CUDA events on the launch stream around the K steps, queued back to back as the subprogram's run loop queues them (m3b_run_steps); max over ranks"
                              if back_to_back else "CUDA events on the launch stream around each step (a step ends with the host reading dt and the validation flag)")},
        "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
        "safe_mode_retries": fallbacks, "wall_s": wall, "exchange": exchange, "scaling_reference": scaling_reference,
    }))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
