"""Multi-GPU parity check (one process per GPU, launched by torch.distributed.run):
the distributed solver against the single-GPU solver and the plain-C oracle on the same inputs.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py
"""
import json
import os
import sys
import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import mara3_b200 as m3


def block_rel_err(a, b):
    scale = np.abs(b).max(axis=(-2, -1), keepdims=True)
    return float((np.abs(a - b) / np.where(scale > 0, scale, 1.0)).max())


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    report = {}

    def fresh_id():
        # an NCCL unique id serves exactly one communicator
        box = [m3.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        return box[0]

    cases = [(dict(depth=3, block_size=32, focus_factor=1e3, domain_radius=6.0), 12, True),
             (dict(depth=2, block_size=64), 34, True),             # runs into the safe-mode retries (steps 23-33)
             (dict(depth=4, block_size=64, focus_factor=1e3), 5, False),
             (dict(depth=4, block_size=32), 6, True),              # nested tree: refinement jumps across the rank boundary
             (dict(depth=6, block_size=64), 4, False),             # config-4-like nesting (136 leaves, levels 2-6)
             (dict(depth=3, block_size=16, conserve_linear_p=0, fixed_dt=1), 4, True),   # advance_q: whole-block ghosts everywhere
             (dict(depth=4, block_size=32, conserve_linear_p=0, fixed_dt=1), 3, False)]  # advance_q: strip kernel (QMODE) + any-tree at the jumps
    if os.environ.get("MGC_QUICK"):
        cases = [(dict(depth=3, block_size=32, focus_factor=1e3, domain_radius=6.0), 6, False), (dict(depth=6, block_size=64), 3, False)]
    for cfg, steps, with_oracle in cases:
        s = m3.Solver(cfg, device=local, rank=rank, nranks=world, nccl_unique_id=fresh_id())
        u = s.create_solution()
        fallbacks, dts = 0, []
        for _ in range(steps):
            dt, fb = s.next_solution(u)
            fallbacks += fb
            dts.append(dt)
        mine = torch.from_numpy(u.conserved_u).cuda()
        parts = [torch.empty((n, 3, s.block_size, s.block_size), dtype=torch.float64, device="cuda")
                 for n in [len(range(s.num_global_blocks * r // world, s.num_global_blocks * (r + 1) // world)) for r in range(world)]]
        dist.all_gather(parts, mine)
        scal = torch.from_numpy(u.scalars).cuda()
        allscal = [torch.empty_like(scal) for _ in range(world)]
        dist.all_gather(allscal, scal)
        same_scalars = all(bool(torch.equal(allscal[0], x)) for x in allscal)     # every rank did the same bookkeeping

        if rank == 0:
            U = torch.cat(parts).cpu().numpy()
            one = m3.Solver(cfg, device=local)
            v = one.create_solution()
            fb1, dts1 = 0, []
            for _ in range(steps):
                dt, fb = one.next_solution(v)
                fb1 += fb
                dts1.append(dt)
            entry = {"blocks": s.num_global_blocks, "steps": steps, "fallbacks": [fallbacks, fb1],
                     "field_err_vs_single_gpu": block_rel_err(U, v.conserved_u),
                     "dt_err_vs_single_gpu": float(np.max(np.abs(np.array(dts) - np.array(dts1)) / np.array(dts1))),
                     "scalar_err_vs_single_gpu": float(np.max(np.abs(u.scalars - v.scalars) / np.maximum(np.abs(v.scalars), 1e-3))),
                     "ranks_agree_on_scalars": same_scalars}
            if with_oracle:
                from oracle_util import OracleMesh, OracleSolution
                o = OracleSolution(OracleMesh(cfg))
                fbo = 0
                for _ in range(steps):
                    fbo += o.next_solution()[1]
                entry["field_err_vs_oracle"] = block_rel_err(U, o.conserved_u)
                entry["fallbacks"].append(fbo)
            report[json.dumps(cfg)] = entry
        dist.barrier()
    if rank == 0:
        print(json.dumps(report, indent=1))
        ok = all(e["field_err_vs_single_gpu"] <= 1e-12 and e["ranks_agree_on_scalars"] and len(set(e["fallbacks"])) == 1
                 and e.get("field_err_vs_oracle", 0.0) <= 1e-9 for e in report.values())
        print("MULTI_GPU_CHECK", "PASS" if ok else "FAIL")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
