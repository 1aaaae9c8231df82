"""Robustness check (one process per GPU under torch.distributed.run, 2 ranks): rank 1 stops stepping after three steps;
rank 0 must come back from its next step with an error that names the stalled peer -- not hang on a flag.
The deadline is shortened through M3B_SPIN_DEADLINE_MS (read when the solver is created).  Covers the default path (peer-memory
transport, steps queued ahead); with M3B_TRANSPORT=nccl or pipelining off the results travel through NCCL collectives, whose
waits are NCCL's own."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("M3B_SPIN_DEADLINE_MS", "1500")
import torch, torch.distributed as dist
import mara3_b200 as m3

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
box = [m3.nccl_unique_id() if rank == 0 else None]
dist.broadcast_object_list(box, src=0)
s = m3.Solver(dict(depth=4, block_size=64, focus_factor=1e3), device=local, rank=rank, nranks=world, nccl_unique_id=box[0])
u = s.create_solution()
for _ in range(3):
    s.next_solution(u)
s.synchronize()
dist.barrier()
verdict = "FAIL: no error"
if rank == 0:
    t0 = time.time()
    try:
        for _ in range(3):
            s.next_solution(u)
        s.synchronize()
    except m3.Mara3Error as e:
        took = time.time() - t0
        ok = "rank 1" in str(e) and "deadline" in str(e) and took < 30.0
        verdict = ("PASS" if ok else "FAIL") + f": after {took:.1f} s: {e}"
    print("STALLED_PEER_CHECK", verdict, flush=True)
else:
    time.sleep(8.0)         # the stalled peer: alive, but it takes no step
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if verdict.startswith("PASS") or rank != 0 else 1)
