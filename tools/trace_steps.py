"""Development aid: per-segment device timeline of the pipelined step (M3B_TRACE=1), single or multi rank."""
import os, sys, time
os.environ["M3B_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
import mara3_b200 as m3
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
depth = int(sys.argv[1]) if len(sys.argv) > 1 else 6
uid = None
if world > 1:
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    box = [m3.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    uid = box[0]
s = m3.Solver(dict(depth=depth, block_size=64, focus_factor=1e3, mach_number=10.0), device=local, rank=rank, nranks=world, nccl_unique_id=uid)
u = s.create_solution()
s.run_steps(u, 10); s.synchronize()
t0 = time.perf_counter(); n = 200; s.run_steps(u, n); s.synchronize(); t1 = time.perf_counter()
if rank == 0: print(f"depth {depth} world {world}: {(t1 - t0) / n * 1e3:.3f} ms/step, {4 ** depth * 4096 * n / (t1 - t0) * 1e-9:.2f} Gzps")
del u, s
if world > 1: dist.destroy_process_group()
