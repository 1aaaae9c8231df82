"""`binary key=value ...` on several GPUs: torchrun --nproc-per-node N tools/run_binary.py binary depth=6 ...
One process per GPU; rank 0 prints the run-loop lines and writes the (gathered) HDF5 products."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
import mara3_b200 as m3

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
uid = None
if world > 1:
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    box = [m3.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    uid = box[0]
code = m3.binary_main(sys.argv[1:], device=local, rank=rank, nranks=world, nccl_unique_id=uid)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
sys.exit(code)
