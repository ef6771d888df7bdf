"""per-step device times of the slab-decomposed 16M run (debug aid): torchrun ... scripts/slab_steps.py"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from realtimeparticles_b200 import _abi as abi, sharded
rank, lr, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
box, grid = (80, 40, 40), (240, 120, 120)
nx = 512 // world
x0 = -40.0 + rank * (80.0 / world)
pos = abi.gen_box_grid((nx, 256, 128), (x0, -20.0, -20.0), (x0 + 80.0 / world, 0.0, 0.0))
dev = torch.device("cuda", lr)
eng = sharded.CudaSlabEngine(int(len(pos) * 1.15) + (4 * 14400 * 40 if world > 1 else 0), box, grid, lr, jacobi=3)
sd = sharded.SlabDecomposition(eng, grid, rank, world)
sd.load_owned(torch.from_numpy(pos).to(dev), torch.zeros((len(pos), 4), device=dev))
sd.profile = True
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
for i in range(n):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    sd.step()
    torch.cuda.synchronize()
    w = (time.perf_counter() - t0) * 1e3
    ph = sd.stats["phases"]
    if rank == 0 and (w > 34 or i % 10 == 0):
        print(i, round(w, 1), "sum_phases", round(sum(ph.values()), 1), ph, sd.stats.get("migrated_out"), flush=True)
if world > 1:
    dist.destroy_process_group()
