"""Cost of running the sweeps by row phase, without any transport in the picture: two slab ranks on ONE GPU (LocalSlabGroup),
overlap on / off, device time per step. Under `ncu --metrics gpu__time_duration.sum` the launch list gives the kernel
times of both schedules (scripts/kernel_times.py).   python scripts/phase_overhead.py [particles_x] [overlap]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from realtimeparticles_b200 import sharded  # noqa: E402

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 192
overlap = bool(int(sys.argv[2])) if len(sys.argv) > 2 else True
BOX, GRID = (80, 40, 40), (240, 120, 120)
# a block of nx x 96 x 96 particles at the lattice spacing of the 16M dam (2 per cell edge), centred on the slab face x = 0
sp = BOX[0] / GRID[0] / 2.0
ix, iy, iz = np.meshgrid(np.arange(nx), np.arange(96), np.arange(96), indexing="ij")
pos0 = np.zeros((nx * 96 * 96, 4), np.float32)
pos0[:, 0] = (ix.ravel() - nx / 2 + 0.5) * sp
pos0[:, 1] = -BOX[1] / 2 + 0.5 + iy.ravel() * sp
pos0[:, 2] = (iz.ravel() - 48 + 0.5) * sp
vel0 = np.zeros_like(pos0)
n = len(pos0)
G = 2 * 96 * 96 * 8
sds = []
for r in range(2):
    mine = sharded.split_initial_state(pos0, BOX, GRID, r, 2)
    cap = int(len(mine) * 1.1) + 2 * 8192 + 2 * G
    eng = sharded.CudaSlabEngine(cap, BOX, GRID, 0, jacobi=3)
    sd = sharded.SlabDecomposition(eng, GRID, rank=r, world=2, ghost_cap=G, migrate_cap=8192, overlap=overlap)
    sd.load_owned(torch.from_numpy(pos0[mine]).cuda(), torch.from_numpy(vel0[mine]).cuda())
    sds.append(sd)
grp = sharded.LocalSlabGroup(sds)
for _ in range(3):
    grp.step()
torch.cuda.synchronize()
t0 = time.perf_counter()
K = 5
for _ in range(K):
    grp.step()
torch.cuda.synchronize()
print("particles", n, "overlap", overlap, "ms/step (both ranks, serialised exchanges)", round((time.perf_counter() - t0) / K * 1e3, 3))
