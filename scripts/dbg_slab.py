import sys, numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from oracle import oracle_py as O
from realtimeparticles_b200 import _abi, sharded
import slab_helpers as SH
BOX, GRID = (10, 10, 10), (30, 30, 30)
world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
pos0 = O.gen_box_grid((48, 32, 32), (-5.0, -5.0, -5.0), (3.0, 0.0, 0.0))
vel0 = np.zeros_like(pos0); vel0[:, 0] = 12.0
n = len(pos0); jacobi = 3
sds = []
for r in range(world):
    sd = sharded.SlabDecomposition(sharded.CudaSlabEngine(n, BOX, GRID, 0, jacobi=jacobi), GRID, rank=r, world=world)
    mine = sharded.split_initial_state(pos0, BOX, GRID, r, world)
    sd.load_owned(torch.from_numpy(pos0[mine]).cuda(), torch.from_numpy(vel0[mine]).cuda())
    sds.append(sd)
grp = sharded.LocalSlabGroup(sds)
h = _abi.Handle(_abi.FLUIDS, n, n, BOX, GRID)
h.set_fluid_params(_abi.FluidParams(450.0, 600.0, 0.010, 3, 1, 0.006, 0.001, 4, 1, 0.0004, 0.0001), jacobi)
h.upload("p_pos", pos0); h.upload("p_vel", vel0); h.reset_ids()
for s in range(steps):
    grp.step()
    h.step(_abi.STEP_PHYSICS); h.sync()
    parts = [sd.owned_state() for sd in sds]
    for sd in sds: sd.e.sync()
    pos = np.concatenate([p.cpu().numpy() for p, _ in parts])
    ref = h.download("p_pos")
    from scipy.spatial import cKDTree
    d, j = cKDTree(ref[:, :3]).query(pos[:, :3])
    print(s, "owned", [sd.n_owned for sd in sds], "sum", len(pos), "finite", np.isfinite(pos).all(), "dmax %.3g" % d.max(), "unique", len(np.unique(j)), "migr", [sd.stats.get("migrated_out") for sd in sds])
