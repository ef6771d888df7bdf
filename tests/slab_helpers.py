"""Helpers for the slab-decomposition tests: a CPU engine (the oracle behind the SlabDecomposition engine interface) and
multi-process runners. TEST INFRASTRUCTURE (imports oracle/)."""
import contextlib
import os

import numpy as np
import torch

from oracle import oracle_py as O


class OracleSlabEngine:
    """The oracle as a local slab engine (CPU tensors): lets the gloo tests drive realtimeparticles_b200/sharded.py."""

    def __init__(self, capacity, box, grid, jacobi=3):
        self.w = O.World(O.FLUIDS, capacity, 0, box, grid)
        self.w.set_fluid_params(O.default_fluid_params(), jacobi)
        self.capacity, self.jacobi, self.vorticity = capacity, jacobi, True
        self.n_owned = 0

    def stream_context(self):
        return contextlib.nullcontext()

    def _t(self, name):
        a = self.w.field(name)
        if a.dtype == np.uint32:
            a = a.view(np.int32)
        return torch.from_numpy(a)

    def set_counts(self, n_owned, n_local):
        self.n_owned = n_owned
        self.w.set_nb_particles(n_local)

    def pos(self):
        return self._t("POS")

    def vel(self):
        return self._t("VEL")

    def keys_in(self):
        return self._t("CELL_ID")

    def pred_in(self):
        return self._t("PRED_POS")

    def pred_cur(self):
        return self._t("PRED_POS")

    # ---- exchange helpers (library kernels on the CUDA engine, torch indexing here)
    # (the CUDA engine refreshes |vorticity|, which it keeps as a scalar; the oracle's confinement sweep takes
    #  fast_length(vort[e]) of the vector, so here the vector travels)
    BUF = dict(PRED_IN="PRED_POS", PRED_CUR="PRED_POS", LAMBDA="CONST_FACTOR", VEL_SORTED="VEL", VORT4="VORT", VEL_CONFINED="VEL",
               POS="POS", VEL="VEL")
    ROW = dict(PRED_IN=4, PRED_CUR=4, LAMBDA=1, VEL_SORTED=4, VORT4=4, VEL_CONFINED=4, POS=4, VEL=4)

    def row_width(self, name):
        return self.ROW[name]

    def pack(self, name, idx, out):
        i = idx.to(torch.int64)
        ok = i >= 0
        t = self._t(self.BUF[name])
        out[ok] = t[i[ok]]
        if t.dim() == 2:
            out[~ok] = torch.tensor([float("inf")] * 3 + [0.0])
        else:
            out[~ok] = 0.0

    def unpack(self, name, idx, src):
        i = idx.to(torch.int64)
        ok = i >= 0
        self._t(self.BUF[name])[i[ok]] = src[ok]

    def inverse_perm(self, out):
        n = self.w.N
        perm = self._t("PERM")[:n].to(torch.int64)
        out[perm] = torch.arange(n, dtype=torch.int32)

    def check_ghosts(self, idx, next_epoch):
        pass  # the oracle has no neighbour lists

    def new(self, shape, dtype):
        return torch.empty(shape, dtype=dtype)

    def refresh_buffers(self, name, last=False):
        if name == "DENSITY_LAMBDA":
            return ["LAMBDA"]
        if name == "CORRECTION":
            return ["PRED_CUR"] + (["VEL_SORTED"] if last else [])
        if name == "VORTICITY":
            return ["VORT4"]
        if name == "CONFINEMENT":
            return ["VEL_CONFINED"]
        return []

    def stage(self, name, it=0, last=False):
        w = self.w
        seq = {"PREDICT": ["PREDICT_POS", "FILL_CELL_IDS"], "GHOST_KEYS": ["FILL_CELL_IDS"],
               "SORT": ["SORT_BY_CELL", "BUILD_CELL_TABLE"],
               "DENSITY_LAMBDA": ["APPLY_BOUNDARY", "DENSITY", "CONSTRAINT_FACTOR"],
               "CORRECTION": ["CONSTRAINT_CORRECTION", "CORRECT_POS"] + (["UPDATE_VEL"] if last else []),
               "VORTICITY": ["VORTICITY"], "CONFINEMENT": ["VORTICITY_CONFINEMENT"], "XSPH": ["XSPH", "UPDATE_POS"],
               "DROP_GHOSTS": []}[name]
        if name == "DROP_GHOSTS":  # owned particles (unsorted index < n_owned) first, sorted order kept
            n_loc = self.w.N
            keep = np.nonzero(self.w.field("PERM")[:n_loc] < self.n_owned)[0]
            for f in ("POS", "VEL"):
                a = self.w.field(f)
                a[:len(keep)] = a[keep]
            return
        if name == "PREDICT":
            w.reset_ids()
        for s in seq:
            w.run_stage(s)

    def sync(self):
        pass


def match_particles(a, b):
    """max distance between two particle sets matched by nearest neighbour (must be a bijection)"""
    from scipy.spatial import cKDTree
    assert a.shape == b.shape, (a.shape, b.shape)
    d, j = cKDTree(b[:, :3]).query(a[:, :3])
    assert len(np.unique(j)) == len(j), "not a one-to-one match"
    return float(d.max()), j


def _cpu_worker(rank, world, port, cfg, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    from realtimeparticles_b200 import sharded
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pos0 = cfg["pos"]
    eng = OracleSlabEngine(cfg["capacity"], cfg["box"], cfg["grid"], cfg["jacobi"])
    sd = sharded.SlabDecomposition(eng, cfg["grid"], rank, world)
    mine = sharded.split_initial_state(pos0, cfg["box"], cfg["grid"], rank, world)
    sd.load_owned(torch.from_numpy(pos0[mine]), torch.from_numpy(cfg["vel"][mine]))
    hist = []
    for _ in range(cfg["steps"]):
        sd.step()
        hist.append(dict(sd.stats))
    p, v = sd.owned_state()
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), pos=p.numpy(), vel=v.numpy(), migrated=sum(h.get("migrated_out", 0) for h in hist),
             ghosts=hist[-1].get("ghosts", 0))
    dist.barrier()
    dist.destroy_process_group()


def _gpu_worker(rank, world, port, cfg, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    from realtimeparticles_b200 import sharded
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    pos0 = cfg["pos"]
    eng = sharded.CudaSlabEngine(cfg["capacity"], cfg["box"], cfg["grid"], rank, jacobi=cfg["jacobi"])
    sd = sharded.SlabDecomposition(eng, cfg["grid"], rank, world)
    mine = sharded.split_initial_state(pos0, cfg["box"], cfg["grid"], rank, world)
    sd.load_owned(torch.from_numpy(pos0[mine]).cuda(rank), torch.from_numpy(cfg["vel"][mine]).cuda(rank))
    hist = []
    for _ in range(cfg["steps"]):
        sd.step()
        hist.append(dict(sd.stats))
    p, v = sd.owned_state()
    eng.sync()
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), pos=p.cpu().numpy(), vel=v.cpu().numpy(),
             migrated=sd.stats.get("migrated_out_total", sum(h.get("migrated_out", 0) for h in hist)), ghosts=hist[-1].get("ghosts", 0))
    dist.barrier()
    dist.destroy_process_group()


def run_sharded(worker, world, cfg, out_dir, port=29571):
    import torch.multiprocessing as mp
    mp.spawn(worker, args=(world, port, cfg, out_dir), nprocs=world, join=True)
    parts = [np.load(os.path.join(out_dir, "rank%d.npz" % r)) for r in range(world)]
    return (np.concatenate([p["pos"] for p in parts]), np.concatenate([p["vel"] for p in parts]),
            int(sum(p["migrated"] for p in parts)), [int(p["ghosts"]) for p in parts], [len(p["pos"]) for p in parts])
