"""Shared builders for the parity tests: the same initial state loaded into the oracle (CPU restatement) and into the
CUDA backend through the C ABI. Configs follow SURVEY.md section 8(d)."""
import numpy as np

from oracle import oracle_py as O
from realtimeparticles_b200 import _abi

M130K = 131072


def rel_err(a, b):
    """norm-wise relative error max|a-b| / max|b| (the 1e-5 bar of BASELINE.json's north_star)."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    den = max(np.abs(b).max(), 1e-30) if b.size else 1.0
    return float(np.abs(a - b).max() / den) if b.size else 0.0


def pad_positions(verts, M):
    pos = np.full((M, 4), np.inf, np.float32)
    pos[:, 3] = 0.0
    pos[:len(verts)] = verts
    return pos


def fluid_params(dim=3, **kw):
    o, a = O.default_fluid_params(dim), _abi.FluidParams(450.0, 600.0, 0.010, dim, 1, 0.006, 0.001, 4, 1, 0.0004, 0.0001)
    for k, v in kw.items():
        setattr(o, k, v)
        setattr(a, k, v)
    return o, a


def cloud_params(dim=3, **kw):
    o = O.default_cloud_params(dim)
    a = _abi.CloudParams(dim, 0.01, 450.0, 10.0, 0.10, 0.0005, 5.0, 0.3485, 0.07, 1, 600.0, 0.75, 1.0)
    for k, v in kw.items():
        setattr(o, k, v)
        setattr(a, k, v)
    return o, a


class Pair:
    """oracle world + CUDA handle holding identical state"""

    def __init__(self, model, M, N, box=(10, 10, 10), grid=(30, 30, 30), dim=3, cap=0, gpu=True):
        self.w = O.World(model, M, N, box, grid, dim, cap)
        self.h = _abi.Handle(model, M, N, box, grid, dim, cap) if gpu else None
        self.model, self.M, self.N = model, M, N

    def upload(self, fid, arr):
        self.w.upload(fid, arr)
        if self.h:
            self.h.upload(O.F[fid], arr)

    def both(self, fn_w, fn_h):
        fn_w(self.w)
        if self.h:
            fn_h(self.h)

    def step(self, flags=O.STEP_PHYSICS, cam=(32.0, -1.2, 0.0), oracle=True):
        if oracle:
            self.w.step(flags, cam)
        if self.h:
            self.h.step(flags, cam)
            self.h.sync()

    def get(self, fid):
        return self.w.download(fid), (self.h.download(O.F[fid]) if self.h else None)


def make_fluids(M=M130K, N=None, res=(64, 64, 32), start=(-5.0, -5.0, -5.0), end=(5.0, 0.0, 0.0), jacobi=3, box=(10, 10, 10),
                grid=(30, 30, 30), gpu=True, verts=None, **fp_kw):
    """Config 3: PBF dam break (Fluids.cpp:341-346), I Jacobi iterations."""
    if verts is None:
        verts = O.gen_box_grid(res, start, end)
    N = len(verts) if N is None else N
    p = Pair(O.FLUIDS, M, N, box, grid, gpu=gpu)
    fo, fa = fluid_params(**fp_kw)
    p.w.set_fluid_params(fo, jacobi)
    if p.h:
        p.h.set_fluid_params(fa, jacobi)
    p.upload("POS", pad_positions(verts[:N] if N <= len(verts) else verts, M))
    p.upload("VEL", np.zeros((M, 4), np.float32))
    p.both(lambda w: w.reset_ids(), lambda h: h.reset_ids())
    return p


def make_boids(M=M130K, N=512, res=(8, 8, 8), box=(10, 10, 10), grid=(30, 30, 30), gpu=True, dim=3, verts=None):
    """Configs 1/2: boids sphere (Boids.cpp:300-318), vel = pos."""
    if verts is None:
        b = [float(x) for x in box]
        verts = O.gen_sphere_grid(res, (b[0] / -6.0, b[1] / -6.0, b[2] / -6.0), (b[0] / 6.0, b[1] / 6.0, b[2] / 6.0))
    p = Pair(O.BOIDS, M, N, box, grid, dim=dim, gpu=gpu)
    pos = pad_positions(verts[:N], M)
    p.upload("POS", pos)
    p.upload("VEL", pos)
    p.both(lambda w: w.reset_ids(), lambda h: h.reset_ids())
    return p


def make_clouds(M=M130K, N=M130K, box=(10, 20, 10), grid=(30, 60, 30), jacobi=2, gpu=True, seed=1, region=None, **cp_kw):
    """Config 4: clouds cumulus region (Clouds.cpp:451-455), glibc rand() fill, T = env(y), vapor = 0.75 sat(T)."""
    b = [float(x) for x in box]
    if region is None:
        region = ((b[0] / -2.0, b[1] / -2.0, b[2] / -2.0), (b[0] / 2.0, b[1] / -4.0, b[2] / 2.0))
    verts = O.gen_random_box(N, region[0], region[1], seed)
    p = Pair(O.CLOUDS, M, N, box, grid, gpu=gpu)
    fo, fa = fluid_params()
    co, ca = cloud_params(**cp_kw)
    p.w.set_fluid_params(fo, jacobi)
    p.w.set_cloud_params(co)
    if p.h:
        p.h.set_fluid_params(fa, jacobi)
        p.h.set_cloud_params(ca)
    p.upload("POS", pad_positions(verts, M))
    p.upload("VEL", np.zeros((M, 4), np.float32))
    p.upload("CLOUD_DENS", np.zeros(M, np.float32))
    p.upload("PART_ID", np.arange(M, dtype=np.float32))
    p.both(lambda w: w.init_clouds_fields(), lambda h: h.init_clouds_fields())
    if p.h:
        # the init kernels are compared separately (test_clouds_init_fields); start the step from identical bits
        p.init_gpu = {f: p.h.download(O.F[f]) for f in ("TEMP", "VAPOR_DENS")}
        for f in ("TEMP", "VAPOR_DENS"):
            p.h.upload(O.F[f], p.w.download(f))
    p.both(lambda w: w.reset_ids(), lambda h: h.reset_ids())
    return p


def pbf_invariants(density, vel, N, rho0=450.0):
    d = np.asarray(density[:N], np.float64)
    v = np.asarray(vel[:N, :3], np.float64)
    return float(np.abs(d / rho0 - 1.0).mean()), float(0.5 * (v * v).sum())
