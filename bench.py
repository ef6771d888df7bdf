#!/usr/bin/env python
"""Headless benchmark of the hot path (BASELINE.json metric: particle-updates/sec, % of HBM roofline).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

N = 1 (default): PBF 3D dam break, 131072 particles, 3 Jacobi iterations, vorticity confinement + XSPH
                 (BASELINE.json configs[2], the config its 2000 steps/s target is quoted on).
A "step" is one pass of the hot path (the physics part of Fluids::update(), Fluids.cpp:400-457) over the whole state.

JSON keys (one line on stdout, rank 0):
  value      whole-job particle-updates/s, state resident in HBM, CUDA-event timed on the launching stream, L2 flushed
             between timed steps (the 17 MB state would otherwise live in the 126 MB L2)
  e2e        same metric through the public API (models.Fluids.update()) with HOST buffers: H2D of p_pos/p_vel from
             pinned memory and D2H of p_pos inside the timed region, every step
  roofline   dominant kernel: algorithmic bytes per launch / its event-timed duration vs MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the reference's kernels compiled for the CPU (oracle/_ref; else the oracle port) on this box's host cores, bounded sample
--impl reference times that CPU port alone, all host threads, on the same workload and metric.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N130K = 131072
FLUID_DEFAULTS = (450.0, 600.0, 0.010, 3, 1, 0.006, 0.001, 4, 1, 0.0004, 0.0001)
JACOBI = 3

# algorithmic bytes per particle per launch; t = cell-table bytes per particle per pass
def algorithmic_bytes(ncells, n, sort_passes=2):
    t = 8.0 * ncells / n
    return {
        "predictPosition+fillCellIDs": 48 + 20,
        "radixSort(onesweep)": 4 + 16 * sort_passes,
        "gather+cellTable": 100 + 4 + t,
        "adjustEndCell": t,
        "densityLambda": (20 + t) + (24 + t),  # fld_computeDensity + fld_computeConstraintFactor, fused
        "correction": (36 + t) + 48 + 32,  # constraintCorrection + correctPosition + boundary (fused)
        "vorticity": 48 + t,
        "confinement": 64 + t,
        "xsph": 32 + (48 + t) + 32,  # copy + xsph + updatePosition (fused)
    }


PBF_BYTES_PER_PARTICLE_STEP = 486.6 + 164.9 * JACOBI  # 981 B at I=3, P=2


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons during the timed region (NVML, ~5 ms period)."""

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.samples, self.reasons, self.stop_flag, self.max_mhz, self.ok = [], set(), False, None, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.dev = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.dev, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake": 0x80, "sync_boost": 0x10, "applications_clocks": 0x2}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.dev, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.dev)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.005)

    def result(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s)}


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


# ------------------------------------------------------------------------------------------------ reference arm

def cpu_world(pos, which):
    """The 130k PBF workload on the host cores: which = "reference" -> the reference's OWN kernel sources compiled for the
    CPU (oracle/_ref/libref_kernels.so, built by oracle/ref/Makefile where /root/reference exists; the .so travels),
    "port" -> the oracle restatement (oracle/rtp_oracle.c). Returns (world, threads) or None when not available."""
    import ctypes
    import numpy as np
    from oracle import oracle_py as O
    n_thr = os.cpu_count() or 1
    if which == "reference":
        from oracle import ref_py as R
        if not R.available():
            return None
        mod = R
        try:  # torchrun exports OMP_NUM_THREADS=1: use every host core anyway
            ctypes.CDLL("libgomp.so.1").omp_set_num_threads(n_thr)
        except OSError:
            n_thr = int(os.environ.get("OMP_NUM_THREADS", n_thr))
    else:
        mod = O
        O.lib().orc_set_threads(n_thr)
        n_thr = O.max_threads()
    w = mod.World(O.FLUIDS, N130K, N130K)
    w.set_fluid_params(O.default_fluid_params(), JACOBI)
    w.upload("POS", pos)
    w.upload("VEL", np.zeros((N130K, 4), np.float32))
    w.reset_ids()
    return w, n_thr


def parity_figures(abi, pos, device):
    """CUDA path vs the CPU checkers on ONE step of the 130k dam break from the same initial state: max-norm relative error
    per field against the reference's own kernel sources (oracle/_ref, IEEE arithmetic without contraction) and against the
    oracle restatement (canonical fused arithmetic). Part of the cpu_baseline leg: the checkers are
    never on the product path."""
    import numpy as np
    from oracle import oracle_py as O
    h = abi.Handle(abi.FLUIDS, N130K, N130K, (10, 10, 10), (30, 30, 30), 3, 0, device)
    h.set_fluid_params(abi.FluidParams(*FLUID_DEFAULTS), JACOBI)
    h.upload("p_pos", pos)
    h.upload("p_vel", np.zeros((N130K, 4), np.float32))
    h.reset_ids()
    h.step(abi.STEP_PHYSICS)
    h.sync()
    ours = {f: h.download(n) for f, n in (("POS", "p_pos"), ("VEL", "p_vel"), ("DENSITY", "p_density"), ("CONST_FACTOR", "p_constFactor"),
                                           ("CELL_ID", "p_cellID"), ("PERM", "RadixSortIndices"), ("START_END_CELL", "c_startEndPartID"))}
    h.close()
    out = {"step": "1 step of pbf_dam_130k_I3_vorticity_xsph", "norm": "max|a-b| / max|b|"}
    for kind in ("reference", "port"):
        cw = cpu_world(pos, kind)
        if cw is None:
            continue
        w = cw[0]
        w.step(O.STEP_PHYSICS)
        r = {}
        for f, a in ours.items():
            b = w.download(f)
            if a.dtype.kind in "iu":
                r[f] = "bit-exact" if np.array_equal(a, b) else "%d mismatches" % int((a != b).sum())
            else:
                r[f] = float(np.abs(a.astype(np.float64) - b).max() / max(np.abs(b).max(), 1e-30))
        out["vs_reference_kernels_on_cpu" if kind == "reference" else "vs_oracle_port"] = r
    return out


def reference_opencl_figures(pos, steps=12, warm=4):
    """The reference's unmodified .cl kernels on this box's OpenCL GPU device (oracle/_ref/libref_ocl.so, NVIDIA's OpenCL
    driver, the reference's build options): device time of the MODEL kernels per step from OpenCL profiling events. The
    reference's radix sort is replaced by a host-side stable sort in this runner, so its 2 x 21 sort launches per frame
    are NOT in the figure: it is a lower bound of the reference's step time on this GPU, reported beside the CPU arm."""
    import numpy as np
    from oracle import ocl_py
    if not ocl_py.available():
        raise RuntimeError("oracle/_ref/libref_ocl.so not built")
    r = ocl_py.OclFluids(N130K, N130K, jacobi=JACOBI)
    r.upload("p_pos", pos)
    r.upload("p_vel", np.zeros((N130K, 4), np.float32))
    r.reset_ids()
    for _ in range(warm):
        r.step()
    r.kernel_times(reset=True)
    for _ in range(steps):
        r.step()
    t = r.kernel_times()
    per_step_us = {k: round(v[0] / steps, 2) for k, v in t.items() if v[1]}
    total = sum(per_step_us.values())
    out = {"device": r.device, "steps": steps, "kernel_us_per_step": per_step_us, "model_kernels_ms_per_step": round(total * 1e-3, 4),
           "value": N130K / (total * 1e-6), "unit": "particle-updates/s (model kernels only, sort excluded)",
           "build_options": "-cl-denorms-are-zero -cl-fast-relaxed-math + the -D constants of Fluids.cpp:104-119"}
    r.close()
    return out


def time_cpu_world(w, warmup, steps, budget_s):
    from oracle import oracle_py as O
    for _ in range(warmup):
        w.step(O.STEP_PHYSICS)
    done, t0 = 0, time.perf_counter()
    while done < steps and (time.perf_counter() - t0) < budget_s:
        w.step(O.STEP_PHYSICS)
        done += 1
    return done, time.perf_counter() - t0


def headline_config(world):
    """`config` of the JSON line: the same dict for both arms (the reference arm runs on OUR arm's config: the l2 /
    parallelism entries describe how the GPU arm is timed)."""
    return {"workload": "pbf_dam_130k_I3_vorticity_xsph", "particles": N130K, "jacobi_iterations": JACOBI,
            "grid": [30, 30, 30], "box": [10, 10, 10], "l2": "flushed before every timed step (256 MiB memset, untimed)",
            "parallelism": "replicas" if world > 1 else "single"}


def run_reference(args):
    """The reference's CPU implementation of the path. The OpenCL runtime it needs does not exist here, so its kernel
    sources (physics/ocl/kernels/*.cl, unmodified) are compiled for the CPU through oracle/ref/ocl_shim.hpp into
    oracle/_ref/libref_kernels.so and run over all host threads (one OpenMP work-item loop per kernel launch, the same
    launch sequence as Fluids::update); when that library is absent the oracle port is timed instead. Bounded to ~150 s."""
    rank, _, world = dist_env()
    if rank != 0:
        return
    from oracle import oracle_py as O
    pos = O.gen_box_grid((64, 64, 32), (-5.0, -5.0, -5.0), (5.0, 0.0, 0.0))
    kind = "reference"
    cw = cpu_world(pos, kind)
    if cw is None:
        kind = "port"
        cw = cpu_world(pos, kind)
    w, cores = cw
    budget = 150.0
    wu = min(args.warmup, 5)  # (a CPU step takes 0.3 s: the driver's W = 5 is honoured)
    done, dt = time_cpu_world(w, wu, args.steps, budget)
    value = N130K * done / dt
    sample = "%d of %d requested steps of the full 131072-particle PBF step (time-bounded to %ds)" % (done, args.steps, int(budget))
    out = {
        "impl": "reference", "metric": "particle-updates/sec", "value": value, "unit": "particle-updates/s",
        "n_gpus": args.gpus, "steps": done, "warmup": wu, "ms_per_step": 1e3 * dt / max(done, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": headline_config(max(args.gpus, 1)),
        "cpu_baseline": {"value": value, "unit": "particle-updates/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "particle-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if kind == "reference":  # for orientation: the fused restatement of the same kernels (the oracle) on the same cores
        w2, c2 = cpu_world(pos, "port")
        d2, t2 = time_cpu_world(w2, 1, max(2, min(done, 6)), 30.0)
        out["oracle_port"] = {"value": N130K * d2 / t2, "unit": "particle-updates/s", "cores": c2, "steps": d2}
    print(json.dumps(out), flush=True)


# ------------------------------------------------------------------------------------------------ our arm

# BASELINE.json configs -> (particles, algorithmic bytes per particle-step from BASELINE.md section 3)
WORKLOADS = {
    "pbf_dam_130k_I3_vorticity_xsph": (N130K, 486.6 + 164.9 * 3),
    "boids_130k": (N130K, 275.0),
    "clouds_130k_I2": (N130K, 823.0 + 218.0 * 2),
    "pbf_dam_16m_I3_vorticity_xsph": (1 << 24, 486.6 + 164.9 * 3 + 16.0),
}


def make_workload(abi, name, device):
    """Synthetic initial state of a BASELINE.json config loaded into a fresh handle."""
    import numpy as np
    n = WORKLOADS[name][0]
    if name.startswith("pbf_dam_130k"):
        pos = abi.gen_box_grid((64, 64, 32), (-5.0, -5.0, -5.0), (5.0, 0.0, 0.0))
        h = abi.Handle(abi.FLUIDS, n, n, (10, 10, 10), (30, 30, 30), 3, 0, device)
        h.set_fluid_params(abi.FluidParams(*FLUID_DEFAULTS), JACOBI)
        vel = np.zeros((n, 4), np.float32)
    elif name.startswith("pbf_dam_16m"):
        pos = abi.gen_box_grid((512, 256, 128), (-40.0, -20.0, -20.0), (40.0, 0.0, 0.0))
        h = abi.Handle(abi.FLUIDS, n, n, (80, 40, 40), (240, 120, 120), 3, 0, device)
        h.set_fluid_params(abi.FluidParams(*FLUID_DEFAULTS), JACOBI)
        vel = np.zeros((n, 4), np.float32)
    elif name.startswith("boids"):
        pos = abi.gen_sphere_grid((64, 64, 32), (-10 / 6.0,) * 3, (10 / 6.0,) * 3)
        h = abi.Handle(abi.BOIDS, n, n, (10, 10, 10), (30, 30, 30), 3, 0, device)
        vel = pos
    elif name.startswith("clouds"):
        pos = abi.gen_random_box(n, (-5.0, -10.0, -5.0), (5.0, -5.0, 5.0), 1)
        h = abi.Handle(abi.CLOUDS, n, n, (10, 20, 10), (30, 60, 30), 3, 0, device)
        h.set_fluid_params(abi.FluidParams(*FLUID_DEFAULTS), 2)
        h.set_cloud_params(abi.CloudParams(3, 0.01, 450.0, 10.0, 0.10, 0.0005, 5.0, 0.3485, 0.07, 1, 600.0, 0.75, 1.0))
        vel = np.zeros((n, 4), np.float32)
    else:
        raise SystemExit("unknown workload " + name)
    h.upload("p_pos", pos)
    h.upload("p_vel", vel)
    if name.startswith("clouds"):
        h.upload("p_cloudDens", np.zeros(n, np.float32))
        h.upload("p_partID", np.arange(n, dtype=np.float32))
        h.init_clouds_fields()
    h.reset_ids()
    return h, pos


def make_pbf(abi, device):
    return make_workload(abi, "pbf_dam_130k_I3_vorticity_xsph", device)


def quick_bench(abi, torch, name, device, steps, warmup, flush):
    """steps/s of another BASELINE.json config (same timing rules: events on the library stream, L2 flush per step)"""
    dev = torch.device("cuda", device)
    h, _ = make_workload(abi, name, device)
    n, bpp = WORKLOADS[name]
    stream = torch.cuda.ExternalStream(h.stream(), device=dev)
    with torch.cuda.stream(stream):
        h.step_n(warmup, abi.STEP_PHYSICS)
    h.sync()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    with torch.cuda.stream(stream):
        for a, b in ev:
            if n * 150 < (100 << 20):
                flush.zero_()
            a.record(stream)
            h.step_n(1, abi.STEP_PHYSICS)
            b.record(stream)
    h.sync()
    ms = sum(a.elapsed_time(b) for a, b in ev)
    peak, _ = measured_peak()
    out = {"particles": n, "steps": steps, "warmup": warmup, "steps_per_s": round(steps / (ms * 1e-3), 2),
           "value": n * steps / (ms * 1e-3), "unit": "particle-updates/s", "launches_per_step": h.last_launch_count(),
           "algorithmic_bytes_per_particle": bpp, "whole_step_frac_of_hbm": round(n * steps / (ms * 1e-3) * bpp / 1e9 / peak, 4)}
    if name.startswith("pbf_dam_16m"):
        # where HBM is really in play (the state is 20x the L2): per-kernel algorithmic GB/s, events between launches of the
        # next 4 steps after the timed window
        h.enable_profiling(True)
        acc, cnt, reps = {}, {}, 4
        for _ in range(reps):
            h.step(abi.STEP_PHYSICS)
            for nm, v in h.stage_times():
                acc[nm] = acc.get(nm, 0.0) + v
                cnt[nm] = cnt.get(nm, 0) + 1
        h.enable_profiling(False)
        ab = algorithmic_bytes(240 * 120 * 120, n, sort_passes=3)
        tot = sum(acc.values()) / reps
        kern = {}
        for nm in acc:
            per = acc[nm] / cnt[nm]
            gbs = ab[nm] * n / (per * 1e-3) / 1e9 if nm in ab else None
            kern[nm] = {"ms_per_launch": round(per, 4), "launches_per_step": cnt[nm] // reps, "share": round(acc[nm] / reps / tot, 4),
                        "algorithmic_GBps": None if gbs is None else round(gbs, 1), "frac_of_hbm": None if gbs is None else round(gbs / peak, 4)}
        out["kernels_16m"] = kern
        dom = max((k for k in kern if k in ab), key=lambda k: kern[k]["share"])
        out["roofline_16m"] = {"bound": "hbm", "kernel": dom, "achieved": kern[dom]["algorithmic_GBps"], "peak": peak, "unit": "GB/s",
                               "frac": kern[dom]["frac_of_hbm"], "window": "steps %d-%d of the 16M dam break, un-graphed launches" % (
                                   warmup + steps + 1, warmup + steps + reps)}
    h.close()
    return out


def run_slab_16m(args, torch, abi, rank, local_rank, world, steps=10, warmup=8):
    """BASELINE.json configs[4]: PBF dam break scaled to 2^24 particles (box 80x40x40, grid 240x120x120), x-slab
    decomposed over `world` GPUs with halo exchange + migration (realtimeparticles_b200/sharded.py). Strong scaling:
    the total problem is fixed; value = 2^24 * steps / max-over-ranks device time. 8 warm-up steps: the first step in
    which particles migrate (step 7 of this initial state) pays a one-off ~90 ms set-up; the same step window
    (steps 10-19) is timed for every N, N = 1 included (same code path with no neighbour), so the driver can form
    strong-scaling efficiencies from like-for-like numbers; `invariants` lets it check that every N computed the same
    physics."""
    import numpy as np
    import torch.distributed as dist
    from realtimeparticles_b200 import sharded
    box, grid, total = (80, 40, 40), (240, 120, 120), 1 << 24
    nx = 512 // world  # lattice planes per slab: spacing 80/512 = 0.15625 is exact in fp32, so sub-blocks reproduce the global lattice
    x0 = -40.0 + rank * (80.0 / world)
    pos = abi.gen_box_grid((nx, 256, 128), (x0, -20.0, -20.0), (x0 + 80.0 / world, 0.0, 0.0))
    n_own = len(pos)
    # fixed capacities per slab face: a ghost region of 2 x-layers (120 x 120 cells each, 4.9 particles per cell in the
    # initial lattice; room for 8) and a migration message
    ghost_cap = sharded.GHOST_LAYERS * 120 * 120 * 8 if world > 1 else 0
    migrate_cap = (1 << 15) if world > 1 else 1
    capacity = int(n_own * 1.10) + 2 * migrate_cap + 2 * ghost_cap
    dev = torch.device("cuda", local_rank)
    eng = sharded.CudaSlabEngine(capacity, box, grid, local_rank, jacobi=JACOBI)
    sd = sharded.SlabDecomposition(eng, grid, rank, world, ghost_cap=ghost_cap, migrate_cap=migrate_cap)
    sd.load_owned(torch.from_numpy(pos).to(dev), torch.zeros((n_own, 4), device=dev))
    sd.warm_up_code_paths()
    for _ in range(warmup):
        sd.step()
    sd.profile = True
    sd.step()
    phases = sd.stats.get("phases")
    if world > 1:
        allph = [None] * world
        dist.all_gather_object(allph, phases)
        phases = allph
    sd.profile = False
    eng.sync()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    with eng.stream_context():
        e0.record(eng.stream)
    marks = []
    for _ in range(steps):
        sd.step()
        ev = torch.cuda.Event(enable_timing=True)
        with eng.stream_context():
            ev.record(eng.stream)
        marks.append(ev)
    with eng.stream_context():
        e1.record(eng.stream)
    eng.sync()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    t = torch.tensor([e0.elapsed_time(e1), wall * 1e3], dtype=torch.float64, device=dev)
    owned = torch.tensor([sd.n_owned], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(owned, op=dist.ReduceOp.SUM)
    ms = float(t[0].item())
    per_step = [a.elapsed_time(b) for a, b in zip([e0] + marks[:-1], marks)]
    # aggregate invariants of the state after the timed window (north_star: mean PBF density error and kinetic energy must
    # agree between decompositions within 1 %): sums over the OWNED particles of every rank
    sd.check()  # static layout: device-side capacity flags + the current particle count (a host synchronisation, untimed)
    with eng.stream_context():
        if sd.static:
            # own region [0, S) at the last sort: rows that held a particle (the sort's permutation maps sorted row -> row)
            n_loc = sd.S + 2 * sd.ghost_cap
            perm = eng.field("RadixSortIndices", u32=True)[:n_loc].to(torch.int64)
            keys = eng.field("p_cellID", u32=True)[:n_loc]  # sorted cell ids: "no particle" rows carry an id beyond the grid
            own = (perm < sd.S) & (keys >= 0) & (keys < sd.grid[0] * sd.grid[1] * sd.grid[2])
        else:
            n_loc = sd.n_owned + sd.ghost_rows
            own = eng.field("RadixSortIndices", u32=True)[:n_loc] < sd.n_owned  # rows of the last sort that are not ghosts
        dens = eng.field("p_density")[:n_loc][own].double()
        v = eng.vel()[:sd.n_owned, :3].double()
        inv = torch.stack([(dens / 450.0 - 1.0).abs().sum(), 0.5 * (v * v).sum(), torch.tensor(float(sd.n_owned), dtype=torch.float64, device=dev)])
    eng.sync()
    if world > 1:
        dist.all_reduce(inv, op=dist.ReduceOp.SUM)
    invariants = {"mean_abs_density_error": float(inv[0] / inv[2]), "kinetic_energy": float(inv[1]), "after_step": warmup + 1 + steps}
    bpp = WORKLOADS["pbf_dam_16m_I3_vorticity_xsph"][1]
    peak, _ = measured_peak()
    out = {"workload": "pbf_dam_16m_I3_vorticity_xsph, x-slab decomposition (2 ghost layers, 2I+3 halo refreshes + migration per step)",
           "particles": total, "particles_owned_sum": int(owned.item()), "n_gpus": world, "steps": steps, "warmup": warmup,
           "scaling": "strong", "ms_per_step": ms / steps, "steps_per_s": steps / (ms * 1e-3),
           "value": total * steps / (ms * 1e-3), "unit": "particle-updates/s", "wall_ms_per_step": float(t[1].item()) / steps,
           "invariants": invariants, "ghost_capacity_rows_per_face": ghost_cap,
           "per_step_ms_rank0": [round(x, 2) for x in per_step], "per_rank": {k: v for k, v in sd.stats.items() if k != "phases"}, "phases_ms_per_rank": phases, "algorithmic_bytes_per_particle": bpp,
           "whole_step_frac_of_hbm_per_gpu": round(total * steps / (ms * 1e-3) * bpp / 1e9 / peak / world, 4)}
    eng.h.close()
    return out


def run_slab_group_library(world, one_gpu):
    """`rtp_headless slabs <world>`: the 16.7M-particle dam on `world` slabs / GPUs driven by rtp_slab_group (C ABI, one host
    thread, no Python, no NCCL), same step window (steps 10-19); ms_per_step is HOST wall clock (enqueue + drain)."""
    import subprocess
    exe = os.path.join(ROOT, "realtimeparticles_b200", "lib", "rtp_headless")
    try:
        r = subprocess.run([exe, "slabs", str(world), "10", "9"], capture_output=True, text=True, timeout=240)
        d = json.loads(r.stdout.strip().splitlines()[-1])
    except Exception as e:
        return {"error": repr(e)[:300]}
    if one_gpu:
        d["strong_scaling_efficiency"] = round(one_gpu["ms_per_step"] / (world * d["ms_per_step"]), 4)
        ke = one_gpu["invariants"]["kinetic_energy"]
        d["kinetic_energy_vs_one_gpu"] = abs(d["kinetic_energy"] - ke) / ke
    return d


def run_ours(args):
    import numpy as np
    import torch
    from realtimeparticles_b200 import _abi as abi, models

    rank, local_rank, world = dist_env()
    if not torch.cuda.is_available() or abi.lib().rtp_device_count() == 0:
        raise SystemExit("bench.py: no CUDA device -- the CUDA backend has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("TORCH_NCCL_HIGH_PRIORITY", "1")  # the slab refreshes travel while sweeps fill the SMs
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    h, pos0 = make_pbf(abi, local_rank)
    stream = torch.cuda.ExternalStream(h.stream(), device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    K, W = args.steps, max(args.warmup, 3)
    flags = abi.STEP_PHYSICS

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (also builds the CUDA graph of one step)
    with torch.cuda.stream(stream):
        for _ in range(W):
            h.step_n(1, flags)
    h.sync()
    launches_per_step = h.last_launch_count()

    # ---- timed region: K steps, L2 flushed before each, per-step events on the launching stream
    sampler = ClockSampler(local_rank)
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    stops = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    barrier()
    sampler.start()
    t_wall0 = time.perf_counter()
    with torch.cuda.stream(stream):
        for i in range(K):
            flush.zero_()
            starts[i].record(stream)
            h.step_n(1, flags)
            stops[i].record(stream)
    h.sync()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    sampler.stop_flag = True
    sampler.join()
    ms = sum(s.elapsed_time(e) for s, e in zip(starts, stops))
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    value = world * N130K * K / (ms_total * 1e-3)

    # ---- same K steps back to back without the flush (state L2-resident: the regime the app runs in)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        h.step_n(K, flags)
        e1.record(stream)
    h.sync()
    warm_ms = e0.elapsed_time(e1)

    slab = None
    if world > 1 and not args.no_slab:
        h.close()
        h = None
        del flush
        torch.cuda.empty_cache()
        try:
            slab = run_slab_16m(args, torch, abi, rank, local_rank, world)
            # the 1-GPU anchor of the strong-scaling efficiency: the SAME code path and step window with one slab, on rank 0
            # (the other ranks wait), so that the efficiency is formed from like-for-like numbers of one box
            dist.barrier()
            if rank == 0:
                one = run_slab_16m(args, torch, abi, 0, local_rank, 1)
                slab["one_gpu_same_window"] = {k: one[k] for k in ("ms_per_step", "value", "invariants")}
                slab["strong_scaling_efficiency"] = round(one["ms_per_step"] / (world * slab["ms_per_step"]), 4)
            dist.barrier()
            # the same decomposition driven from inside the library by ONE host thread (rtp_slab_group, csrc/slab_group.cu:
            # slabs pull their neighbours' rows over peer memory) through the C++ headless harness, on all GPUs of the box;
            # the other ranks wait on a CPU-side (gloo) barrier so that no NCCL kernel spins on their GPUs meanwhile
            torch.cuda.empty_cache()
            cpu_group = dist.new_group(backend="gloo")
            dist.barrier(group=cpu_group)
            if rank == 0:
                slab["library_driver"] = run_slab_group_library(world, slab.get("one_gpu_same_window"))
            dist.barrier(group=cpu_group)
        except Exception as e:  # keep the headline line even if the large config cannot run here
            slab = dict(slab or {}, error=repr(e)[:300])
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        h, pos0 = make_pbf(abi, local_rank)
        stream = torch.cuda.ExternalStream(h.stream(), device=dev)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- per-kernel durations (events between launches, L2 flushed before each step)
    h.enable_profiling(True)
    acc, cnt, reps = {}, {}, 20
    for _ in range(reps):
        with torch.cuda.stream(stream):
            flush.zero_()
        h.step(flags)
        for name, v in h.stage_times():
            acc[name] = acc.get(name, 0.0) + v
            cnt[name] = cnt.get(name, 0) + 1
    h.enable_profiling(False)
    abytes = algorithmic_bytes(27000, N130K)
    peak, peak_src = measured_peak()
    kernels = {}
    step_ms = sum(acc.values()) / reps
    for name in acc:
        per_launch_ms = acc[name] / cnt[name]
        b = abytes.get(name)
        gbs = (b * N130K / (per_launch_ms * 1e-3) / 1e9) if b else None
        kernels[name] = {"ms_per_launch": round(per_launch_ms, 5), "launches_per_step": cnt[name] // reps,
                         "share": round(acc[name] / reps / step_ms, 4),
                         "algorithmic_GBps": None if gbs is None else round(gbs, 1),
                         "frac_of_hbm": None if gbs is None else round(gbs / peak, 4)}
    dom = max((k for k in kernels if abytes.get(k)), key=lambda k: kernels[k]["share"])
    kernels_window = ("steps %d-%d of the dam break (after the %d warm-up + %d timed + %d L2-resident steps), events between "
                      "UN-GRAPHED launches with the L2 flushed before each step: their sum exceeds ms_per_step, which replays "
                      "one CUDA graph per step over steps %d-%d" % (W + 2 * K + 1, W + 2 * K + reps, W, K, K, W + 1, W + K))
    roofline = {"bound": "hbm", "kernel": dom, "achieved": kernels[dom]["algorithmic_GBps"], "peak": peak, "unit": "GB/s",
                "frac": kernels[dom]["frac_of_hbm"], "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": abytes[dom] * N130K,
                "whole_step": {"algorithmic_bytes_per_particle": PBF_BYTES_PER_PARTICLE_STEP,
                               "achieved": round(value / world * PBF_BYTES_PER_PARTICLE_STEP / 1e9, 1),
                               "frac": round(value / world * PBF_BYTES_PER_PARTICLE_STEP / 1e9 / peak, 4)},
                "note": "neighbour sweeps test ~460 candidate pairs per particle on 16-64 B of compulsory traffic: instruction-"
                        "issue bound (ncu: 65-71% of the issue slots busy, profiles/r02b_ncu_pbf130k_step.txt), far below the HBM line "
                        "by construction (SURVEY 8d); traffic = ncu dram bytes per launch with a cold L2, dominated by the "
                        "neighbour lists a step writes once (tile filter + build) and its seven later sweeps walk instead of "
                        "re-testing the candidates (DESIGN.md section 5)"}
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tj = json.load(f)
        roofline["traffic"] = tj.get(dom)
        # what actually bounds the sweeps: warp-instruction issue (ncu smsp__issue_active, committed capture)
        roofline["issue_active_pct_ncu"] = tj.get("issue_active_pct")
        roofline["traffic_source"] = tj.get("source")
    except Exception:
        pass

    # ---- e2e through the public API with host buffers
    params = models.ModelParams(currNbParticles=N130K, maxNbParticles=N130K, boxSize=(10, 10, 10), gridRes=(30, 30, 30),
                                pCase=models.PhysicsCase.FLUIDS_DAM, device=local_rank)
    m = models.CreateModel(models.ModelType.FLUIDS, params)
    js = m.getInputJson()
    js["Fluids"]["Nb Jacobi Iterations"][0] = JACOBI
    m.updateInputJson(js)
    m.setStepFlags(abi.STEP_PHYSICS)
    hpos = torch.from_numpy(pos0.copy()).pin_memory()
    hvel = torch.zeros((N130K, 4), dtype=torch.float32).pin_memory()
    ke = max(20, min(K, 500))
    hpos_np, hvel_np = hpos.numpy(), hvel.numpy()

    def e2e_step():
        # the state lives on the HOST between steps: what step k downloads is what step k + 1 uploads, so the simulation
        # advances (the dam collapses over the timed window like in the device-resident run)
        # (page-locked buffers: the four copies are stream-ordered around the step, one synchronisation per frame)
        m.upload("p_pos", hpos_np, blocking=False)
        m.upload("p_vel", hvel_np, blocking=False)
        m.update()
        m.download("p_pos", out=hpos_np, blocking=False)
        m.download("p_vel", out=hvel_np, blocking=False)
        m.sync()
    for _ in range(3):
        e2e_step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(ke):
        e2e_step()
    torch.cuda.synchronize()
    e2e_dt = time.perf_counter() - t0
    e2e_moved = float(np.abs(hpos.numpy()[:, :3] - pos0[:, :3]).max())
    e2e = {"value": N130K * ke / e2e_dt, "unit": "particle-updates/s", "h2d_bytes_per_step": 2 * N130K * 16,
           "d2h_bytes_per_step": 2 * N130K * 16, "steps": ke, "ms_per_step": round(1e3 * e2e_dt / ke, 4),
           "max_displacement_over_run": round(e2e_moved, 4),
           "api": "realtimeparticles_b200.models.Fluids: upload(p_pos,p_vel) + update() + download(p_pos,p_vel) + sync(), pinned host "
                  "buffers, stream-ordered copies (rtp_upload_async / rtp_download_async), one synchronisation per frame; the "
                  "downloaded state is the next step's upload"}

    # ---- what Model::update() costs in the app: physics + render-side kernels + camera sort (Fluids.cpp:400-471)
    hf, _ = make_pbf(abi, local_rank)
    sf = torch.cuda.ExternalStream(hf.stream(), device=dev)
    full_flags = abi.STEP_PHYSICS | abi.STEP_RENDER_AUX | abi.STEP_CAMERA_SORT
    with torch.cuda.stream(sf):
        hf.step_n(W, full_flags)
    hf.sync()
    evf = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(min(K, 500))]
    with torch.cuda.stream(sf):
        for a, b in evf:
            flush.zero_()
            a.record(sf)
            hf.step_n(1, full_flags)
            b.record(sf)
    hf.sync()
    full_ms = sum(a.elapsed_time(b) for a, b in evf)
    full_update = {"steps_per_s": round(len(evf) / (full_ms * 1e-3), 1), "ms_per_step": round(full_ms / len(evf), 4), "steps": len(evf),
                   "launches_per_step": hf.last_launch_count(),
                   "what": "rtp_step(PHYSICS | RENDER_AUX | CAMERA_SORT): the whole Fluids::update() of the app, L2 flushed per step"}
    hf.close()

    # ---- the same step SUSTAINED: the driver's short window (steps W+1 .. W+K) covers the start of the dam break, when the
    # lattice is barely disturbed and the lists are at their shortest; this is the figure over 2000 more steps of the collapse
    sustained = None
    if K < 1000:
        hs, _ = make_pbf(abi, local_rank)
        ss = torch.cuda.ExternalStream(hs.stream(), device=dev)
        with torch.cuda.stream(ss):
            hs.step_n(50, flags)
        hs.sync()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(2000)]
        with torch.cuda.stream(ss):
            for a, b in evs:
                flush.zero_()
                a.record(ss)
                hs.step_n(1, flags)
                b.record(ss)
        hs.sync()
        sus_ms = sum(a.elapsed_time(b) for a, b in evs)
        sustained = {"steps": len(evs), "window": "steps 51-2050 of the dam break, L2 flushed before every step", "ms_per_step": round(sus_ms / len(evs), 4),
                     "steps_per_s": round(len(evs) / (sus_ms * 1e-3), 1), "value": N130K * len(evs) / (sus_ms * 1e-3), "unit": "particle-updates/s"}
        hs.close()

    # ---- CPU baseline on this box's host cores, bounded sample: the reference's own kernels compiled for the CPU
    # (oracle/_ref) when that library travelled here, and the oracle port beside it
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        for kind, (nsteps, budget) in (("reference", (40, 14.0)), ("port", (20, 8.0))):
            cw = cpu_world(pos0, kind)
            if cw is None:
                continue
            n, dt = time_cpu_world(cw[0], 1, nsteps, budget)
            r = {"value": N130K * n / dt, "unit": "particle-updates/s", "cores": cw[1], "kind": kind,
                 "sample": "%d full steps of the same 131072-particle PBF workload after 1 warm-up (%.1f s)" % (n, dt)}
            if cpu is None:
                cpu = r
            else:
                cpu["oracle_port"] = {k: r[k] for k in ("value", "cores", "sample")}
        if cpu is not None:
            try:
                cpu["parity"] = parity_figures(abi, pos0, local_rank)
            except Exception as e:
                cpu["parity"] = {"error": repr(e)[:200]}
            try:
                cpu["reference_opencl_on_this_gpu"] = reference_opencl_figures(pos0)
            except Exception as e:
                cpu["reference_opencl_on_this_gpu"] = {"unavailable": repr(e)[:200]}

    others = {}
    if not args.no_other_workloads:
        for name, (k2, w2) in (("boids_130k", (300, 20)), ("clouds_130k_I2", (300, 20)), ("pbf_dam_16m_I3_vorticity_xsph", (12, 3))):
            try:
                others[name] = quick_bench(abi, torch, name, local_rank, k2, w2, flush)
            except Exception as e:  # e.g. not enough free memory for the 16M config
                others[name] = {"error": str(e)[:200]}

    out = {
        "metric": "particle-updates/sec", "value": value, "unit": "particle-updates/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": headline_config(world),
        "steps_per_s": K / (ms_total * 1e-3),
        "l2_resident": {"value": N130K * K / (warm_ms * 1e-3), "steps_per_s": K / (warm_ms * 1e-3),
                        "note": "same K steps replayed back to back from one CUDA graph, no flush"},
        "e2e": e2e, "full_update": full_update, "gpu_launches": launches_per_step * K * world, "launches_per_step": launches_per_step,
        "roofline": roofline, "kernels": kernels, "kernels_window": kernels_window, "cpu_baseline": cpu, "clocks": sampler.result(),
        "sustained": sustained, "other_workloads": others, "slab_16m": slab,
        "slab_16m_strong_scaling_efficiency": (slab or {}).get("strong_scaling_efficiency"),
        "slab_16m_library_driver_strong_scaling_efficiency": ((slab or {}).get("library_driver") or {}).get("strong_scaling_efficiency"),
        "wall_s_timed_region": round(t_wall, 3),
    }
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-workloads", action="store_true", help="skip the quick boids / clouds / 16M numbers")
    ap.add_argument("--no-slab", action="store_true", help="N > 1: skip the slab-decomposed 16M-particle run")
    ap.add_argument("--slab-only", action="store_true", help="run only the slab-decomposed 16M-particle config (any N)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.slab_only:
        import torch
        from realtimeparticles_b200 import _abi as abi
        rank, local_rank, world = dist_env()
        torch.cuda.set_device(local_rank)
        if world > 1:
            import torch.distributed as dist
            os.environ.setdefault("TORCH_NCCL_HIGH_PRIORITY", "1")  # the slab refreshes travel while sweeps fill the SMs
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        out = run_slab_16m(args, torch, abi, rank, local_rank, world, steps=max(args.steps, 1) if args.steps < 100 else 10)
        if rank == 0:
            print(json.dumps(out), flush=True)
        if world > 1:
            dist.destroy_process_group()
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
