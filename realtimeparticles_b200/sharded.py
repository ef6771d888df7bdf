"""x-slab domain decomposition of the PBF step across the GPUs of one box (new design, the reference is
single-device).

Rank r owns the x-layers [xlo, xhi) of the GLOBAL grid. `cell1D` is x-major (grid.cl:33-35), so the rank's particles
are a contiguous range of the global cell-sorted order and every rank keeps global cell ids: the single-GPU kernels run
unchanged on "owned particles + ghost copies of the neighbours' two boundary layers".

Per step (one process per GPU, neighbour-only point-to-point exchanges, no collective on the data path):
  1. predict + cell ids of the owned particles; particles whose predicted cell left the slab MIGRATE to the neighbour
     (pos, vel), which predicts them again (same arithmetic, same bits);
  2. HALO: predicted positions of the owned particles in the two layers next to each slab face go to that neighbour
     and are appended as ghosts; everything is sorted by cell together (ghost layer 2 covers the reference's
     stale-grid look-ups: a particle's centre cell is recomputed from its moved position, fluids.cl:88-89);
  3. every stage whose output a neighbour sweep reads is followed by a ghost REFRESH of exactly that field:
     lambda after density+lambda, predPos (and vel on the last iteration) after correction, |vorticity| after the
     vorticity sweep, the confined velocity before XSPH: 2 I + 3 exchanges per step;
  4. ghosts are dropped, the owned particles stay in cell-sorted order for the next step.
Ghost rows are never computed locally: the sweeps skip them (csrc/sweep.cuh isGhostRow) and the refreshes bring the
owner's values before anybody reads them. Results equal the single-GPU run up to the order of particles inside a cell
(migrated particles are appended), i.e. up to fp32 summation order; cell ids and the cell table are identical.

What keeps the host out of the step (CUDA engine; `static_layout`):
  * the row layout of a rank never changes: rows [0, S) are the rank's own region -- its particles, cell-sorted, at the
    front, "no particle" rows behind them, the last 2 x migrate_cap rows reserved as arrival slots --, followed by one ghost
    region of ghost_cap rows per slab face. Every launch covers the same rows every step, so NOTHING has to come back to
    the host inside a step: the Python driver only enqueues, it runs ahead of the GPU, and its cost is hidden. Leavers are
    marked "no particle" on the device, arrivals land in the arrival slots, the end-of-step compaction moves the particles
    to the front again. Capacity violations raise a device flag that is read every `check_every` steps.
  * ghost regions and exchange buffers have a FIXED capacity: a halo / refresh message is always `ghost_cap` rows, padded
    with "no particle" rows (+inf positions: their cell key is beyond the grid, so they sort behind every real particle,
    appear in no cell range and are skipped by the sweeps like any ghost). No count has to reach the host before the next
    launch; the counts of a step are checked against the capacity one step later, together with the one read-back below;
  * (engines without static_layout -- the oracle in the CPU tests -- read the migration counts back once per step: the
    number of owned particles sizes their launches);
  * gather / scatter of exchange rows, the inverse permutation and the ghost displacement check are kernels of the
    library behind the C ABI (rtp_shard_pack / unpack / inverse_perm / check_ghosts), on the handle's stream.
The transport is pluggable: torch.distributed P2P batches (NCCL over NVLink on GPUs, gloo on CPU) when every rank is a
process, or LocalSlabGroup, which runs several ranks in ONE process (tests: 2 and 4 slabs on one device; C-side hosts can
do the same with the stage API). The engine is duck-typed so that the CPU tests drive the same code over the oracle.
"""
import os

import torch
import torch.distributed as dist

from . import _abi

GHOST_LAYERS = 2


class _CudaArray:
    """zero-copy torch view of library device memory (__cuda_array_interface__)"""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"data": (ptr, False), "shape": tuple(shape), "typestr": typestr, "version": 3,
                                         "strides": None}


class CudaSlabEngine:
    """Local engine: one rtp handle (fluids model, GLOBAL box/grid) + torch views of the buffers the exchanges touch."""

    ST = dict(PREDICT=0, GHOST_KEYS=1, SORT=2, DENSITY_LAMBDA=3, CORRECTION=4, VORTICITY=5, CONFINEMENT=6, XSPH=7, DROP_GHOSTS=8,
              PREDICT_FROM=9)
    BUF = dict(KEYS_IN=0, PRED_IN=1, PRED_CUR=2, LAMBDA=3, VEL_SORTED=4, VORT_NORM=5, VEL_CONFINED=6, LIST_BUILD_POS=7,
               LIST_INVALID=8, POS=9, VEL=10, ROW_BOUNDS=11)
    ROW = dict(PRED_IN=4, PRED_CUR=4, LAMBDA=1, VEL_SORTED=4, VORT_NORM=1, VEL_CONFINED=4, POS=4, VEL=4)

    def __init__(self, capacity, box, grid, device, fluid_params=None, jacobi=3):
        self.device = torch.device("cuda", device)
        self.h = _abi.Handle(_abi.FLUIDS, capacity, 0, box, grid, 3, 0, device)
        self.capacity = capacity
        self.jacobi = jacobi
        fp = fluid_params or _abi.FluidParams(450.0, 600.0, 0.010, 3, 1, 0.006, 0.001, 4, 1, 0.0004, 0.0001)
        self.vorticity = bool(fp.isVorticityConfEnabled)
        self.h.set_fluid_params(fp, jacobi)
        self.stream = torch.cuda.ExternalStream(self.h.stream(), device=self.device)

    def stream_context(self):
        return torch.cuda.stream(self.stream)

    def _view(self, ptr, nbytes, f4=False, u32=False):
        if u32:
            return torch.as_tensor(_CudaArray(ptr, (nbytes // 4,), "<i4"), device=self.device)
        if f4:
            return torch.as_tensor(_CudaArray(ptr, (nbytes // 16, 4), "<f4"), device=self.device)
        return torch.as_tensor(_CudaArray(ptr, (nbytes // 4,), "<f4"), device=self.device)

    def buf(self, name):
        p, n = self.h.shard_buffer(self.BUF[name])
        return self._view(p, n, f4=self.ROW.get(name) == 4, u32=name in ("KEYS_IN", "LIST_INVALID", "ROW_BOUNDS"))

    def field(self, name, f4=False, u32=False):
        p = self.h.device_ptr(name)
        return self._view(p, self.h.field_bytes(name), f4=f4, u32=u32)

    # ---- interface used by SlabDecomposition
    def set_counts(self, n_owned, n_local):
        self.h.set_nb_particles(n_local)
        self.h.shard_set_owned(n_owned)

    def pos(self):
        return self.field("p_pos", f4=True)

    def vel(self):
        return self.field("p_vel", f4=True)

    def keys_in(self):
        return self.buf("KEYS_IN")

    def pred_in(self):
        return self.buf("PRED_IN")

    ROWS = dict(ALL=0, BOUNDARY=1, INTERIOR=2)

    def stage(self, name, it=0, last=False, rows="ALL"):
        self.h.shard_stage_rows(self.ST[name], it, last, self.ROWS[rows])

    # ---- overlap of a ghost refresh with the sweeps of the interior rows (include/rtp_cuda.h: rtp_shard_set_interior ...)
    overlap_capable = True

    def set_interior(self, cell_lo, cell_hi, max_boundary_rows):
        self.h.shard_set_interior(cell_lo, cell_hi, max_boundary_rows)
        self.exchange_stream = torch.cuda.ExternalStream(self.h.shard_exchange_stream(), device=self.device)

    def exchange_fork(self):
        self.h.shard_exchange_fork()

    def exchange_done(self):
        self.h.shard_exchange_done()

    def exchange_join(self):
        self.h.shard_exchange_join()

    def refresh_buffers(self, name, last=False):
        """internal buffers (cell-sorted index space) whose ghost rows must be refreshed after stage `name`"""
        if name == "DENSITY_LAMBDA":
            return ["LAMBDA"]
        if name == "CORRECTION":
            return ["PRED_CUR"] + (["VEL_SORTED"] if last else [])
        if name == "VORTICITY":
            return ["VORT_NORM"]
        if name == "CONFINEMENT":
            return ["VEL_CONFINED"]
        return []

    def row_width(self, name):
        return self.ROW[name]

    def pack(self, name, idx, out):
        """out[k] = buffer[idx[k]] (idx int32 device tensor, -1 = "no particle" padding)"""
        self.h.shard_pack(self.BUF[name], idx.data_ptr(), idx.numel(), out.data_ptr())

    def unpack(self, name, idx, src):
        self.h.shard_unpack(self.BUF[name], idx.data_ptr(), idx.numel(), src.data_ptr())

    def inverse_perm(self, out):
        self.h.shard_inverse_perm(out.data_ptr())

    def check_ghosts(self, idx, next_epoch):
        self.h.shard_check_ghosts(idx.data_ptr(), idx.numel(), next_epoch)

    def new(self, shape, dtype):
        return torch.empty(shape, dtype=dtype, device=self.device)

    static_layout = True  # the sweeps of a sharded handle skip "no particle" rows (+inf position): see SlabDecomposition

    def classify(self, n, layer_below, layer_from, below, above):
        """below[i] / above[i] (bool tensors): row i < n holds a particle whose predicted cell x-layer is < layer_below /
        >= layer_from (one kernel over the cell ids and p_pos)"""
        self.h.shard_classify(n, layer_below, layer_from, below.data_ptr(), above.data_ptr())

    def clear_rows(self, idx):
        """the rows idx (int32, -1 = padding) hold no particle any more"""
        self.h.shard_clear_rows(idx.data_ptr(), idx.numel())

    def sync(self):
        self.h.sync()


class _Exchange:
    """one neighbour exchange: per side a list of (send tensor, receive tensor) pairs; None for a missing neighbour"""

    def __init__(self, left, right, overlapped=False):
        self.left, self.right = left, right
        self.overlapped = overlapped  # the transport runs on the engine's exchange stream (between exchange_fork and _done)


class SlabDecomposition:
    """One rank of the x-slab decomposition. `engine` is a CudaSlabEngine (or any object with the same interface)."""

    def __init__(self, engine, grid, rank=None, world=None, group=None, ghost_cap=None, migrate_cap=None, overlap=None):
        self.e = engine
        self.grid = tuple(grid)
        self.group = group
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world = dist.get_world_size(group) if world is None else world
        rx = self.grid[0]
        self.xlo = (rx * self.rank) // self.world
        self.xhi = (rx * (self.rank + 1)) // self.world
        if self.world > 1 and (self.xhi - self.xlo) < 2 * GHOST_LAYERS:
            raise ValueError("slab thinner than 2 x ghost layers")
        self.n_owned = 0
        self.stats = {}
        self.profile = False  # record per-phase device times (ms) of the last step into stats["phases"]
        self._marks = []
        # fixed capacities (rows) of a ghost region / a migration message per side
        self.left = self.rank - 1 if self.rank > 0 else None
        self.right = self.rank + 1 if self.rank < self.world - 1 else None
        sides = (self.left is not None) + (self.right is not None)
        self.ghost_cap = int(ghost_cap) if ghost_cap is not None else (engine.capacity // 8 if sides else 0)
        self.migrate_cap = int(migrate_cap) if migrate_cap is not None else max(self.ghost_cap // 4, 1)
        self.ghost_rows = self.ghost_cap * sides
        self._pending_counts = None  # device tensor with the halo counts of the previous step (checked one step late)
        self.static = bool(getattr(engine, "static_layout", False)) and sides > 0
        self.check_every = 16  # static layout: how often the device-side capacity flag is read back
        self._steps = 0
        e = engine
        if sides:
            f32, i64 = torch.float32, torch.int64
            self._mig_s = {s: e.new((2, self.migrate_cap, 4), f32) for s in "lr"}
            self._mig_r = {s: e.new((2, self.migrate_cap, 4), f32) for s in "lr"}
            self._cnt_s = {s: e.new((1,), i64) for s in "lr"}
            self._cnt_r = {s: e.new((1,), i64) for s in "lr"}
            self._row_s = {w: {s: e.new((self.ghost_cap, w) if w > 1 else (self.ghost_cap,), f32) for s in "lr"} for w in (1, 4)}
            self._row_r = {w: {s: e.new((self.ghost_cap, w) if w > 1 else (self.ghost_cap,), f32) for s in "lr"} for w in (1, 4)}
            # static layout: [field of a refresh][row width] -> (2 faces, ghost_cap rows): both faces in one pack / unpack
            self._rows2_s = [{w: e.new((2, self.ghost_cap, w) if w > 1 else (2, self.ghost_cap), f32) for w in ws} for ws in ((1, 4), (4,))]
            self._rows2_r = [{w: e.new((2, self.ghost_cap, w) if w > 1 else (2, self.ghost_cap), f32) for w in ws} for ws in ((1, 4), (4,))]
            for d in self._rows2_r:
                for t in d.values():
                    t.zero_()
            self._inv = e.new((e.capacity,), torch.int32)
            if self.static:
                # rows [0, S): own region (arrival slots at its end); [S, S + 2 * ghost_cap): ghost regions (left, right)
                self.S = e.capacity - 2 * self.ghost_cap
                self.A0 = self.S - 2 * self.migrate_cap
                if self.A0 <= 0:
                    raise ValueError("slab capacity too small for the ghost regions and arrival slots")
                self._err = torch.zeros(1, dtype=torch.int64, device=e.device)
                self._mask_l = torch.zeros(self.S, dtype=torch.bool, device=e.device)
                self._mask_r = torch.zeros(self.S, dtype=torch.bool, device=e.device)
                self._migrated = torch.zeros(1, dtype=torch.int64, device=e.device)
                self._inf_rows = torch.zeros((max(self.ghost_cap, self.migrate_cap), 4), device=e.device)
                self._inf_rows[:, :3] = float("inf")
        # static layout: the refresh after a stage travels while the next stage sweeps the interior rows (x-layers two or
        # more from both faces: no ghost among their neighbours, nobody's ghost); RTP_SLAB_OVERLAP=0 / overlap=False: in order
        if overlap is None:
            overlap = os.environ.get("RTP_SLAB_OVERLAP", "1") != "0"
        self.overlap = bool(overlap) and self.static and bool(getattr(engine, "overlap_capable", False))
        if self.overlap:
            plane = self.grid[1] * self.grid[2]
            lo = self.xlo + (GHOST_LAYERS if self.left is not None else 0)
            hi = self.xhi - (GHOST_LAYERS if self.right is not None else 0)
            # rows outside the interior: per face at most ghost_cap ghosts and ghost_cap owned rows of the two face layers
            engine.set_interior(lo * plane, max(lo, hi) * plane, 2 * self.ghost_cap * sides)

    # ---- initial distribution: every rank is given the full initial state and keeps the particles of its slab
    def slab_of(self, keys):
        layer = torch.div(keys, self.grid[1] * self.grid[2], rounding_mode="floor")  # (keys are non-negative int32)
        return layer.clamp_(max=self.grid[0] - 1)  # index RES (particle exactly on the +x wall) belongs to the last layer

    def load_owned(self, pos, vel):
        """pos/vel: (n, 4) float32 tensors on the engine's device holding ONLY this rank's particles, any order"""
        n = pos.shape[0]
        if self.static:
            if n > self.A0:
                raise RuntimeError("slab capacity exceeded: %d particles > %d rows before the arrival slots" % (n, self.A0))
            with self.e.stream_context():
                p, v = self.e.pos(), self.e.vel()
                p[:, :3] = float("inf")  # "no particle" everywhere ...
                p[:, 3] = 0.0
                v.zero_()
                p[:n] = pos  # ... except the rank's particles at the front
                v[:n] = vel
                self.e.set_counts(self.S, self.S)
            self.n_owned = n
            return
        if n + self.ghost_rows > self.e.capacity:
            raise RuntimeError("slab capacity exceeded: %d owned + %d ghost rows > %d" % (n, self.ghost_rows, self.e.capacity))
        with self.e.stream_context():
            self.e.set_counts(n, n)
            self.e.pos()[:n] = pos
            self.e.vel()[:n] = vel
        self.n_owned = n

    def _global(self, r):
        return r if self.group is None else dist.get_global_rank(self.group, r)

    def _sides(self):
        return [(s, p) for s, p in (("l", self.left), ("r", self.right)) if p is not None]

    # ---- transport over torch.distributed: one batched P2P per exchange point
    def _run_dist(self, x):
        ops = []
        for peer, pairs in ((self.left, x.left), (self.right, x.right)):
            if peer is None or pairs is None:
                continue
            for send, recv in pairs:
                ops.append(dist.P2POp(dist.isend, send, self._global(peer), self.group))
                ops.append(dist.P2POp(dist.irecv, recv, self._global(peer), self.group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()

    def _apply_migration(self, pos, vel, n, holes, arrivals):
        """arrivals (k_in, 8) fill the slots `holes` of the leavers; a surplus is appended, a deficit is filled from the tail"""
        k_out, k_in = holes.numel(), arrivals.shape[0]
        m = min(k_out, k_in)
        if m:
            pos[holes[:m]] = arrivals[:m, :4]
            vel[holes[:m]] = arrivals[:m, 4:]
        if k_in > k_out:
            extra = k_in - k_out
            if n + extra + self.ghost_rows > pos.shape[0]:
                raise RuntimeError("slab capacity exceeded: %d > %d" % (n + extra + self.ghost_rows, pos.shape[0]))
            pos[n:n + extra] = arrivals[m:, :4]
            vel[n:n + extra] = arrivals[m:, 4:]
            n += extra
        elif k_out > k_in:
            rest = holes[m:]
            n_new = n - rest.numel()
            tail = torch.arange(n_new, n, device=rest.device)
            movers = tail[~torch.isin(tail, rest)]
            fill = rest[rest < n_new]
            if fill.numel():
                pos[fill] = pos[movers]
                vel[fill] = vel[movers]
            n = n_new
        return n

    def warm_up_code_paths(self):
        """Run the migration bookkeeping once on scratch tensors of realistic size so that no lazily loaded torch kernel,
        autotuned launch or allocator growth lands inside a timed step (the first use of a torch op costs milliseconds;
        the first migrating step of the 16M dam used to be 4 ms slower than its neighbours)."""
        dev = self.e.pos().device
        rows = 8192 + self.ghost_rows
        with self.e.stream_context():
            for k_out, k_in in (() if self.static else ((700, 100), (100, 700), (300, 300), (0, 5), (5, 0))):
                p = torch.zeros((rows, 4), device=dev)
                v = torch.zeros((rows, 4), device=dev)
                idx = torch.nonzero_static(torch.arange(4096, device=dev) % 5 == 0, size=1024, fill_value=-1).flatten().to(torch.int32)
                holes = torch.cat([idx[:k_out // 2], idx[400:400 + k_out - k_out // 2]]).to(torch.int64)
                arr = torch.cat([torch.cat([torch.ones((k_in // 2, 4), device=dev)] * 2, dim=1),
                                 torch.cat([torch.ones((k_in - k_in // 2, 4), device=dev)] * 2, dim=1)], dim=0)
                self._apply_migration(p, v, 4000, holes, arr)
            lay = self.slab_of(torch.arange(4096, device=dev, dtype=torch.int32))
            self._compact(lay < 1, 64)
            torch.cat([torch.ones(1, device=dev, dtype=torch.int64)] * 3).cpu().tolist()
        self.e.sync()

    def _mark(self, name):
        if self.profile and hasattr(self.e, "stream"):
            ev = torch.cuda.Event(enable_timing=True)
            ev.record(self.e.stream)
            self._marks.append((name, ev))

    def step(self):
        """one step over torch.distributed (every rank is a process)"""
        for x in self.step_gen():
            # the P2P batch is ordered after the pack kernels on the engine's stream (the exchange stream for a refresh
            # that overlaps the next sweep)
            with (torch.cuda.stream(self.e.exchange_stream) if x.overlapped else self.e.stream_context()):
                self._run_dist(x)

    def step_gen(self):
        """the step as a generator: yields an _Exchange at every point where rows travel between neighbours (the caller
        performs it: SlabDecomposition.step over torch.distributed, LocalSlabGroup inside one process). The code between
        two exchange points runs with the engine's stream current; the generator is never suspended inside that context."""
        e = self.e
        self._marks = []
        inner = self._step_static() if self.static else self._step()
        while True:
            with e.stream_context():
                try:
                    x = next(inner)
                except StopIteration:
                    break
            yield x
        if self.profile and self._marks:
            e.sync()
            ph = {}
            for (_, a), (name, b) in zip(self._marks[:-1], self._marks[1:]):
                ph[name] = ph.get(name, 0.0) + a.elapsed_time(b)
            self.stats["phases"] = {k: round(v, 3) for k, v in ph.items()}

    def _face_masks(self, layer_below, layer_from):
        """static layout: {"l": rows of the own region holding a particle whose predicted x-layer is < layer_below,
        "r": ... >= layer_from}"""
        e, S = self.e, self.S
        if hasattr(e, "classify"):
            e.classify(S, layer_below, layer_from, self._mask_l, self._mask_r)
            return {"l": self._mask_l, "r": self._mask_r}
        layer = self.slab_of(e.keys_in()[:S])
        alive = torch.isfinite(e.pos()[:S, 0])
        return {"l": alive & (layer < layer_below), "r": alive & (layer >= layer_from)}

    def _compact(self, mask, cap):
        """indices of the set entries of `mask`, padded with -1 to `cap` rows (no host synchronisation), and their number"""
        return torch.nonzero_static(mask, size=cap, fill_value=-1).flatten().to(torch.int32), mask.sum()

    def _step(self):
        e, n = self.e, self.n_owned
        jacobi = e.jacobi
        sides = self._sides()
        pos, vel = e.pos(), e.vel()
        self._mark("begin")

        # 1. predict, migrate
        e.set_counts(n, n)
        e.stage("PREDICT")
        if sides:
            layer = self.slab_of(e.keys_in()[:n])
            idx = {}
            for s, _ in sides:
                idx[s], cnt = self._compact((layer < self.xlo) if s == "l" else (layer >= self.xhi), self.migrate_cap)
                self._cnt_s[s].copy_(cnt.reshape(1))
                e.pack("POS", idx[s], self._mig_s[s][0])
                e.pack("VEL", idx[s], self._mig_s[s][1])
            yield _Exchange(*[[(self._mig_s[s], self._mig_r[s]), (self._cnt_s[s], self._cnt_r[s])] if p is not None else None
                              for s, p in (("l", self.left), ("r", self.right))])
            # the one host synchronisation of the step: how many particles left and arrived (+ last step's halo counts)
            vals = [self._cnt_s[s] for s, _ in sides] + [self._cnt_r[s] for s, _ in sides]
            if self._pending_counts is not None:
                vals.append(self._pending_counts)
            host = torch.cat([v.reshape(-1) for v in vals]).cpu().tolist()
            k = len(sides)
            out_c, in_c, halo_c = host[:k], host[k:2 * k], host[2 * k:]
            if max(out_c + in_c) > self.migrate_cap:
                raise RuntimeError("migration capacity exceeded: %d > %d rows" % (max(out_c + in_c), self.migrate_cap))
            if halo_c and max(halo_c) > self.ghost_cap:
                raise RuntimeError("ghost capacity exceeded in the previous step: %d > %d rows" % (max(halo_c), self.ghost_cap))
            self.stats["migrated_out"] = int(sum(out_c))
            if sum(out_c) + sum(in_c):
                holes = torch.cat([idx[s][:c] for (s, _), c in zip(sides, out_c)]).to(torch.int64)
                arrivals = torch.cat([torch.cat([self._mig_r[s][0][:c], self._mig_r[s][1][:c]], dim=1) for (s, _), c in zip(sides, in_c)], dim=0)
                # only the handful of migrating rows are touched: arrivals fill the slots of the leavers, a surplus is
                # appended, a deficit is filled from the tail (order inside a cell is not preserved for moved rows)
                n = self._apply_migration(pos, vel, n, holes, arrivals)
                e.set_counts(n, n)
                e.stage("PREDICT")
        self._mark("predict+migrate")

        # 2. halo of predicted positions into the fixed-capacity ghost regions, sort everything by cell
        G = self.ghost_cap
        n_loc = n + self.ghost_rows
        if sides:
            if n_loc > e.capacity:
                raise RuntimeError("slab capacity exceeded: %d > %d" % (n_loc, e.capacity))
            layer = self.slab_of(e.keys_in()[:n])
            src, counts = {}, []
            for s, _ in sides:
                m = (layer < self.xlo + GHOST_LAYERS) if s == "l" else (layer >= self.xhi - GHOST_LAYERS)
                src[s], c = self._compact(m, G)
                counts.append(c.reshape(1))
                e.pack("PRED_IN", src[s], self._row_s[4][s])
            self._pending_counts = torch.cat(counts)
            yield _Exchange(*[[(self._row_s[4][s], self._row_r[4][s])] if p is not None else None for s, p in (("l", self.left), ("r", self.right))])
            pred_in = e.pred_in()
            off = {}
            for k, (s, _) in enumerate(sides):
                off[s] = n + k * G
                pred_in[off[s]:off[s] + G] = self._row_r[4][s]
                pos[off[s]:off[s] + G] = self._row_r[4][s]
            vel[n:n_loc] = 0.0
            e.set_counts(n, n_loc)
            e.stage("GHOST_KEYS")
        self.stats.update(owned=n, ghosts=self.ghost_rows)
        self._mark("halo")
        e.stage("SORT")
        self._mark("sort")

        refresh = None
        if sides:
            inv = self._inv[:n_loc]
            e.inverse_perm(inv)
            send = {s: torch.where(src[s] >= 0, inv[src[s].clamp(min=0).to(torch.int64)], src[s]) for s, _ in sides}
            recv = {s: inv[off[s]:off[s] + G] for s, _ in sides}
            ghost_idx = torch.cat([recv[s] for s, _ in sides])

            def refresh(names):
                for name in names:
                    w = e.row_width(name)
                    for s, _ in sides:
                        e.pack(name, send[s], self._row_s[w][s])
                    yield _Exchange(*[[(self._row_s[w][s], self._row_r[w][s])] if p is not None else None
                                      for s, p in (("l", self.left), ("r", self.right))])
                    for s, _ in sides:
                        e.unpack(name, recv[s], self._row_r[w][s])
        self._mark("index-maps")

        # 3. the solver stages, each followed by the refresh of what it produced
        for it in range(jacobi):
            last = it == jacobi - 1
            e.stage("DENSITY_LAMBDA", it)
            self._mark("compute")
            if refresh:
                yield from refresh(e.refresh_buffers("DENSITY_LAMBDA"))
                self._mark("refresh")
            e.stage("CORRECTION", it, last)
            self._mark("compute")
            if refresh:
                yield from refresh(e.refresh_buffers("CORRECTION", last))
                e.check_ghosts(ghost_idx, it + 1)  # the owner's kernel checks the particles it moves; its ghosts are checked here
                self._mark("refresh")
        if e.vorticity:
            e.stage("VORTICITY", jacobi)
            self._mark("compute")
            if refresh:
                yield from refresh(e.refresh_buffers("VORTICITY"))
                self._mark("refresh")
            e.stage("CONFINEMENT", jacobi)
            self._mark("compute")
            if refresh:
                yield from refresh(e.refresh_buffers("CONFINEMENT"))
                self._mark("refresh")
            e.stage("XSPH", jacobi)
            self._mark("compute")

        # 4. drop the ghosts; owned particles stay in cell-sorted order
        if n_loc > n:
            e.stage("DROP_GHOSTS")
        e.set_counts(n, n)
        self.n_owned = n
        self._mark("compact")
        self._mark("end")

    def _step_static(self):
        """one step on the static row layout (module docstring): no host synchronisation"""
        e = self.e
        jacobi = e.jacobi
        sides = self._sides()
        S, G, Mc, A0 = self.S, self.ghost_cap, self.migrate_cap, self.A0
        pos, vel = e.pos(), e.vel()
        self._mark("begin")

        # 1. predict; particles whose predicted cell left the slab migrate: their row is cleared here, the neighbour gets
        #    them in its arrival slots
        e.set_counts(S, S)
        e.stage("PREDICT")
        masks = self._face_masks(self.xlo, self.xhi)
        self._err += torch.isfinite(pos[A0:S, 0]).any().to(torch.int64)  # the arrival slots must be free
        for s, _ in sides:
            idx, cnt = self._compact(masks[s], Mc)
            self._cnt_s[s].copy_(cnt.reshape(1))
            self._migrated += cnt
            self._err += (cnt > Mc).to(torch.int64) * 2
            e.pack("POS", idx, self._mig_s[s][0])
            e.pack("VEL", idx, self._mig_s[s][1])
            e.clear_rows(idx)
        yield _Exchange(*[[(self._mig_s[s], self._mig_r[s])] if p is not None else None for s, p in (("l", self.left), ("r", self.right))])
        for k, s in enumerate("lr"):
            a = A0 + k * Mc
            if (self.left if s == "l" else self.right) is not None:
                pos[a:a + Mc] = self._mig_r[s][0]  # (the message is padded with "no particle" rows)
                vel[a:a + Mc] = self._mig_r[s][1]
        if hasattr(e, "ST") and "PREDICT_FROM" in e.ST:
            e.stage("PREDICT_FROM", A0)  # the arrivals need their prediction and cell id
        else:
            e.stage("PREDICT")  # (everything again: same arithmetic, same bits for the rest)
        self._mark("predict+migrate")

        # 2. halo of predicted positions into the ghost regions, sort everything by cell
        masks = self._face_masks(self.xlo + GHOST_LAYERS, self.xhi - GHOST_LAYERS)
        src, off = {}, {"l": S, "r": S + G}
        for s, _ in sides:
            src[s], c = self._compact(masks[s], G)
            self._err += (c > G).to(torch.int64) * 4
            e.pack("PRED_IN", src[s], self._row_s[4][s])
        yield _Exchange(*[[(self._row_s[4][s], self._row_r[4][s])] if p is not None else None for s, p in (("l", self.left), ("r", self.right))])
        pred_in = e.pred_in()
        for s in "lr":
            rows = self._row_r[4][s] if (self.left if s == "l" else self.right) is not None else self._inf_rows[:G]
            pred_in[off[s]:off[s] + G] = rows
            pos[off[s]:off[s] + G] = rows
        vel[S:S + 2 * G] = 0.0
        n_loc = S + 2 * G
        e.set_counts(S, n_loc)
        e.stage("GHOST_KEYS")
        self.stats.update(owned=self.n_owned, ghosts=2 * G)
        self._mark("halo")
        e.stage("SORT")
        self._mark("sort")

        inv = self._inv[:n_loc]
        e.inverse_perm(inv)
        send = {s: torch.where(src[s] >= 0, inv[src[s].clamp(min=0).to(torch.int64)], src[s]) for s, _ in sides}
        recv = {s: inv[off[s]:off[s] + G] for s, _ in sides}
        ghost_idx = torch.cat([recv[s] for s, _ in sides])

        # both faces are packed / unpacked by ONE kernel launch each: index lists and buffers of the two sides are contiguous
        none = torch.full((G,), -1, dtype=torch.int32, device=inv.device)
        send_all = torch.cat([send.get("l", none), send.get("r", none)])
        recv_all = torch.cat([recv.get("l", none), recv.get("r", none)])

        def refresh(names, check_epoch=None):
            # all fields a stage produced travel in ONE exchange (one send + one receive per neighbour)
            if self.overlap:
                e.exchange_fork()
            bufs = []
            for k, name in enumerate(names):
                w = e.row_width(name)
                sb, rb = self._rows2_s[k][w], self._rows2_r[k][w]
                e.pack(name, send_all, sb)
                bufs.append((name, sb, rb))
            yield _Exchange(*[[(sb[k2], rb[k2]) for _, sb, rb in bufs] if p is not None else None
                              for k2, p in ((0, self.left), (1, self.right))], overlapped=self.overlap)
            for name, sb, rb in bufs:
                e.unpack(name, recv_all, rb)
            if check_epoch is not None:
                e.check_ghosts(ghost_idx, check_epoch)
            if self.overlap:
                e.exchange_done()
        self._mark("index-maps")

        # 3. the solver stages, each followed by the refresh of what it produced. With overlap, a stage sweeps its interior
        #    rows first -- while the refresh of the previous stage is still travelling -- and the rest once it has arrived.
        stages = []
        for it in range(jacobi):
            last = it == jacobi - 1
            stages.append(("DENSITY_LAMBDA", it, False, e.refresh_buffers("DENSITY_LAMBDA"), None))
            stages.append(("CORRECTION", it, last, e.refresh_buffers("CORRECTION", last), it + 1))
        if e.vorticity:
            stages.append(("VORTICITY", jacobi, False, e.refresh_buffers("VORTICITY"), None))
            stages.append(("CONFINEMENT", jacobi, False, e.refresh_buffers("CONFINEMENT"), None))
            stages.append(("XSPH", jacobi, False, [], None))
        for name, it, last, fields, check_epoch in stages:
            if self.overlap:
                e.stage(name, it, last, rows="INTERIOR")
                e.exchange_join()
                e.stage(name, it, last, rows="BOUNDARY")
            else:
                e.stage(name, it, last)
            self._mark("compute")
            if fields:
                yield from refresh(fields, check_epoch)
                self._mark("refresh")

        # 4. particles to the front of the own region, in cell-sorted order; everything else is "no particle"
        e.stage("DROP_GHOSTS")
        e.set_counts(S, S)
        self._mark("compact")
        self._mark("end")
        self._steps += 1
        if self._steps % self.check_every == 0:
            self.check()

    def check(self):
        """static layout: read the device-side flags back (a host synchronisation); raises if a capacity was exceeded"""
        if not self.static:
            return
        with self.e.stream_context():
            rows_err = self.e.buf("ROW_BOUNDS")[3:4].to(torch.int64) if self.overlap else torch.zeros_like(self._err)
            vals = torch.cat([self._err, self._migrated, torch.isfinite(self.e.pos()[:self.S, 0]).sum().reshape(1), rows_err]).cpu().tolist()
        if vals[3]:
            raise RuntimeError("a sweep launched by row phase was sized too small for its rows (the row bounds of a step ago "
                               "were outrun): results since the last check are invalid")
        if vals[0]:
            raise RuntimeError("slab capacity exceeded on the device (flags %d: 1 = arrival slots in use, 2 = migration message, "
                               "4 = ghost region)" % vals[0])
        self.stats["migrated_out_total"] = int(vals[1])
        self.n_owned = int(vals[2])

    def owned_state(self):
        if self.static:
            self.check()
        n = self.n_owned
        with self.e.stream_context():
            return self.e.pos()[:n].clone(), self.e.vel()[:n].clone()


class LocalSlabGroup:
    """All ranks of a decomposition inside ONE process (any mix of devices, e.g. 4 slabs on one GPU): the ranks' steps are
    generators advanced in lock step; at every exchange point the rows are copied between the ranks' buffers."""

    def __init__(self, slabs):
        self.slabs = list(slabs)

    def step(self):
        gens = [sd.step_gen() for sd in self.slabs]
        while True:
            xs = []
            for g in gens:
                try:
                    xs.append(next(g))
                except StopIteration:
                    xs.append(None)
            if all(x is None for x in xs):
                return
            if any(x is None for x in xs):
                raise RuntimeError("the ranks of a LocalSlabGroup left the step at different exchange points")
            for sd in self.slabs:
                sd.e.sync()  # the rows to send are complete
                if hasattr(sd.e, "device"):
                    torch.cuda.synchronize(sd.e.device)  # (also on the exchange stream)
            for r, x in enumerate(xs):
                if x.right is not None:  # my right-going rows are my right neighbour's rows "from the left", and vice versa
                    for (send, _), (_, recv) in zip(x.right, xs[r + 1].left):
                        recv.copy_(send)
                    for (_, recv), (send, _) in zip(x.right, xs[r + 1].left):
                        recv.copy_(send)
            for sd in self.slabs:
                if hasattr(sd.e, "device"):
                    torch.cuda.synchronize(sd.e.device)


def split_initial_state(pos, box, grid, rank, world):
    """rows of `pos` (numpy (n,4)) whose cell x-layer belongs to `rank` (same cell formula as grid.cl:14-24 on the
    initial positions; any consistent assignment works, the first step migrates by predicted position anyway)"""
    import numpy as np
    h = np.float32(np.float32(box[0]) / np.float32(grid[0]))
    w = np.float32(box[0] / 2.0)
    x = np.floor((np.clip(pos[:, 0], -w, w) + w) / h).astype(np.int64).clip(max=grid[0] - 1)
    lo, hi = (grid[0] * rank) // world, (grid[0] * (rank + 1)) // world
    return np.nonzero((x >= lo) & (x < hi))[0]
