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
Ghost values computed locally (their neighbourhoods are incomplete) are never read before the refresh overwrites them.
Results equal the single-GPU run up to the order of particles inside a cell (migrated particles are appended), i.e. up
to fp32 summation order; cell ids and the cell table are identical.

The orchestration below is plumbing (index bookkeeping + torch.distributed send/recv on tensors that alias the
library's device buffers); all per-particle work is done by the CUDA kernels through the stage API of
include/rtp_cuda.h (rtp_shard_*). The engine is duck-typed so that the CPU tests can drive the same code over gloo.
"""
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

    ST = dict(PREDICT=0, GHOST_KEYS=1, SORT=2, DENSITY_LAMBDA=3, CORRECTION=4, VORTICITY=5, CONFINEMENT=6, XSPH=7, DROP_GHOSTS=8)
    BUF = dict(KEYS_IN=0, PRED_IN=1, PRED_CUR=2, LAMBDA=3, VEL_SORTED=4, VORT_NORM=5, VEL_CONFINED=6, LIST_BUILD_POS=7,
               LIST_INVALID=8)

    def __init__(self, capacity, box, grid, device, fluid_params=None, jacobi=3):
        self.device = torch.device("cuda", device)
        self.h = _abi.Handle(_abi.FLUIDS, capacity, 0, box, grid, 3, 0, device)
        self.capacity = capacity
        self.jacobi = jacobi
        fp = fluid_params or _abi.FluidParams(450.0, 600.0, 0.010, 3, 1, 0.006, 0.001, 4, 1, 0.0004, 0.0001)
        self.vorticity = bool(fp.isVorticityConfEnabled)
        self.h.set_fluid_params(fp, jacobi)
        self.stream = torch.cuda.ExternalStream(self.h.stream(), device=self.device)
        self.dmax_sq = self.h.shard_list_dmax_sq()

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
        return self._view(p, n, f4=name in ("PRED_IN", "PRED_CUR", "VEL_SORTED", "VEL_CONFINED", "LIST_BUILD_POS"),
                          u32=name in ("KEYS_IN", "LIST_INVALID"))

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

    def perm(self):
        return self.field("RadixSortIndices", u32=True)

    def stage(self, name, it=0, last=False):
        self.h.shard_stage(self.ST[name], it, last)

    def refresh_fields(self, name, last=False):
        """tensors (cell-sorted index space) whose ghost rows must be refreshed after stage `name`"""
        if name == "DENSITY_LAMBDA":
            return [self.buf("LAMBDA")]
        if name == "CORRECTION":
            return [self.buf("PRED_CUR")] + ([self.buf("VEL_SORTED")] if last else [])
        if name == "VORTICITY":
            return [self.buf("VORT_NORM")]
        if name == "CONFINEMENT":
            return [self.buf("VEL_CONFINED")]
        return []

    def pred_cur(self):
        return self.buf("PRED_CUR")

    def list_state(self):
        try:
            return self.buf("LIST_BUILD_POS"), self.buf("LIST_INVALID"), self.dmax_sq
        except _abi.RtpError:
            return None

    def sync(self):
        self.h.sync()


class SlabDecomposition:
    """One rank of the x-slab decomposition. `engine` is a CudaSlabEngine (or any object with the same interface)."""

    def __init__(self, engine, grid, rank=None, world=None, group=None):
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

    # ---- initial distribution: every rank is given the full initial state and keeps the particles of its slab
    def slab_of(self, keys):
        layer = torch.div(keys.to(torch.int64), self.grid[1] * self.grid[2], rounding_mode="floor")
        return layer.clamp_(max=self.grid[0] - 1)  # index RES (particle exactly on the +x wall) belongs to the last layer

    def load_owned(self, pos, vel):
        """pos/vel: (n, 4) float32 tensors on the engine's device holding ONLY this rank's particles, any order"""
        n = pos.shape[0]
        with self.e.stream_context():
            self.e.set_counts(n, n)
            self.e.pos()[:n] = pos
            self.e.vel()[:n] = vel
        self.n_owned = n

    # ---- exchanges with the left (rank-1) and right (rank+1) slab owners
    def _peers(self):
        return (self.rank - 1 if self.rank > 0 else None), (self.rank + 1 if self.rank < self.world - 1 else None)

    def _global(self, r):
        return r if self.group is None else dist.get_global_rank(self.group, r)

    def _exchange_counts(self, n_left, n_right, like):
        left, right = self._peers()
        ops, bufs = [], {}
        for peer, n, tag in ((left, n_left, "l"), (right, n_right, "r")):
            if peer is None:
                continue
            s = torch.tensor([n], dtype=torch.int64, device=like.device)
            rbuf = torch.zeros(1, dtype=torch.int64, device=like.device)
            bufs[tag] = (s, rbuf)
            ops.append(dist.P2POp(dist.isend, s, self._global(peer), self.group))
            ops.append(dist.P2POp(dist.irecv, rbuf, self._global(peer), self.group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        return (int(bufs["l"][1].item()) if "l" in bufs else 0), (int(bufs["r"][1].item()) if "r" in bufs else 0)

    def _exchange(self, send_left, send_right, n_from_left, n_from_right):
        """send rows to the neighbours, receive n_from_* rows from them (row shape/dtype of the sends).
        Every exchange posts exactly one send and one receive per existing neighbour, also for empty payloads (one pad
        row is appended), so the NCCL operation pattern never changes: a new pattern (e.g. the first send-only
        migration) otherwise costs a one-off 50-100 ms connection set-up in the middle of the run."""
        left, right = self._peers()
        ops = []
        tail = tuple(send_left.shape[1:])
        pad = torch.zeros((1,) + tail, dtype=send_left.dtype, device=send_left.device)
        recv_l = torch.empty((n_from_left + 1,) + tail, dtype=send_left.dtype, device=send_left.device)
        recv_r = torch.empty((n_from_right + 1,) + tail, dtype=send_left.dtype, device=send_left.device)
        if left is not None:
            ops.append(dist.P2POp(dist.isend, torch.cat([send_left, pad], dim=0), self._global(left), self.group))
            ops.append(dist.P2POp(dist.irecv, recv_l, self._global(left), self.group))
        if right is not None:
            ops.append(dist.P2POp(dist.isend, torch.cat([send_right, pad], dim=0), self._global(right), self.group))
            ops.append(dist.P2POp(dist.irecv, recv_r, self._global(right), self.group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        return recv_l[:n_from_left], recv_r[:n_from_right]

    def _apply_migration(self, pos, vel, n, holes, arrivals):
        """arrivals (k_in, 8) fill the slots `holes` of the leavers; a surplus is appended, a deficit is filled from the tail"""
        k_out, k_in = holes.numel(), arrivals.shape[0]
        m = min(k_out, k_in)
        if m:
            pos[holes[:m]] = arrivals[:m, :4]
            vel[holes[:m]] = arrivals[:m, 4:]
        if k_in > k_out:
            extra = k_in - k_out
            if n + extra > pos.shape[0]:
                raise RuntimeError("slab capacity exceeded: %d > %d" % (n + extra, pos.shape[0]))
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
        """Run the (rare) migration bookkeeping once on scratch tensors so that no lazily loaded torch kernel or allocator
        growth lands inside a timed step (first use of a torch op costs tens of ms)."""
        dev = self.e.pos().device
        with self.e.stream_context():
            for k_out, k_in in ((3, 1), (1, 3), (2, 2)):
                p = torch.zeros((16, 4), device=dev)
                v = torch.zeros((16, 4), device=dev)
                holes = torch.arange(1, 1 + k_out, device=dev) * 3
                self._apply_migration(p, v, 12, holes, torch.ones((k_in, 8), device=dev))
                torch.cat([p[holes], v[holes]], dim=1)
            lay = self.slab_of(torch.arange(8, device=dev, dtype=torch.int32))
            torch.nonzero(lay < 1).flatten()
            torch.nonzero((lay < 1) | (lay >= 2)).flatten()
        self.e.sync()

    def _mark(self, name):
        if self.profile and hasattr(self.e, "stream"):
            ev = torch.cuda.Event(enable_timing=True)
            ev.record(self.e.stream)
            self._marks.append((name, ev))

    def step(self):
        e = self.e
        self._marks = []
        with e.stream_context():
            self._mark("begin")
            self._step()
            self._mark("end")
        if self.profile and self._marks:
            e.sync()
            ph = {}
            for (_, a), (name, b) in zip(self._marks[:-1], self._marks[1:]):
                ph[name] = ph.get(name, 0.0) + a.elapsed_time(b)
            self.stats["phases"] = {k: round(v, 3) for k, v in ph.items()}

    def _step(self):
        e, n = self.e, self.n_owned
        jacobi = e.jacobi
        left, right = self._peers()
        pos, vel = e.pos(), e.vel()

        # 1. predict, migrate
        e.set_counts(n, n)
        e.stage("PREDICT")
        if self.world > 1:
            layer = self.slab_of(e.keys_in()[:n])
            go_l = (layer < self.xlo) if left is not None else torch.zeros_like(layer, dtype=torch.bool)
            go_r = (layer >= self.xhi) if right is not None else torch.zeros_like(layer, dtype=torch.bool)
            idx_l, idx_r = torch.nonzero(go_l).flatten(), torch.nonzero(go_r).flatten()
            n_in_l, n_in_r = self._exchange_counts(idx_l.numel(), idx_r.numel(), pos)
            self.stats["migrated_out"] = int(idx_l.numel() + idx_r.numel())
            k_out, k_in = idx_l.numel() + idx_r.numel(), n_in_l + n_in_r
            # every rank takes part in the (possibly empty) migration exchange: its neighbours cannot know it has nothing
            in_l, in_r = self._exchange(torch.cat([pos[idx_l], vel[idx_l]], dim=1), torch.cat([pos[idx_r], vel[idx_r]], dim=1),
                                        n_in_l, n_in_r)
            if k_out + k_in:
                # only the handful of migrating rows are touched: arrivals fill the slots of the leavers, a surplus is
                # appended, a deficit is filled from the tail (order inside a cell is not preserved for moved rows)
                n = self._apply_migration(pos, vel, n, torch.cat([idx_l, idx_r]), torch.cat([in_l, in_r], dim=0))
                e.set_counts(n, n)
                e.stage("PREDICT")
        self._mark("predict+migrate")

        # 2. halo of predicted positions, sort everything by cell
        ng_l = ng_r = 0
        if self.world > 1:
            layer = self.slab_of(e.keys_in()[:n])
            src_l = torch.nonzero(layer < self.xlo + GHOST_LAYERS).flatten() if left is not None else layer.new_empty(0)
            src_r = torch.nonzero(layer >= self.xhi - GHOST_LAYERS).flatten() if right is not None else layer.new_empty(0)
            ng_l, ng_r = self._exchange_counts(src_l.numel(), src_r.numel(), pos)
            pred_in = e.pred_in()
            gl, gr = self._exchange(pred_in[src_l], pred_in[src_r], ng_l, ng_r)
            ng = ng_l + ng_r
            if n + ng > e.capacity:
                raise RuntimeError("slab capacity exceeded: %d > %d" % (n + ng, e.capacity))
            ghosts = torch.cat([gl, gr], dim=0)
            pred_in[n:n + ng] = ghosts
            pos[n:n + ng] = ghosts
            vel[n:n + ng] = 0.0
            e.set_counts(n, n + ng)
            e.stage("GHOST_KEYS")
        n_loc = n + ng_l + ng_r
        self.stats.update(owned=n, ghosts=ng_l + ng_r)
        self._mark("halo")
        e.stage("SORT")
        self._mark("sort")

        refresh = None
        if self.world > 1:
            perm = e.perm()[:n_loc].to(torch.int64)
            inv = torch.empty_like(perm)
            inv[perm] = torch.arange(n_loc, device=perm.device)
            send_l, send_r = inv[src_l], inv[src_r]
            recv_l, recv_r = inv[n:n + ng_l], inv[n + ng_l:n_loc]

            def refresh(tensors):
                for t in tensors:
                    rl, rr = self._exchange(t[send_l], t[send_r], ng_l, ng_r)
                    if ng_l:
                        t[recv_l] = rl
                    if ng_r:
                        t[recv_r] = rr

            ghost_idx = torch.cat([recv_l, recv_r])
            lists = e.list_state()

            def check_ghost_displacement(next_epoch):
                # the owner's kernel checks the particles it moves; a ghost moved by ITS owner is checked here
                if lists is None or ghost_idx.numel() == 0:
                    return
                build_pos, invalid, dmax_sq = lists
                d = e.pred_cur()[ghost_idx, :3] - build_pos[ghost_idx, :3]
                moved = ((d * d).sum(dim=1) > dmax_sq).any().to(invalid.dtype)
                invalid[next_epoch] = torch.maximum(invalid[next_epoch], moved)
        else:
            def check_ghost_displacement(next_epoch):
                return
        self._mark("index-maps")

        # 3. the solver stages, each followed by the refresh of what it produced
        for it in range(jacobi):
            last = it == jacobi - 1
            e.stage("DENSITY_LAMBDA", it)
            self._mark("compute")
            if refresh:
                refresh(e.refresh_fields("DENSITY_LAMBDA"))
                self._mark("refresh")
            e.stage("CORRECTION", it, last)
            self._mark("compute")
            if refresh:
                refresh(e.refresh_fields("CORRECTION", last))
                check_ghost_displacement(it + 1)
                self._mark("refresh")
        if e.vorticity:
            e.stage("VORTICITY", jacobi)
            self._mark("compute")
            if refresh:
                refresh(e.refresh_fields("VORTICITY"))
                self._mark("refresh")
            e.stage("CONFINEMENT", jacobi)
            self._mark("compute")
            if refresh:
                refresh(e.refresh_fields("CONFINEMENT"))
                self._mark("refresh")
            e.stage("XSPH", jacobi)
            self._mark("compute")

        # 4. drop the ghosts; owned particles stay in cell-sorted order
        if self.world > 1 and n_loc > n:
            e.stage("DROP_GHOSTS")
        e.set_counts(n, n)
        self.n_owned = n
        self._mark("compact")

    def owned_state(self):
        n = self.n_owned
        with self.e.stream_context():
            return self.e.pos()[:n].clone(), self.e.vel()[:n].clone()


def split_initial_state(pos, box, grid, rank, world):
    """rows of `pos` (numpy (n,4)) whose cell x-layer belongs to `rank` (same cell formula as grid.cl:14-24 on the
    initial positions; any consistent assignment works, the first step migrates by predicted position anyway)"""
    import numpy as np
    h = np.float32(np.float32(box[0]) / np.float32(grid[0]))
    w = np.float32(box[0] / 2.0)
    x = np.floor((np.clip(pos[:, 0], -w, w) + w) / h).astype(np.int64).clip(max=grid[0] - 1)
    lo, hi = (grid[0] * rank) // world, (grid[0] * (rank + 1)) // world
    return np.nonzero((x >= lo) & (x < hi))[0]
