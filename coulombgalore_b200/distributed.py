"""Multi-GPU evaluation: i-particles are sharded, j-particles are exchanged.

Every rank owns a contiguous slice of the particles.  One step = all-gather of the SoA inputs (positions, charges,
dipoles) over the process group, evaluation of this rank's shard of i-groups against all j, and a scalar all-reduce of
the energy / pair count.  Per-particle outputs stay sharded: rank r's kernels write only the particles of its i-groups.
The collective layer is torch.distributed (NCCL on GPUs; gloo in the CPU tests); the evaluator is injected, so the same
plumbing is tested on CPU with the oracle standing in for the CUDA evaluator.
"""
import torch
import torch.distributed as dist


def owned_range(n, rank, world):
    return n * rank // world, n * (rank + 1) // world


class ShardedStep:
    def __init__(self, n_total, n_arrays, device, dtype=torch.float64, group=None):
        self.n, self.group = n_total, group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.counts = [owned_range(n_total, r, self.world)[1] - owned_range(n_total, r, self.world)[0] for r in range(self.world)]
        self.even = len(set(self.counts)) == 1
        self.full = [torch.empty(n_total, dtype=dtype, device=device) for _ in range(n_arrays)]
        self._coalescing_manager = getattr(dist, "_coalescing_manager", None)
        self.coalesce = (self._coalescing_manager is not None and self.world > 1 and torch.device(device).type == "cuda"
                         and dist.get_backend(group) == "nccl")

    def gather(self, owned):
        """owned: this rank's slices, one tensor per SoA array -> the full arrays on every rank."""
        if self.world == 1:
            for f, o in zip(self.full, owned):
                f.copy_(o)
            return self.full
        if self.even and self.coalesce:
            # one NCCL group (a single fused launch) for all SoA arrays instead of one collective per array.  The manager is a
            # private torch API: a signature that differs from the one probed here switches the coalescing off for good.
            try:
                cm = self._coalescing_manager(group=self.group, device=self.full[0].device, async_ops=False)
            except TypeError:
                self.coalesce = False
                cm = None
            if cm is not None:
                with cm:
                    for f, o in zip(self.full, owned):
                        dist.all_gather_into_tensor(f, o, group=self.group)
                return self.full
        for f, o in zip(self.full, owned):
            if self.even:
                dist.all_gather_into_tensor(f, o, group=self.group)
            else:  # uneven slices: gather padded blocks, then compact
                cmax = max(self.counts)
                pad = torch.zeros(cmax, dtype=f.dtype, device=f.device)
                pad[: o.numel()] = o
                buf = torch.empty(self.world * cmax, dtype=f.dtype, device=f.device)
                dist.all_gather_into_tensor(buf, pad, group=self.group)
                off = 0
                for r, c in enumerate(self.counts):
                    f[off:off + c] = buf[r * cmax:r * cmax + c]
                    off += c
        return self.full

    def reduce_scalars(self, scalars):
        """scalars: tensor [U_shard, pairs_shard] -> totals on every rank."""
        if self.world > 1:
            dist.all_reduce(scalars, op=dist.ReduceOp.SUM, group=self.group)
        return scalars


class _DeviceBlock:
    """Zero-copy torch view of device memory owned by the library (numba-style __cuda_array_interface__)."""

    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f8", "data": (int(ptr), False), "version": 2, "strides": None}


class PeerExchange(ShardedStep):
    """ShardedStep whose gather is this repo's own kernel over peer memory instead of an NCCL all-gather.

    Every rank keeps its owned SoA slices in a block of IPC-shareable device memory (two blocks, used alternately); the
    other ranks map it (cudaIpcOpenMemHandle through cg_peer_open) and `gather` PULLS all blocks over NVLink / NVSwitch
    with one kernel (cg_peer_gather) on the evaluator's stream.  One small all-reduce per step orders the ranks: after
    it, every rank has written block (step % 2) and has finished reading block ((step - 1) % 2) -- the pull of step
    s-1 precedes the all-reduce of step s in stream order -- so the owner may rewrite that one in step s + 1.
    Needs equal slices, one process per GPU on one node and CUDA IPC; `PeerExchange.create` falls back to the NCCL
    ShardedStep (with the reason in `.fallback_reason`) where that is not available."""

    def __init__(self, evaluator, n_total, n_arrays, device, group=None):
        import ctypes as C
        from ._capi import check, lib
        super().__init__(n_total, n_arrays, device, group=group)
        if not self.even:
            raise RuntimeError("PeerExchange needs equal slices on all ranks")
        self.ev, self.lib, self.C = evaluator, lib, C
        self.device = torch.device(device)
        self.count = self.counts[self.rank]
        nbytes = max(16, n_arrays * self.count * 8)
        self.own_ptr, handles = [], []
        for _ in range(2):
            p, h = C.c_void_p(), C.create_string_buffer(64)
            if lib.cg_peer_alloc(self.device.index, nbytes, C.byref(p), h) != 0:
                handles = None  # still take part in the collective below, so that no rank is left waiting
                break
            self.own_ptr.append(p)
            handles.append(h.raw)
        everyone = [None] * self.world
        dist.all_gather_object(everyone, handles, group=self.group)
        if any(e is None for e in everyone):
            raise RuntimeError("cg_peer_alloc / cudaIpcGetMemHandle failed on rank(s) %s" % [r for r, e in enumerate(everyone) if e is None])
        self.block_ptrs = []  # [buffer][rank] -> device pointer valid in this process
        self._opened = []
        for b in range(2):
            row = []
            for r in range(self.world):
                if r == self.rank:
                    row.append(self.own_ptr[b].value)
                else:
                    p = C.c_void_p()
                    check(lib.cg_peer_open(self.device.index, everyone[r][b], C.byref(p)), "cg_peer_open")
                    self._opened.append(p)
                    row.append(p.value)
            self.block_ptrs.append((C.c_void_p * self.world)(*row))
        self.counts_c = (C.c_int64 * self.world)(*self.counts)
        self.full_c = (C.c_void_p * n_arrays)(*[t.data_ptr() for t in self.full])
        # torch views of the two owned blocks: [n_arrays, count]
        self.owned_blocks = [torch.as_tensor(_DeviceBlock(self.own_ptr[b].value, (n_arrays, max(self.count, 1))), device=self.device)
                             for b in range(2)]
        self.flag = torch.zeros(1, dtype=torch.float32, device=self.device)
        self.step = 0
        self.fallback_reason = None

    @staticmethod
    def create(evaluator, n_total, n_arrays, device, group=None):
        """PeerExchange where CUDA IPC works and slices are equal, else the NCCL ShardedStep (all ranks decide alike)."""
        reason = None
        try:
            ex = PeerExchange(evaluator, n_total, n_arrays, device, group=group)
        except Exception as e:  # noqa: BLE001 -- any failure means: use the collective library
            ex, reason = None, "%s: %s" % (type(e).__name__, e)
        ok = torch.tensor([0.0 if ex is None else 1.0], device=device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if ok.item() < 1.0:
            if ex is not None:
                ex.close()
            ex = ShardedStep(n_total, n_arrays, device, group=group)
            ex.fallback_reason = reason or "a peer rank could not map the blocks"
        return ex

    def gather(self, owned):
        # The two-block protocol orders publish -> all-reduce -> pull by STREAM order: the evaluator's kernels (cg_peer_gather)
        # must run on the stream the copies and the collective are enqueued on, i.e. torch's current stream.
        self.ev.set_stream(torch.cuda.current_stream(self.device).cuda_stream)
        b = self.step & 1
        self.step += 1
        blk = self.owned_blocks[b]
        # publish: this rank's new positions / charges / dipoles go into its block (one copy for a stacked
        # [n_arrays, count] tensor; an integrator would write them there directly: see next_block())
        if owned is None:
            pass
        elif torch.is_tensor(owned):
            blk[:, : owned.shape[1]].copy_(owned)
        else:
            for a, o in enumerate(owned):
                blk[a, : o.numel()].copy_(o)
        dist.all_reduce(self.flag, group=self.group)  # everyone has published b and is done reading the other block
        from ._capi import check
        check(self.lib.cg_peer_gather(self.ev.h, self.world, self.block_ptrs[b], self.counts_c, len(self.full), self.full_c),
              "cg_peer_gather", self.ev.h)
        return self.full

    def next_block(self):
        """The [n_arrays, count] view of the block the next gather() publishes: write the new owned data here and call
        gather(None) to skip the copy."""
        return self.owned_blocks[self.step & 1]

    def close(self):
        torch.cuda.synchronize(self.device)
        if dist.is_initialized():
            dist.barrier(group=self.group)  # nobody reads a block that is about to be freed
        for p in self._opened:
            self.lib.cg_peer_close(p)
        self._opened = []
        for p in self.own_ptr:
            self.lib.cg_peer_free(p)
        self.own_ptr = []


class HaloExchange:
    """Exchange step that moves only what a rank needs (SURVEY.md 8e "halo-only exchange").

    Rank t evaluates a block of cell rows and touches only the particles in the cell layers within one cut-off of it
    (`cg_shard_window`).  Per step every rank sorts its owned particles into one segment per destination inside its
    peer-mapped block (`cg_halo_publish_step`) and pulls just its segments from all blocks over NVLink
    (`cg_halo_pull_step`).  The ranks are ordered on the device -- a `published` word per block that the pull kernels spin
    on, acknowledgement words that free a block for rewriting two steps later -- so a step contains no collective at all
    (CG_HALO_SYNC=nccl: one small all-reduce per step instead).  The evaluator then works on the received subset; `ids` maps
    subset entries back to global particle indices.

        ex = HaloExchange(ev, n_total, box, n_payload)         # payload arrays: x, y, z first, then q / mu as needed
        arrays, ids, n_dev = ex.gather(owned2d)                # owned2d: [n_payload, count] of this rank
        ex.set_system(ev, arrays, ...)                         # cg_set_system_device + cg_set_count_device
    """

    def __init__(self, evaluator, n_total, box, n_payload, device, group=None, capacity=None):
        import ctypes as C
        from ._capi import check, lib
        self.ev, self.lib, self.C, self.check = evaluator, lib, C, check
        self.group, self.device = group, torch.device(device)
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.n_total, self.n_payload, self.narrays = n_total, n_payload, n_payload + 1  # + the global index
        self.counts = [owned_range(n_total, r, self.world)[1] - owned_range(n_total, r, self.world)[0] for r in range(self.world)]
        self.count = self.counts[self.rank]
        self.lo = owned_range(n_total, self.rank, self.world)[0]
        self.box = [float(b) for b in box]
        check(lib.cg_set_grid_count(evaluator.h, n_total), "cg_set_grid_count", evaluator.h)
        boxc = (C.c_double * 3)(*self.box)
        zlo, zhi = (C.c_double * self.world)(), (C.c_double * self.world)()
        for t in range(self.world):
            w = (C.c_double * 2)()
            check(lib.cg_shard_window(evaluator.h, boxc, t, self.world, n_total, w), "cg_shard_window", evaluator.h)
            zlo[t], zhi[t] = w[0], w[1]
        self.zlo, self.zhi = zlo, zhi
        nbytes = lib.cg_halo_block_bytes(self.world, self.narrays, self.count)
        self.own_ptr, handles = [], []
        for _ in range(2):
            p, h = C.c_void_p(), C.create_string_buffer(64)
            if lib.cg_peer_alloc(self.device.index, nbytes, C.byref(p), h) != 0:
                handles = None  # still take part in the collective below, so that no rank is left waiting
                break
            self.own_ptr.append(p)
            handles.append(h.raw)
        everyone = [None] * self.world
        dist.all_gather_object(everyone, handles, group=group)
        if any(e is None for e in everyone):
            raise RuntimeError("cg_peer_alloc / cudaIpcGetMemHandle failed on rank(s) %s" % [r for r, e in enumerate(everyone) if e is None])
        self.block_ptrs, self._opened = [], []
        for b in range(2):
            row = []
            for r in range(self.world):
                if r == self.rank:
                    row.append(self.own_ptr[b].value)
                else:
                    p = C.c_void_p()
                    check(lib.cg_peer_open(self.device.index, everyone[r][b], C.byref(p)), "cg_peer_open")
                    self._opened.append(p)
                    row.append(p.value)
            self.block_ptrs.append((C.c_void_p * self.world)(*row))
        self.counts_c = (C.c_int64 * self.world)(*self.counts)
        # what this rank receives: at most everything; the engine is told a tighter bound after the first step
        self.capacity = int(capacity or n_total)
        self.recv = torch.zeros((self.narrays, self.capacity), dtype=torch.float64, device=self.device)
        self.recv_c = (C.c_void_p * self.narrays)(*[self.recv[a].data_ptr() for a in range(self.narrays)])
        self.n_out = torch.zeros(2, dtype=torch.int32, device=self.device)
        self.ids_owned = torch.arange(self.lo, self.lo + self.count, dtype=torch.float64, device=self.device)
        self.flag = torch.zeros(1, dtype=torch.float32, device=self.device)
        self.step, self.n_bound = 0, None
        # CG_HALO_SYNC=nccl: order the ranks with one small all-reduce per step instead of the device-side flags
        import os
        self.device_flags = os.environ.get("CG_HALO_SYNC", "flags") != "nccl"
        self._n_host = torch.zeros(2, dtype=torch.int32).pin_memory() if self.device.type == "cuda" else torch.zeros(2, dtype=torch.int32)
        self._n_event = torch.cuda.Event() if self.device.type == "cuda" else None
        self._n_pending, self.last_count = False, 0
        # the pull kernel posts the received count straight into the pinned words: no copy operation in the stream
        self._mirror = self.device.type == "cuda" and lib.cg_halo_set_count_mirror(evaluator.h, self._n_host.data_ptr()) == 0

    @staticmethod
    def create(evaluator, n_total, box, n_payload, device, group=None):
        """HaloExchange where CUDA IPC works on every rank, else None with the reason (all ranks decide alike)."""
        ex, reason = None, None
        try:
            ex = HaloExchange(evaluator, n_total, box, n_payload, device, group=group)
        except Exception as e:  # noqa: BLE001 -- any failure means: use another exchange
            reason = "%s: %s" % (type(e).__name__, e)
        ok = torch.tensor([0.0 if ex is None else 1.0], device=device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if ok.item() < 1.0:
            if ex is not None:
                ex.close()
            return None, reason or "a peer rank could not map the blocks"
        return ex, None

    def gather(self, owned2d):
        """owned2d: [n_payload, count] float64 on the device (row 2 = z).  -> (list of payload arrays [bound], ids [bound],
        n_out int32[2] on the device: particles received, overflow flag).  `bound` >= the received count."""
        C = self.C
        # publish -> all-reduce -> pull are ordered by STREAM order (see PeerExchange.gather): bind the evaluator to torch's
        # current stream, whatever stream it was created with
        self.ev.set_stream(torch.cuda.current_stream(self.device).cuda_stream)
        self._track_bound()
        b = self.step & 1
        self.step += 1
        src = (C.c_void_p * self.narrays)(*([owned2d[a].data_ptr() for a in range(self.n_payload)] + [self.ids_owned.data_ptr()]))
        self._keep = owned2d
        if self.device_flags:
            # the ranks are ordered on the device: `published` / acknowledgement words in the block headers (cg_b200.h)
            s = self.step - 1
            prev = self.block_ptrs[(s - 1) & 1] if s >= 1 else None
            self.check(self.lib.cg_halo_publish_step(self.ev.h, self.world, self.narrays, 2, self.count, src, self.zlo, self.zhi, self.box[2],
                                                     self.own_ptr[b], s, self.rank, prev, self.n_out.data_ptr() + 4), "cg_halo_publish_step", self.ev.h)
            self.check(self.lib.cg_halo_pull_step(self.ev.h, self.world, self.rank, self.block_ptrs[b], self.counts_c, self.narrays, self.recv_c,
                                                  self.capacity, self.n_out.data_ptr(), s), "cg_halo_pull_step", self.ev.h)
        else:
            self.check(self.lib.cg_halo_publish(self.ev.h, self.world, self.narrays, 2, self.count, src, self.zlo, self.zhi, self.box[2],
                                                self.own_ptr[b]), "cg_halo_publish", self.ev.h)
            dist.all_reduce(self.flag, group=self.group)  # every owner has published block b; nobody still reads the other block
            self.check(self.lib.cg_halo_pull(self.ev.h, self.world, self.rank, self.block_ptrs[b], self.counts_c, self.narrays, self.recv_c,
                                             self.capacity, self.n_out.data_ptr()), "cg_halo_pull", self.ev.h)
        if self.n_bound is None:  # first step: one host read to size the engine's launches (later steps reuse it)
            got = self.n_out.cpu()
            if int(got[1]) != 0:
                raise RuntimeError("HaloExchange: capacity %d too small" % self.capacity)
            self.n_bound = min(self.capacity, int(got[0]) + int(got[0]) // 8 + 1024)
        else:  # the count of this step travels to pinned memory behind the pull; a later gather() reads it without waiting
            if not self._mirror:
                self._n_host.copy_(self.n_out, non_blocking=True)
            self._n_event.record(torch.cuda.current_stream(self.device))
            self._n_pending = True
        nb = self.n_bound
        return [self.recv[a, :nb] for a in range(self.n_payload)], self.recv[self.n_payload, :nb], self.n_out

    def exchange(self, ev, owned2d, has_q=True, has_mu=False):
        """gather() + set_system() in ONE library call (cg_halo_exchange_step): for small shards the host time between the
        launches is what the GPUs wait for.  owned2d: [n_payload, count] float64 on the device, rows x, y, z[, q][, mu].
        -> number of rows of the evaluator's outputs (the bound of the received count).  Device-side ordering only."""
        C = self.C
        if not self.device_flags or self.n_bound is None:  # the first step sizes the bound (one host read)
            sub, _, _ = self.gather(owned2d)
            self.set_system(ev, sub, has_q=has_q, has_mu=has_mu)
            return self.n_bound
        self.ev.set_stream(torch.cuda.current_stream(self.device).cuda_stream)
        self._track_bound()
        s = self.step
        self.step += 1
        b = s & 1
        cache = self.__dict__.setdefault("_src_cache", {})  # pointer tables of the caller's (persistent, e.g. double-buffered) input blocks
        key = owned2d.data_ptr()
        if key not in cache:
            if len(cache) > 8:
                cache.clear()
            cache[key] = (C.c_void_p * self.narrays)(*([owned2d[a].data_ptr() for a in range(self.n_payload)] + [self.ids_owned.data_ptr()]))
            self._boxc = (C.c_double * 3)(*self.box)
        self._src = cache[key]
        self._keep = owned2d
        self.check(self.lib.cg_halo_exchange_step(ev.h, self.world, self.rank, self.narrays, 2, self.count, self._src, self.zlo, self.zhi,
                                                  self._boxc, self.own_ptr[b], self.block_ptrs[b], self.block_ptrs[(s - 1) & 1] if s >= 1 else None,
                                                  self.counts_c, self.recv_c, self.capacity, self.n_bound, self.n_out.data_ptr(), s,
                                                  int(has_q), int(has_mu)), "cg_halo_exchange_step", ev.h)
        ev.n = self.n_bound
        if not self._mirror:
            self._n_host.copy_(self.n_out, non_blocking=True)
        self._n_event.record(torch.cuda.current_stream(self.device))
        self._n_pending = True
        return self.n_bound

    def ids(self):
        """Global particle index (as float64) of every row of the evaluator's outputs after exchange()."""
        return self.recv[self.n_payload, : self.n_bound]

    def collect_owned(self, src, dst):
        """Rows of the rank's own particles from outputs in received order (lists of [n_bound, w] device tensors) into `dst`
        (lists of [count, w] tensors, the order of the owned arrays given to exchange()): cg_halo_collect_owned, one launch."""
        C = self.C
        n = len(src)
        width = src[0].shape[1] if src[0].dim() > 1 else 1
        cache = self.__dict__.setdefault("_collect_cache", {})
        key = (tuple(t.data_ptr() for t in src), tuple(t.data_ptr() for t in dst))
        if key not in cache:
            if len(cache) > 8:
                cache.clear()
            cache[key] = ((C.c_void_p * n)(*key[0]), (C.c_void_p * n)(*key[1]))
        sp, dp = cache[key]
        rc = self.lib.cg_halo_collect_owned(self.ev.h, self.ids().data_ptr(), self.n_out.data_ptr(), self.n_bound, self.lo, self.count, n,
                                            width, sp, dp)
        if rc != 0:
            raise RuntimeError("cg_halo_collect_owned failed (%d)" % rc)

    def _track_bound(self):
        """Grow the bound handed to the engine before the received count reaches it.  A step that does exceed it is not
        silent: the engine raises its overflow flag (energy NaN, pair count -1) when the device count is larger than the
        arrays it was given."""
        if self.n_bound is None or not self._n_pending or not self._n_event.query():
            return
        self._n_pending = False
        got = int(self._n_host[0])
        self.last_count = got
        if got + got // 32 + 256 > self.n_bound:
            self.n_bound = min(self.capacity, got + got // 8 + 1024)

    def set_system(self, ev, arrays, has_q=True, has_mu=False, pbc=True):
        """Hand the received subset to the evaluator: arrays = x, y, z[, q][, mux, muy, muz]."""
        k = 3
        q = arrays[k] if has_q else None
        k += 1 if has_q else 0
        mu = (arrays[k], arrays[k + 1], arrays[k + 2]) if has_mu else None
        ev.set_system_device(arrays[0], arrays[1], arrays[2], q, mu, self.box, pbc)
        self.check(self.lib.cg_set_count_device(ev.h, self.n_out.data_ptr()), "cg_set_count_device", ev.h)

    def check_overflow(self):
        """Host read of the overflow flag and the received count (synchronises)."""
        got = self.n_out.cpu()
        if int(got[1]) == 2:
            raise RuntimeError("HaloExchange: a peer rank did not publish / acknowledge its block in time")
        if int(got[1]) != 0 or (self.n_bound is not None and int(got[0]) > self.n_bound):
            raise RuntimeError("HaloExchange: received %d particles, bound %s, capacity %d" % (int(got[0]), self.n_bound, self.capacity))
        return int(got[0])

    def close(self):
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)
        for p in self._opened:
            self.lib.cg_peer_close(p)
        self._opened = []
        for p in self.own_ptr:
            self.lib.cg_peer_free(p)
        self.own_ptr = []
