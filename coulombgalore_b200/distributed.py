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
            # one NCCL group (a single fused launch) for all SoA arrays instead of one collective per array
            with self._coalescing_manager(group=self.group, device=self.full[0].device, async_ops=False):
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
