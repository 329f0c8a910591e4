"""world_size = 2 over gloo on CPU: the exchange plumbing of coulombgalore_b200/distributed.py (all-gather of the SoA
inputs, sharded evaluation, scalar all-reduce) with the oracle standing in for the CUDA evaluator.  The sum of the two
ranks' shards must reproduce the single-rank result exactly (forces) / to rounding (energy)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import oracle as O
from zoo import oracle_params


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_sides, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from coulombgalore_b200.distributed import ShardedStep, owned_range
        for n_side in n_sides:
            x, y, z, q, mu = O.generate(n_side, 3.0, 99, "port")
            N = x.size
            lo, hi = owned_range(N, rank, world)
            step = ShardedStep(N, 4, "cpu")
            owned = [torch.from_numpy(a[lo:hi].copy()) for a in (x, y, z, q)]
            full = [t.numpy() for t in step.gather(owned)]
            assert all(np.array_equal(f, a) for f, a in zip(full, (x, y, z, q)))  # every rank now holds all j
            s = O.Scheme(oracle_params("wolf"), "port")
            L = n_side * 3.0
            r = s.batched(full[0], full[1], full[2], full[3], None, [L] * 3, True, O.MODE_ION, O.FORCE | O.ENERGY, lo, hi)  # this rank's i-block
            sc = step.reduce_scalars(torch.tensor([r["energy"], float(r["pairs"])], dtype=torch.float64))
            cmax = max(step.counts)
            mine = torch.zeros((cmax, 3), dtype=torch.float64)
            mine[: hi - lo] = torch.from_numpy(r["force"])
            blocks = [torch.zeros((cmax, 3), dtype=torch.float64) for _ in range(world)]
            dist.all_gather(blocks, mine)
            if rank == 0:
                force = torch.cat([b[:c] for b, c in zip(blocks, step.counts)]).numpy().copy()
                out.put((n_side, sc.numpy().copy(), force))
    finally:
        dist.destroy_process_group()


def test_two_ranks_reproduce_one():
    n_sides = [8, 9]  # 512 particles (even split) and 729 (uneven split)
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_sides, out)) for r in range(2)]
    for p in procs:
        p.start()
    results = [out.get(timeout=240) for _ in n_sides]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for n_side, sc, force in results:
        x, y, z, q, mu = O.generate(n_side, 3.0, 99, "port")
        L = n_side * 3.0
        ref = O.Scheme(oracle_params("wolf"), "port").batched(x, y, z, q, None, [L] * 3, True, O.MODE_ION, O.FORCE | O.ENERGY)
        assert int(sc[1]) == ref["pairs"]
        assert abs(sc[0] - ref["energy"]) <= 1e-13 * ref["energy_abs"]
        assert np.array_equal(force, ref["force"])
