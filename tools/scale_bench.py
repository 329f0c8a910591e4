#!/usr/bin/env python
"""Multi-GPU scaling of the BASELINE configs 2-5: one process per GPU, every rank owns N/R particles.
One step = NCCL all-gather of the SoA inputs + evaluation of this rank's shard of i-groups + scalar all-reduce; timed on
the device, max over ranks.  Launch:  python -m torch.distributed.run --nnodes=1 --nproc-per-node R --master-addr 127.0.0.1
--master-port 29511 tools/scale_bench.py [cfg ...]      (R = 1 works without torchrun).  One JSON line per config."""
import json, os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import coulombgalore_b200 as cg
from coulombgalore_b200.distributed import HaloExchange, PeerExchange, ShardedStep, owned_range

INF = float("inf")
CONFIGS = {
    "2w": dict(pot=lambda: cg.Wolf(12, 0.25), n=100, seed=1002, mode=cg.MODE_ION, what=cg.FORCE, falg=33),
    "3": dict(pot=lambda: cg.ReactionField(12, 80, 1, False), n=64, seed=1003, mode=cg.MODE_DIPOLE, what=cg.ENERGY | cg.FIELD | cg.FORCE, falg=101),
    "4": dict(pot=lambda: cg.Ewald(10, 0.3, INF, 30.0), n=160, seed=1004, mode=cg.MODE_MULTIPOLE, what=cg.ENERGY | cg.FORCE, falg=133),
    "5": dict(pot=lambda: cg.Splined().spline(cg.Poisson, 12, 4, 3), n=256, seed=1005, mode=cg.MODE_MULTIPOLE, what=cg.ENERGY | cg.FORCE, falg=153),
}


def main():
    names = [a for a in sys.argv[1:] if not a.startswith("-")] or ["2w", "3", "4", "5"]
    steps, warmup = 10, 3
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    for nm in names:
        c = CONFIGS[nm]
        n_side, a = c["n"], 3.0
        N, L = n_side ** 3, n_side * a
        ev = cg.BatchedEvaluator(c["pot"](), local)
        ev.set_stream(stream.cuda_stream)
        ev.set_shard(rank, world)
        arr = [torch.empty(N, dtype=torch.float64, device=dev) for _ in range(7)]
        ev.generate_lattice_device(n_side, a, c["seed"], arr[0], arr[1], arr[2], arr[3], (arr[4], arr[5], arr[6]))
        use = [0, 1, 2] + ([3] if c["mode"] != cg.MODE_DIPOLE else []) + ([4, 5, 6] if c["mode"] != cg.MODE_ION else [])
        lo, hi = owned_range(N, rank, world)
        owned = [arr[k][lo:hi].clone() for k in use]
        # CG_EXCHANGE = halo (default): pull only the particles of the rank's window; peer: pull everything (own kernel);
        # nccl: the collective library's all-gather
        mode_x = os.environ.get("CG_EXCHANGE", "halo") if world > 1 else "nccl"
        hx = HaloExchange.create(ev, N, [L] * 3, len(use), dev)[0] if mode_x == "halo" else None
        ex = ShardedStep(N, len(use), dev) if mode_x != "peer" else PeerExchange.create(ev, N, len(use), dev)
        del arr
        f = torch.empty((N, 3), dtype=torch.float64, device=dev)
        e = torch.empty((N, 3), dtype=torch.float64, device=dev) if c["what"] & cg.FIELD else None
        sc = torch.zeros(2, dtype=torch.float64, device=dev)
        flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)

        marks = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(steps)]

        if isinstance(ex, PeerExchange) or hx is not None:
            owned = torch.stack(owned)  # one publish copy per step instead of one per array

        def step(mk=None):
            if hx is not None:
                sub, ids, n_out = hx.gather(owned)
                if mk:
                    mk[0].record()
                hx.set_system(ev, sub, has_q=3 in use, has_mu=4 in use)
                nb = sub[0].numel()
                ev.eval_device(c["mode"], c["what"], force=f[:nb], field=None if e is None else e[:nb], scalars=sc)
            else:
                full = ex.gather(owned)
                if mk:
                    mk[0].record()
                g = dict(zip(use, full))
                mu = (g[4], g[5], g[6]) if 4 in g else None
                ev.set_system_device(g[0], g[1], g[2], g.get(3), mu, [L] * 3, True)
                ev.eval_device(c["mode"], c["what"], force=f, field=e, scalars=sc)
            if mk:
                mk[1].record()
            ex.reduce_scalars(sc)
            if mk:
                mk[2].record()

        def barrier():
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()

        for _ in range(warmup):
            step()
        barrier()
        e0 = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        e1 = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        pair_ms, sort_ms = [], []
        for s in range(steps):
            flush.fill_(s & 255)
            barrier()
            e0[s].record()
            step(marks[s])
            e1[s].record()
            e1[s].synchronize()
            t = ev.timings()
            pair_ms.append(t["pair_ms"])
            sort_ms.append(t["sort_ms"])
        barrier()
        ms = torch.tensor([statistics.mean(e0[s].elapsed_time(e1[s]) for s in range(steps)), statistics.mean(pair_ms), statistics.mean(sort_ms),
                           statistics.mean(e0[s].elapsed_time(marks[s][0]) for s in range(steps)),
                           statistics.mean(marks[s][0].elapsed_time(marks[s][1]) for s in range(steps)),
                           statistics.mean(marks[s][1].elapsed_time(marks[s][2]) for s in range(steps))],
                          dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        pairs = sc[1].item()  # all-reduced total
        if rank == 0:
            print(json.dumps({"cfg": nm, "n_gpus": world, "n_particles": N, "pairs_per_step": pairs, "ms_per_step": ms[0].item(),
                              "pair_kernel_ms_max": ms[1].item(), "sort_ms_max": ms[2].item(),
                              "gather_ms_max": ms[3].item(), "eval_ms_max": ms[4].item(), "allreduce_ms_max": ms[5].item(), "pairs_per_s": pairs / ms[0].item() * 1e3,
                              "tflops_alg": pairs * c["falg"] / ms[0].item() * 1e3 / 1e12, "energy": sc[0].item(),
                              "exchange": "%s of %d fp64 arrays (%d B/particle) + all-reduce of 2 scalars" %
                              ("halo pull (%d of %d particles received)" % (hx.check_overflow(), N) if hx is not None else
                               "peer-memory pull kernel" if isinstance(ex, PeerExchange) else "NCCL all-gather", len(use), 8 * len(use)),
                              "exchange_fallback": getattr(ex, "fallback_reason", None)}), flush=True)
        if isinstance(ex, PeerExchange):
            ex.close()
        if hx is not None:
            hx.close()
        ev.close()
        del f, e, flush, ex, owned
        torch.cuda.empty_cache()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
