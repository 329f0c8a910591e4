#!/usr/bin/env python
"""Kernel-level timing of the BASELINE configs on one GPU (tuning aid; the judged numbers come from bench.py).
usage: python tools/quick_bench.py [cfg ...]   cfg in {1,2w,2f,2m,3,4s,4,5s,5,q}  (4s / 5s = configs 4 / 5 at n = 100 / 128;
2m = Wolf + Fennell in one pass, as bench.py evaluates config 2; q = Ewald with quadrupoles, n = 64)"""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import coulombgalore_b200 as cg

INF = float("inf")
CONFIGS = {
    "1": dict(pot=lambda: cg.qPotential(14, 3), n=16, a=1.75, seed=1001, pbc=False, mode=cg.MODE_ION, what=cg.ENERGY | cg.FORCE, falg=42),
    "2w": dict(pot=lambda: cg.Wolf(12, 0.25), n=100, a=3.0, seed=1002, pbc=True, mode=cg.MODE_ION, what=cg.FORCE, falg=33),
    "2f": dict(pot=lambda: cg.Fennell(12, 0.25), n=100, a=3.0, seed=1002, pbc=True, mode=cg.MODE_ION, what=cg.FORCE, falg=39),
    "3": dict(pot=lambda: cg.ReactionField(12, 80, 1, False), n=64, a=3.0, seed=1003, pbc=True, mode=cg.MODE_DIPOLE, what=cg.ENERGY | cg.FIELD | cg.FORCE, falg=101),
    "4s": dict(pot=lambda: cg.Ewald(10, 0.3, INF, 30.0), n=100, a=3.0, seed=1004, pbc=True, mode=cg.MODE_MULTIPOLE, what=cg.ENERGY | cg.FORCE, falg=133),
    "4": dict(pot=lambda: cg.Ewald(10, 0.3, INF, 30.0), n=160, a=3.0, seed=1004, pbc=True, mode=cg.MODE_MULTIPOLE, what=cg.ENERGY | cg.FORCE, falg=133),
    "5": dict(pot=lambda: cg.Splined().spline(cg.Poisson, 12, 4, 3), n=256, a=3.0, seed=1005, pbc=True, mode=cg.MODE_MULTIPOLE, what=cg.ENERGY | cg.FORCE, falg=153),
    "q": dict(pot=lambda: cg.Ewald(10, 0.3, INF, INF), n=64, a=3.0, seed=1006, pbc=True, mode=cg.MODE_MULTIPOLE_QUAD, what=cg.ENERGY | cg.FORCE, falg=400),
    "5s": dict(pot=lambda: cg.Splined().spline(cg.Poisson, 12, 4, 3), n=128, a=3.0, seed=1005, pbc=True, mode=cg.MODE_MULTIPOLE, what=cg.ENERGY | cg.FORCE, falg=153),
}

def fused_config2(dev, st):
    """config 2 as bench.py runs it: Wolf and Fennell in one pass over the pairs (cg_eval_multi_device)."""
    import time
    n, N = 100, 100 ** 3
    evs = [cg.BatchedEvaluator(cg.Wolf(12, 0.25)), cg.BatchedEvaluator(cg.Fennell(12, 0.25))]
    for ev in evs:
        ev.set_stream(st.cuda_stream)
    if os.environ.get("CG_SHARD"):  # emulate one rank of an R-rank run
        evs[0].set_shard(*[int(v) for v in os.environ["CG_SHARD"].split(",")])
    evs[1].share_system(evs[0])
    arr = [torch.empty(N, dtype=torch.float64, device=dev) for _ in range(4)]
    evs[0].generate_lattice_device(n, 3.0, 1002, arr[0], arr[1], arr[2], arr[3])
    f = [torch.empty((N, 3), dtype=torch.float64, device=dev) for _ in evs]
    sc = [torch.zeros(2, dtype=torch.float64, device=dev) for _ in evs]
    pm, wm = [], []
    for it in range(8):
        evs[0].set_system_device(arr[0], arr[1], arr[2], arr[3], None, [300.0] * 3, True)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        cg.eval_multi_device(evs, cg.MODE_ION, cg.FORCE, force=f, scalars=sc)
        torch.cuda.synchronize(); t1 = time.perf_counter()
        if it >= 3:
            pm.append(evs[0].timings()["pair_ms"]); wm.append((t1 - t0) * 1e3)
    pairs = sc[0][1].item() + sc[1][1].item()
    p = statistics.median(pm)
    print("cfg 2m  N=%8d pairs=%.4g (Wolf + Fennell, one pass)  pair_kernel %.3f ms  wall %.3f ms -> %.3e pairs/s  %.2f TF(F_alg=33+39) %.1f%% of 37.2" %
          (N, pairs, p, statistics.median(wm), pairs / p * 1e3, pairs / 2 * 72 / p * 1e3 / 1e12, 100 * pairs / 2 * 72 / p * 1e3 / 37.2e12), flush=True)
    for ev in evs:
        ev.close()


def main():
    names = sys.argv[1:] or ["1", "2w", "2f", "2m", "3", "4s", "5s"]
    dev = torch.device("cuda", 0)
    st = torch.cuda.Stream()
    torch.cuda.set_stream(st)
    for nm in names:
        if nm == "2m":
            fused_config2(dev, st)
            continue
        c = CONFIGS[nm]
        N = c["n"] ** 3
        ev = cg.BatchedEvaluator(c["pot"]())
        ev.set_stream(st.cuda_stream)
        if os.environ.get("CG_PRECISION"):
            ev.set_precision(os.environ["CG_PRECISION"])
        if os.environ.get("CG_SHARD"):  # emulate one rank of an R-rank run: CG_SHARD=rank,nranks
            ev.set_shard(*[int(v) for v in os.environ["CG_SHARD"].split(",")])
        arr = [torch.empty(N, dtype=torch.float64, device=dev) for _ in range(7)]
        ev.generate_lattice_device(c["n"], c["a"], c["seed"], arr[0], arr[1], arr[2], arr[3], (arr[4], arr[5], arr[6]))
        L = c["n"] * c["a"]
        need_mu = c["mode"] != cg.MODE_ION
        need_q = c["mode"] != cg.MODE_DIPOLE
        ev.set_system_device(arr[0], arr[1], arr[2], arr[3] if need_q else None, (arr[4], arr[5], arr[6]) if need_mu else None, [L] * 3, c["pbc"])
        if c["mode"] == cg.MODE_MULTIPOLE_QUAD:  # random (neither symmetric nor traceless) quadrupole tensors
            quad = torch.rand((N, 9), dtype=torch.float64, device=dev, generator=torch.Generator(device=dev).manual_seed(c["seed"])) - 0.5
            ev.set_quadrupoles_device(quad)
        f = torch.empty((N, 3), dtype=torch.float64, device=dev)
        e = torch.empty((N, 3), dtype=torch.float64, device=dev)
        sc = torch.zeros(2, dtype=torch.float64, device=dev)
        pm, sm, wm = [], [], []
        import time
        for it in range(8):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            ev.eval_device(c["mode"], c["what"], force=f, field=e if c["what"] & cg.FIELD else None, scalars=sc)
            torch.cuda.synchronize(); t1 = time.perf_counter()
            t = ev.timings()
            if it >= 3:
                pm.append(t["pair_ms"]); sm.append(t["sort_ms"]); wm.append((t1 - t0) * 1e3)
        pairs = sc[1].item()
        p = statistics.median(pm)
        cn = ev.counters()
        print("cfg %-3s N=%8d pairs=%.4g  pair_kernel %.3f ms  sort %.3f ms  wall %.3f ms -> %.3e pairs/s  %.2f TF(F_alg=%d) %.1f%% of 37.2  [entries %d groups %d split %d]" %
              (nm, N, pairs, p, statistics.median(sm), statistics.median(wm), pairs / p * 1e3, pairs * c["falg"] / p * 1e3 / 1e12, c["falg"], 100 * pairs * c["falg"] / p * 1e3 / 37.2e12, cn["sorted_entries"], cn["groups"], cn["split"]), flush=True)
        ev.close()

if __name__ == "__main__":
    main()
