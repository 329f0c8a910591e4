#!/usr/bin/env python
"""bench.py -- pair interactions/s (fp64) of the B200 batched CoulombGalore engine.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W

Headline workload (BASELINE.json configs[1]): Wolf(12, 0.25) and Fennell(12, 0.25) ion-ion forces, N = 100^3 = 1 000 000
charges on the jittered lattice of SURVEY.md 8d (a = 3, L = 300, periodic, seed 1002), cell list under R_c = 12, fp64.
One STEP = the exchange of the inputs (N > 1), the spatial sort of the configuration, one Wolf and one Fennell evaluation of
all per-particle forces on it and the reductions (positions are treated as new every step).  The unit of work is one ordered
in-cutoff pair (i <- j); `value` = pairs evaluated by all ranks per second with the inputs resident in HBM; `e2e` is the same
metric with HOST buffers: H2D of positions / charges and D2H of the forces inside the timing.

Config 1 (4096 charges, open box) is evaluated by the all-pairs tile kernel on one GPU (no sort; TMA-staged j-tiles).

N > 1: every rank owns N/R particles.  Per step every rank sorts its particles into one segment per
destination rank inside a peer-mapped block (cg_halo_publish) and pulls, over NVLink peer memory, only the particles of the
cell layers within one cut-off of its block of cell rows (cg_halo_pull); then it bins what it received and evaluates its
block.  No collective moves particle data; CG_EXCHANGE=peer pulls everything, =nccl uses the library's all-gather.
Total work is fixed -> "scaling": "strong".

`per_config` carries the other BASELINE configs (1, 3, 4, 5) at full size, measured the same way on the same ranks: step and
pair-kernel times, pairs/s, the F_alg fraction of the FP64 peak, and `parity` = the worst normalised error (tolerance 1e-12 ->
1.0) of the forces / fields of the particles of an i-slice that this rank evaluated, against the CPU oracle (oracle/_ref =
the unmodified reference when present).  At N > 1 every rank checks particles of its own shard, i.e. what the halo exchange
delivered.  The oracle is only the checker here and the CPU baseline; nothing timed on the GPU touches it.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

INF = float("inf")
NOMINAL_FP64_TFLOPS = 148 * 64 * 2 * 1.965e9 / 1e12  # 37.2: 148 SMs x 64 DFMA/clk x 2 x 1.965 GHz
TOL = 1e-12


def configs():
    """BASELINE.json configs as SURVEY.md 8d concretises them.  falg = algorithmic flops per ordered in-cutoff pair."""
    import coulombgalore_b200 as cg
    from oracle import oracle as O
    ION, DIP, MP = cg.MODE_ION, cg.MODE_DIPOLE, cg.MODE_MULTIPOLE
    E, F, FLD = cg.ENERGY, cg.FORCE, cg.FIELD
    return {
        "1": dict(title="configs[0]: qPotential(14, order 3) ion-ion energy + force, N = 16^3 = 4096, a = 1.75, open box", n_side=16, a=1.75,
                  seed=1001, pbc=False, mode=ION, what=E | F, cutoff=14.0,
                  schemes=[("qpotential3", lambda: cg.qPotential(14.0, 3), lambda: O.params(O.QPOTENTIAL, cutoff=14.0, order=3), 42.0)]),
        "2": dict(title="configs[1]: Wolf(12, 0.25) + Fennell(12, 0.25) ion-ion forces, N = 100^3, a = 3, periodic", n_side=100, a=3.0,
                  seed=1002, pbc=True, mode=ION, what=F, cutoff=12.0,
                  schemes=[("wolf", lambda: cg.Wolf(12.0, 0.25), lambda: O.params(O.WOLF, cutoff=12.0, alpha=0.25), 33.0),
                           ("fennell", lambda: cg.Fennell(12.0, 0.25), lambda: O.params(O.FENNELL, cutoff=12.0, alpha=0.25), 39.0)]),
        "3": dict(title="configs[2]: ReactionField(12, 80, 1) dipole-dipole energy + field + force, N = 64^3 = 262144 point dipoles, periodic",
                  n_side=64, a=3.0, seed=1003, pbc=True, mode=DIP, what=E | FLD | F, cutoff=12.0,
                  schemes=[("reactionfield", lambda: cg.ReactionField(12.0, 80.0, 1.0, False),
                            lambda: O.params(O.REACTIONFIELD, cutoff=12.0, eps_rf=80.0, eps_r=1.0, shifted=False), 101.0)]),
        "4": dict(title="configs[3]: Ewald(10, 0.3, debye 30) real-space multipole (ion + dipole) energy + force, N = 160^3 = 4096000, periodic",
                  n_side=160, a=3.0, seed=1004, pbc=True, mode=MP, what=E | F, cutoff=10.0,
                  schemes=[("ewald_screened", lambda: cg.Ewald(10.0, 0.3, INF, 30.0),
                            lambda: O.params(O.EWALD, cutoff=10.0, alpha=0.3, eps_sur=INF, debye_length=30.0), 133.0)]),
        "5": dict(title="configs[4]: Splined Poisson(12, C=4, D=3) multipole (ion + dipole) energy + force, N = 256^3 = 16777216, periodic",
                  n_side=256, a=3.0, seed=1005, pbc=True, mode=MP, what=E | F, cutoff=12.0,
                  schemes=[("splined_poisson43", lambda: cg.Splined().spline(cg.Poisson, 12.0, 4, 3),
                            lambda: O.params(O.POISSON, cutoff=12.0, C_=4, D=3, splined=True), 153.0)]),
    }


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n-side", type=int, default=0, help="lattice sites per edge of the headline config (default: BASELINE)")
    ap.add_argument("--configs", default="1,3,4,5", help="per_config entries besides the headline config 2 ('' = none)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [s.strip() for s in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if f[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def cpu_reference_rate(kind, cfg, target_seconds, nthreads=0):
    """The reference's own classes looped over the same pairs on the host cores (oracle/harness.hpp), on a bounded sample
    of i-particles; -> (pairs/s, cores, sample description, kind string).  The time includes the oracle's serial cell-list
    build over all N particles, a small handicap against this baseline that shrinks with the sample."""
    from oracle import oracle as O
    n_side = cfg["n_side"]
    x, y, z, q, _ = O.generate(n_side, cfg["a"], cfg["seed"], kind)
    L = n_side * cfg["a"]
    cores = nthreads or (os.cpu_count() or 1)
    schemes = [O.Scheme(mk(), kind) for _, _, mk, _ in cfg["schemes"]]
    N = x.size

    def run(m):
        t0 = time.perf_counter()
        pairs = 0
        for s in schemes:
            r = s.batched(x, y, z, q, None, [L] * 3, True, O.MODE_ION, O.FORCE, 0, m, cores, scales=False)
            pairs += r["pairs"]
        return pairs, time.perf_counter() - t0

    m = min(N, 16384)
    pairs, dt = run(m)  # probe
    m2 = int(min(N, max(m, m * target_seconds / max(dt, 1e-3))))
    if m2 > m:
        pairs, dt = run(m2)
        m = m2
    cpu_reference_rate.last_seconds = dt
    return pairs / dt, cores, ("first %d of %d i-particles, Wolf + Fennell force, %.1f s incl. the serial cell-list build over all N "
                               "particles" % (m, N, dt)), schemes[0].kind


class Workload:
    """One BASELINE config on this rank: inputs on the device, evaluators, the exchange step for N > 1."""

    def __init__(self, name, cfg, rank, world, local, dev, stream, n_side=0):
        import torch
        import coulombgalore_b200 as cg
        from coulombgalore_b200.distributed import HaloExchange, PeerExchange, ShardedStep, owned_range
        self.cg, self.torch = cg, torch
        self.name, self.cfg, self.rank, self.world, self.dev = name, cfg, rank, world, dev
        self.n_side = n_side or cfg["n_side"]
        self.N = self.n_side ** 3
        self.L = self.n_side * cfg["a"]
        self.box = [self.L] * 3
        mode = cfg["mode"]
        self.has_q, self.has_mu = mode != cg.MODE_DIPOLE, mode != cg.MODE_ION
        self.names = [s[0] for s in cfg["schemes"]]
        self.evs = {s[0]: cg.BatchedEvaluator(s[1](), local) for s in cfg["schemes"]}
        for ev in self.evs.values():
            ev.set_stream(stream.cuda_stream)
        self.lead = self.evs[self.names[0]]
        self.lead.set_shard(rank, world)
        for nm in self.names[1:]:  # same configuration, same cut-off: one binning + sort per step serves all schemes
            self.evs[nm].share_system(self.lead)
        full = [torch.empty(self.N, dtype=torch.float64, device=dev) for _ in range(7)]
        # synthetic inputs generated on the device by the SURVEY.md 8d rule (bit-identical to the oracle generator)
        self.lead.generate_lattice_device(self.n_side, cfg["a"], cfg["seed"], full[0], full[1], full[2], full[3], (full[4], full[5], full[6]))
        torch.cuda.synchronize()
        self.use = [0, 1, 2] + ([3] if self.has_q else []) + ([4, 5, 6] if self.has_mu else [])
        lo, hi = owned_range(self.N, rank, world)
        self.own_lo, self.own_hi = lo, hi
        self.halo = self.exchange = None
        self.halo_reason = None
        if world > 1:
            xmode = os.environ.get("CG_EXCHANGE", "halo") if cfg["pbc"] else "nccl"  # the halo windows are periodic layers
            owned = [full[k][lo:hi].clone() for k in self.use]
            if xmode == "halo":
                self.halo, self.halo_reason = HaloExchange.create(self.lead, self.N, self.box, len(self.use), dev)
                if self.halo is None:
                    xmode = "nccl"
            if self.halo is None:
                self.exchange = ShardedStep(self.N, len(self.use), dev) if xmode == "nccl" else PeerExchange.create(self.lead, self.N, len(self.use), dev)
            self.owned = torch.stack(owned) if (self.halo is not None or isinstance(self.exchange, PeerExchange)) else owned
            self.peer = isinstance(self.exchange, PeerExchange)
            del full
            self.full = None
        else:
            self.full = full
        n_out = self.N
        self.force = torch.empty((n_out, 3), dtype=torch.float64, device=dev) if cfg["what"] & cg.FORCE else None
        self.field = torch.empty((n_out, 3), dtype=torch.float64, device=dev) if cfg["what"] & cg.FIELD else None
        self.scal = {k: torch.zeros(2, dtype=torch.float64, device=dev) for k in self.names}
        # config 2 names two schemes on one configuration: they are evaluated in one pass over the pairs where the engine can
        # (CG_FUSE_SCHEMES=0: one by one), each scheme into its own force array
        self.fused = len(self.names) > 1 and cfg["what"] == cg.FORCE and os.environ.get("CG_FUSE_SCHEMES", "1") != "0"
        self.force_k = [self.force] + [torch.empty_like(self.force) for _ in self.names[1:]] if self.fused else None
        self.ids = None   # halo: global index of every received particle
        self.nb = n_out   # rows of the output buffers the last step could write

    def _set_system(self, arrays):
        g = dict(zip(self.use, arrays))
        mu = (g[4], g[5], g[6]) if 4 in g else None
        self.lead.set_system_device(g[0], g[1], g[2], g.get(3), mu, self.box if self.cfg["pbc"] else None, self.cfg["pbc"])

    def step(self, marks=None):
        """-> own kernel launches of this step."""
        cg, launches = self.cg, 0
        fo, fe = self.force, self.field
        if self.halo is not None:
            # publish + pull + hand-over to the evaluator in one library call: the particles this rank's window holds
            self.nb = self.halo.exchange(self.lead, self.owned, has_q=self.has_q, has_mu=self.has_mu)
            self.ids = self.halo.ids()
            fo = None if fo is None else fo[: self.nb]
            fe = None if fe is None else fe[: self.nb]
            launches += 4  # cg_halo_publish (3 kernels) + cg_halo_pull
        elif self.world > 1:
            arrays = self.exchange.gather(self.owned)  # every particle of every rank
            launches += 2 if self.peer else 0  # publish copy + cg_peer_gather (NCCL's kernels are not ours)
            self._set_system(arrays)
        else:
            self._set_system([self.full[k] for k in self.use])
        if marks:
            marks[0].record()
        if self.fused:  # all schemes in one pass over the pairs (cg_eval_multi_device), each into its own outputs
            k = len(self.names)
            fk = [self.force_k[i][: self.nb] for i in range(k)]
            try:
                cg.eval_multi_device([self.evs[nm] for nm in self.names], self.cfg["mode"], self.cfg["what"], force=fk,
                                     scalars=[self.scal[nm] for nm in self.names])
                launches += self.lead.counters()["launches"]
            except cg.CgError:  # a combination the one-pass evaluation does not cover: one by one from now on
                self.fused = False
        if not self.fused:
            for nm in self.names:  # the first scheme owns (bins, sorts) the system the others borrow
                self.evs[nm].eval_device(self.cfg["mode"], self.cfg["what"], force=fo, field=fe, scalars=self.scal[nm])
                launches += self.evs[nm].counters()["launches"]
        if marks:
            marks[1].record()
        return launches

    def exchange_kind(self):
        if self.world == 1:
            return None
        if self.halo is not None:
            return "halo pull over peer memory (cg_halo_publish + cg_halo_pull): %d of %d particles received by rank %d" % (
                self.halo.check_overflow(), self.N, self.rank)
        if self.peer:
            return "peer-memory pull kernel (cg_peer_gather)"
        return "NCCL all-gather (%s)" % (getattr(self.exchange, "fallback_reason", None) or self.halo_reason or
                                         ("open boundaries" if not self.cfg["pbc"] else "CG_EXCHANGE=nccl"))

    def close(self):
        if self.halo is not None:
            self.halo.close()
        if self.exchange is not None and self.peer:
            self.exchange.close()
        for ev in self.evs.values():
            ev.close()


def time_workload(w, steps, warmup, barrier, flush_buf, sampler=None):
    """W untimed steps, then exactly `steps` timed ones (CUDA events on the launching stream, L2 flushed before each)."""
    import torch
    for _ in range(warmup):
        w.step()
    barrier()
    if sampler:
        sampler.start()
    e0 = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    e1 = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    mk = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(steps)]
    pair_ms = {k: [] for k in w.names}
    sort_ms, launches = [], 0
    barrier()
    # The host enqueues all K steps without waiting for any of them, as a simulation loop would: a step is then timed by its
    # own pair of events on the stream (the L2 flush between two steps stays outside), and the ranks of a multi-GPU run meet
    # only where the exchange makes them, not at a host synchronisation per step.
    for s in range(steps):
        flush_buf.fill_(s & 0xFF)  # evict L2 between timed iterations (untimed)
        e0[s].record()
        launches += w.step(mk[s])
        e1[s].record()
    barrier()
    kh = min(steps, 64)
    for k in w.names:
        pair_ms[k] = [t["pair_ms"] for t in w.evs[k].timings_history(kh)]
    sort_ms = [t["sort_ms"] for t in w.lead.timings_history(kh)]
    clocks = sampler.stop() if sampler else None
    t = [e0[s].elapsed_time(e1[s]) for s in range(steps)]
    pm = {k: statistics.mean(v) for k, v in pair_ms.items()}
    # schemes evaluated in one pass share one kernel launch: its time counts once
    return dict(total_ms=sum(t), pair_ms=pm, pair_total_ms=pm[w.names[0]] if w.fused else sum(pm.values()), sort_ms=statistics.mean(sort_ms),
                exchange_ms=statistics.mean(e0[s].elapsed_time(mk[s][0]) for s in range(steps)),
                eval_ms=statistics.mean(mk[s][0].elapsed_time(mk[s][1]) for s in range(steps)), launches=launches, clocks=clocks)


def parity_slice(w, okind, nthreads=0):
    """Worst normalised error of the per-particle outputs this rank computed for particles of the original-index range
    [0, m) against the CPU oracle (SURVEY.md 8d norms; 1.0 = the 1e-12 tolerance).  m grows with the rank count so that
    every rank finds about 4096 particles of its own shard in the range."""
    import numpy as np
    import torch
    import coulombgalore_b200 as cg
    from oracle import oracle as O
    cfg = w.cfg
    t0 = time.perf_counter()
    x, y, z, q, mu = O.generate(w.n_side, cfg["a"], cfg["seed"], okind)
    m = int(min(w.N, 4096 * w.world))
    out = {"slice": "original indices [0, %d)" % m, "oracle": None, "schemes": {}}
    worst = 0.0
    for idx, (nm, mk, op, _) in enumerate(cfg["schemes"]):
        osch = O.Scheme(op(), okind)
        out["oracle"] = osch.kind
        ev = w.evs[nm]
        chk = None
        if op().splined:  # same tables on both sides (the product's own generation agrees with the reference's to 1e-5 only)
            inner = cg.Poisson(12.0, 4, 3)
            chk = cg.BatchedEvaluator(cg.Splined().spline_from_tables(inner, osch.spline_tables()), w.dev.index)
            chk.set_stream(torch.cuda.current_stream().cuda_stream)
            chk.share_system(w.lead)
        # one more step with the outputs pre-set to NaN: what comes back finite is what this rank's kernels wrote
        for buf in [w.force, w.field] + (w.force_k[1:] if w.fused else []):
            if buf is not None:
                buf.fill_(float("nan"))
        w.step()
        fo = None if w.force is None else w.force[: w.nb]
        fe = None if w.field is None else w.field[: w.nb]
        if w.fused:  # every scheme wrote its own array in the step above
            fo = w.force_k[idx][: w.nb]
        elif chk is not None:
            sc = torch.zeros(2, dtype=torch.float64, device=w.dev)
            chk.eval_device(cfg["mode"], cfg["what"], force=fo, field=fe, scalars=sc)
        elif idx < len(cfg["schemes"]) - 1:  # the buffers hold the last scheme's results: evaluate this one again
            ev.eval_device(cfg["mode"], cfg["what"], force=fo, field=fe, scalars=w.scal[nm])
        torch.cuda.synchronize()
        ref = osch.batched(x, y, z, q if w.has_q else None, mu if w.has_mu else None, w.box, cfg["pbc"], cfg["mode"],
                           cfg["what"] & (cg.FORCE | cg.FIELD), 0, m, nthreads)
        if w.ids is not None:
            gid = w.ids[: w.nb].to(torch.int64)
        else:
            gid = torch.arange(w.nb, device=w.dev)
        first = fo if fo is not None else fe
        # (a j-split evaluation combines zero-filled partial buffers: rows of other shards then come back as exact zeros)
        sel = (gid < m) & torch.isfinite(first[:, 0]) & (first.abs().sum(dim=1) > 0)
        rows = torch.nonzero(sel).flatten()
        g_idx = gid[rows].cpu().numpy()
        res = {"particles": int(rows.numel())}
        for key, buf, skey in (("force", fo, "force_scale"), ("field", fe, "field_scale")):
            if buf is None:
                continue
            g = buf[rows].cpu().numpy()
            r, s = ref[key][g_idx], ref[skey][g_idx]
            den = np.maximum(np.linalg.norm(r, axis=1), s)
            den = np.where(den > 0, den, 1.0)
            err = float(np.max(np.linalg.norm(g - r, axis=1) / den) / TOL) if rows.numel() else float("nan")
            res[key] = err
            worst = max(worst, err) if err == err else float("nan")
        out["schemes"][nm] = res
        if chk is not None:
            chk.close()
    out["worst"] = worst
    out["ok"] = bool(worst <= 1.0) and all(v["particles"] > 0 for v in out["schemes"].values())
    out["seconds"] = time.perf_counter() - t0
    return out


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    from oracle import oracle as O
    okind = "ref" if O.available("ref") else "port"
    CFG = configs()
    head = CFG["2"]
    n_side = args.n_side or head["n_side"]
    N = n_side ** 3
    workload = ("configs[1]: Wolf(12,0.25) + Fennell(12,0.25) ion-ion forces, cell list, N=%d^3=%d charges, a=3, periodic, "
                "R_c=12, fp64; step = spatial sort of the configuration + one Wolf + one Fennell evaluation on it" % (n_side, N))
    config = {"workload": workload, "n_particles": N, "cutoff": head["cutoff"], "l2_policy": "L2 flushed (512 MiB write) before every timed step",
              "parallelism": ("i-group sharding x%d; exchange: every rank pulls the particles of its window over NVLink peer memory "
                              "(CG_EXCHANGE=peer: all particles; =nccl or no CUDA IPC: NCCL all-gather)" % world) if world > 1 else "single GPU"}

    # ------------------------------------------------------------------------------------------------------------
    if args.impl == "reference":
        if rank != 0:
            return 0
        per_step, step_ms = [], []
        sample = kind = ""
        cores = 0
        cfg_ref = dict(head, n_side=n_side)
        for s in range(args.warmup + args.steps):
            rate, cores, sample, kind = cpu_reference_rate(okind, cfg_ref, 4.0)
            if s >= args.warmup:
                per_step.append(rate)
                step_ms.append(1e3 * cpu_reference_rate.last_seconds)
        v = statistics.mean(per_step)
        line = {"impl": "reference", "metric": "pair interactions/s (fp64)", "value": v, "unit": "pairs/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": statistics.mean(step_ms),  # one step = one bounded sample
                "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": v, "unit": "pairs/s", "cores": cores, "kind": "reference" if kind == "reference" else "port",
                                 "sample": sample},
                "e2e": {"value": v, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------------------------------------------------
    import numpy as np
    import torch

    import coulombgalore_b200 as cg

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # one process per GPU: run (and allocate the pinned host buffers) on the CPU cores next to this rank's GPU -- the
        # launcher does not bind the ranks, and copies to a pinned buffer on the other socket run at a fraction of the rate
        try:
            import pynvml
            pynvml.nvmlInit()
            hnd = pynvml.nvmlDeviceGetHandleByIndex(local)
            ncpu = os.cpu_count() or 64
            mask = pynvml.nvmlDeviceGetCpuAffinity(hnd, (ncpu + 63) // 64)
            cpus = {64 * w_ + b for w_, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
            cpus &= os.sched_getaffinity(0)
            if cpus:
                os.sched_setaffinity(0, cpus)
        except Exception:  # noqa: BLE001 -- no NVML, no affinity information: run where the launcher put us
            pass
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(v):
        t = torch.tensor([float(v)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    def allsum(v):
        t = torch.tensor([float(v)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t.item()

    # a dedicated (non-default) stream: everything timed below is enqueued on it, and the CUDA events are recorded on it
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    flush_buf = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)

    # ---- headline: config 2, device-resident ----
    w = Workload("2", head, rank, world, local, dev, stream, n_side=n_side)
    sampler = ClockSampler(local) if rank == 0 else None
    tm = time_workload(w, args.steps, args.warmup, barrier, flush_buf, sampler)
    clocks = tm["clocks"]
    ms_per_step = allmax(tm["total_ms"]) / args.steps
    pairs_by_scheme = {k: w.scal[k][1].item() for k in w.names}  # this rank's ordered in-cutoff pairs per launch
    pairs_per_step = allsum(sum(pairs_by_scheme.values()))
    value = pairs_per_step / (ms_per_step * 1e-3)
    launches = tm["launches"]
    breakdown = {"exchange_ms": allmax(tm["exchange_ms"]), "sort_ms": allmax(tm["sort_ms"]),
                 "pair_kernels_ms": allmax(tm["pair_total_ms"]), "sort_pair_reduce_ms": allmax(tm["eval_ms"]),
                 "exchange_ms_min_rank": -allmax(-tm["exchange_ms"]), "pair_kernels_ms_min_rank": -allmax(-tm["pair_total_ms"]),
                 "note": "max over ranks of the per-rank means (min_rank: the smallest); exchange = publish + wait for the peers' blocks + "
                         "pull, sort_pair_reduce = everything after it"}
    exchange_kind = w.exchange_kind()
    head_parity = None
    if not args.no_parity:
        head_parity = parity_slice(w, okind)

    # ---- end-to-end with host buffers ----
    e2e = None
    if not args.no_e2e:
        e2e = e2e_headline(w, args, rank, world, local, dev, stream, barrier, allmax, allsum)

    peak_tf, _ = w.lead.fp64_peak_probe()
    counters = w.lead.counters()
    w.close()
    del w
    torch.cuda.empty_cache()

    # ---- the other BASELINE configs, full size, same ranks ----
    per_config = {}
    for nm in [c for c in args.configs.split(",") if c]:
        cfg = CFG[nm]
        ksteps = max(3, min(args.steps, 5 if cfg["n_side"] >= 160 else 10))
        wc = Workload(nm, cfg, rank, world, local, dev, stream)
        t = time_workload(wc, ksteps, max(3, min(args.warmup, 3)), barrier, flush_buf)
        ms = allmax(t["total_ms"]) / ksteps
        pr = allsum(sum(wc.scal[k][1].item() for k in wc.names))
        falg_flops = allsum(sum(wc.scal[k][1].item() * f for (k, _, _, f) in cfg["schemes"]))
        pk_ms = allmax(t["pair_total_ms"])
        entry = {"workload": cfg["title"], "n_particles": wc.N, "steps": ksteps, "ms_per_step": ms, "pairs_per_step": pr,
                 "pairs_per_s": pr / (ms * 1e-3), "pair_kernel_ms": pk_ms, "sort_ms": allmax(t["sort_ms"]), "exchange_ms": allmax(t["exchange_ms"]),
                 "tflops_alg_step": falg_flops / (ms * 1e-3) / 1e12, "tflops_alg_kernel": falg_flops / (pk_ms * 1e-3) / 1e12,
                 "frac_of_fp64_peak_kernel": falg_flops / (pk_ms * 1e-3) / 1e12 / (peak_tf * world) if peak_tf else None,
                 "flops_per_pair": {k: f for (k, _, _, f) in cfg["schemes"]}, "split": wc.lead.counters()["split"],
                 "exchange": wc.exchange_kind()}
        if not args.no_parity:
            p = parity_slice(wc, okind)
            oks = allsum(1.0 if p["ok"] else 0.0)
            p["ranks_ok"] = int(oks)
            p["worst_all_ranks"] = allmax(p["worst"] if p["worst"] == p["worst"] else 1e300)
            entry["parity"] = p
        per_config[nm] = entry
        wc.close()
        del wc
        torch.cuda.empty_cache()
    if head_parity is not None:
        head_parity["ranks_ok"] = int(allsum(1.0 if head_parity["ok"] else 0.0))
        head_parity["worst_all_ranks"] = allmax(head_parity["worst"] if head_parity["worst"] == head_parity["worst"] else 1e300)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (pair_kernel): FP64 pipe, measured live with CUDA events ----
    falg = {k: f for (k, _, _, f) in head["schemes"]}
    pk = {}
    for k in falg:
        ms = tm["pair_ms"][k]
        pr = pairs_by_scheme[k]
        pk[k] = {"pair_kernel_ms": ms, "pairs_per_launch": pr, "pairs_per_s_kernel": pr / (ms * 1e-3), "tflops_alg": pr * falg[k] / (ms * 1e-3) / 1e12}
    ach = sum(pairs_by_scheme[k] * falg[k] for k in falg) / (tm["pair_total_ms"] * 1e-3) / 1e12
    nsch = len(falg) if tm["pair_total_ms"] < 0.75 * sum(tm["pair_ms"].values()) else 1
    alg_bytes = (N // world) * (32 + 24 * nsch) + N * 32  # per launch: own i read + force write(s), all j read once
    hbm_peak = None
    try:
        hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    traffic = traffic_source = None
    try:  # dram bytes of one pair_kernel launch of THIS kernel version, from the committed ncu summary (world 1 only)
        tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        if world == 1 and n_side == head["n_side"]:
            traffic, traffic_source = tr["dram_bytes_per_launch"], tr["source"]
    except Exception:
        pass
    kms = tm["pair_ms"]["wolf"]
    fused = tm["pair_total_ms"] < 0.75 * sum(tm["pair_ms"].values())
    roofline = {"bound": "fp64", "kernel": ("pair_kernel<ErfcMultiSrf<2>>: Wolf and Fennell in one pass over the pairs (cg_eval_multi_device), one launch "
                                            "per step" if fused else "pair_kernel (Wolf + Fennell launches)"),
                "kernel_ms_per_step": tm["pair_total_ms"], "achieved": ach, "peak": peak_tf,
                "unit": "TFLOP/s", "frac": ach / peak_tf if peak_tf else None,
                "peak_source": "measured here: register-resident DFMA chain (cg_fp64_peak_probe); MEASURED_PEAKS.json has no fp64 entry",
                "peak_nominal": NOMINAL_FP64_TFLOPS, "frac_of_nominal": ach / NOMINAL_FP64_TFLOPS,
                "flops_per_pair": falg, "traffic": traffic, "traffic_source": traffic_source,
                "hbm": {"algorithmic_bytes_per_launch": alg_bytes, "achieved_gbs": alg_bytes / (kms * 1e-3) / 1e9,
                        "peak_gbs": hbm_peak, "frac": (alg_bytes / (kms * 1e-3) / 1e9 / hbm_peak) if hbm_peak else None,
                        "note": "shown only to establish that the path is not HBM-bound"},
                "per_scheme": pk}

    cpu = None
    if not args.no_cpu_baseline:
        rate, cores, sample, kind = cpu_reference_rate(okind, dict(head, n_side=n_side), 12.0)
        cpu = {"value": rate, "unit": "pairs/s", "cores": cores, "kind": "reference" if kind == "reference" else "port", "sample": sample}

    line = {"metric": "pair interactions/s (fp64)", "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": config, "pairs_per_step": pairs_per_step, "exchange": exchange_kind,
            "step_breakdown": breakdown, "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
            "parity": head_parity, "per_config": per_config, "counters": counters}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def e2e_headline(w, args, rank, world, local, dev, stream, barrier, allmax, allsum):
    """Config 2 through host buffers.  1 GPU: the C ABI's host entry points (cg_set_system + cg_eval_multi_begin / _end), two
    configurations in flight on two sets of handles so that uploads, pair kernels and downloads overlap.  N GPUs: every rank uploads its own
    N/R particles from pinned memory, the halo exchange brings the rest over NVLink, and the forces come back over a second
    stream (double-buffered on the device), so that the download of one step overlaps the upload and kernels of the next."""
    import numpy as np
    import torch
    import coulombgalore_b200 as cg
    from coulombgalore_b200.distributed import HaloExchange
    head, N, box = w.cfg, w.N, w.box
    pots = {s[0]: s[1]() for s in head["schemes"]}
    evh = {k: cg.BatchedEvaluator(p, local) for k, p in pots.items()}
    evh["wolf"].set_shard(rank, world)
    evh["fennell"].share_system(evh["wolf"])
    ksteps = max(1, min(args.steps, 20 if world > 1 else 10))
    halo_h = None
    if world > 1 and w.halo is not None:
        halo_h, _ = HaloExchange.create(evh["wolf"], N, box, 4, dev)
    if halo_h is not None:
        for ev in evh.values():
            ev.set_stream(stream.cuda_stream)
        own_host = w.owned.cpu().pin_memory()  # this rank's slice, pinned
        own_dev = [torch.empty_like(own_host, device=dev) for _ in range(2)]
        sub, _, _ = halo_h.gather(own_dev[0].copy_(own_host))  # sizes the receive bound
        nb, cnt = sub[0].numel(), own_host.shape[1]
        up_stream, copy_stream = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        f_dev = {k: torch.zeros((nb, 3), dtype=torch.float64, device=dev) for k in evh}              # rows in received order
        # What goes back to the host: the forces of the particles THIS rank evaluated (its block of cell rows), compacted by
        # cg_collect_evaluated together with their global indices -- every particle of the system comes back from exactly one
        # rank, whoever owns it.  Rows: 32 per i-group of the shard (sized after one evaluation, with room for the sync-free bound).
        halo_h.set_system(evh["wolf"], sub, has_q=True, has_mu=False)
        cg.eval_multi_device(list(evh.values()), cg.MODE_ION, cg.FORCE, force=[f_dev[k] for k in evh]) if w.fused else \
            evh["wolf"].eval_device(cg.MODE_ION, cg.FORCE, force=f_dev["wolf"])
        torch.cuda.synchronize()
        rows = evh["wolf"].collect_rows_bound()
        rows_exact = 32 * evh["wolf"].counters()["groups"]
        rows += rows // 4 + 1024
        fo_dev = [{k: torch.zeros((rows, 3), dtype=torch.float64, device=dev) for k in evh} for _ in range(2)]
        id_dev = [torch.full((rows,), -1.0, dtype=torch.float64, device=dev) for _ in range(2)]
        f_host = [{k: torch.empty((rows, 3), dtype=torch.float64).pin_memory() for k in evh} for _ in range(2)]
        id_host = [torch.empty((rows,), dtype=torch.float64).pin_memory() for _ in range(2)]
        sc_dev = [{k: torch.zeros(2, dtype=torch.float64, device=dev) for k in evh} for _ in range(2)]
        sc_host = [{k: torch.empty(2, dtype=torch.float64).pin_memory() for k in evh} for _ in range(2)]
        uploaded = [torch.cuda.Event() for _ in range(2)]   # inputs of the step are on the device (upload stream)
        consumed = [torch.cuda.Event() for _ in range(2)]   # the exchange has read them (compute stream)
        done = [torch.cuda.Event() for _ in range(2)]       # kernels finished (compute stream)
        copied = [torch.cuda.Event() for _ in range(2)]     # download finished (copy stream)
        used0 = 32 * (rows_exact // 32 * 51 // 50 + 8)  # what a step downloads (see step_host)
        h2d_bytes, d2h_bytes = own_host.numel() * 8, 2 * (used0 * 3 * 8 + 16) + used0 * 8
        note = ("pinned host buffers; every rank uploads its own N/R particles (upload stream, double-buffered), the halo exchange brings "
                "the rest over NVLink, cg_collect_evaluated compacts the forces of the particles the rank evaluated (its block of cell "
                "rows) together with their global indices, and they come back on a third stream, double-buffered: upload, kernels and "
                "download of consecutive steps overlap; after the run the indices the ranks received back are checked to cover every "
                "particle exactly once")
        state = {"s": 0}

        phase_ev = []  # per step: stream events at the start, after the upload, after the exchange, after the kernels; copy stream: downloaded

        pool = [[torch.cuda.Event(enable_timing=True) for _ in range(5)] for _ in range(ksteps + 8)]  # created outside the timed loop

        def step_host():
            b = state["s"] & 1
            pe = pool[state["s"] % len(pool)]
            state["s"] += 1
            phase_ev.append(pe)
            with torch.cuda.stream(up_stream):
                up_stream.wait_event(consumed[b])  # the exchange of two steps ago has read this buffer
                own_dev[b].copy_(own_host, non_blocking=True)
                uploaded[b].record(up_stream)
            pe[0].record(stream)
            stream.wait_event(uploaded[b])
            pe[1].record(stream)
            halo_h.exchange(evh["wolf"], own_dev[b], has_q=True, has_mu=False)
            consumed[b].record(stream)
            pe[2].record(stream)
            if w.fused:
                cg.eval_multi_device(list(evh.values()), cg.MODE_ION, cg.FORCE, force=[f_dev[k] for k in evh],
                                     scalars=[sc_dev[b][k] for k in evh])
            else:
                for k, ev in evh.items():
                    ev.eval_device(cg.MODE_ION, cg.FORCE, force=f_dev[k], scalars=sc_dev[b][k])
            stream.wait_event(copied[b])  # the download of two steps ago has left this buffer
            evh["wolf"].collect_evaluated([f_dev[k] for k in evh], [fo_dev[b][k] for k in evh], id_dev[b], ids=halo_h.ids())
            done[b].record(stream)
            # rows that can hold a particle: 32 per i-group of the previous evaluation (+2 %: the configuration moves slowly);
            # delivered() below verifies that nothing was cut off
            ng = evh["wolf"].counters()["groups"]
            used = min(rows, 32 * (ng + ng // 50 + 8))
            state["used"] = used
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(done[b])
                id_host[b][:used].copy_(id_dev[b][:used], non_blocking=True)
                for k in evh:
                    f_host[b][k][:used].copy_(fo_dev[b][k][:used], non_blocking=True)
                    sc_host[b][k].copy_(sc_dev[b][k], non_blocking=True)
                copied[b].record(copy_stream)
            pe[3].record(stream)
            pe[4].record(copy_stream)

        def drain():
            copy_stream.synchronize()
            stream.synchronize()
            return int(sum(v[1].item() for v in sc_host[(state["s"] - 1) & 1].values()))

        def delivered():
            """-> (particles whose forces reached this rank's host buffers in the last step, all of them distinct and finite)"""
            b = (state["s"] - 1) & 1
            used = state["used"]
            ids = id_host[b].numpy()[:used]
            ok = ids >= 0
            got = ids[ok].astype(np.int64)
            fine = np.unique(got).size == got.size and all(np.isfinite(f_host[b][k].numpy()[:used][ok]).all() for k in evh)
            return int(got.size), bool(fine)
    else:
        host = [t.cpu().numpy() for t in (w.full[:4] if w.full is not None else [])]
        if not host:  # multi-GPU without CUDA IPC: regenerate the inputs on the host side of this rank
            from oracle import oracle as O
            x, y, z, q, _ = O.generate(w.n_side, head["a"], head["seed"], "ref" if O.available("ref") else "port")
            host = [x, y, z, q]
        hx = [torch.from_numpy(a).pin_memory().numpy() for a in host]
        outs = {k: {"force": torch.empty((N, 3), dtype=torch.float64).pin_memory().numpy()} for k in evh}
        h2d_bytes, d2h_bytes = 4 * 8 * N, 2 * (3 * 8 * N + 16)
        note = "cg_set_system + cg_eval_begin/_end with pinned host buffers, one handle per scheme; every rank uploads all N particles"
        last = {"pairs": 0}

        fused = w.fused and world == 1
        if fused:
            # Two sets of handles used in turn, each set on its own stream (the C ABI's default): while one configuration's
            # pair kernel runs, the next one's inputs go up and the previous one's forces come down.  Every step still uploads
            # its inputs from pinned host memory and delivers both force arrays to pinned host buffers; cg_eval_multi_end is
            # the host's wait for exactly that step.
            depth = 2  # measured (tools/e2e_probe.py): 1 -> 3.54 ms per step, 2 -> 2.20, 3 -> 2.65, 4 -> 2.63
            lanes = [dict(evs=list(evh.values()), outs=[outs[k] for k in evh])]
            for _ in range(depth - 1):
                evs = [cg.BatchedEvaluator(p, local) for p in pots.values()]
                evs[1].share_system(evs[0])
                lanes.append(dict(evs=evs, outs=[{"force": torch.empty((N, 3), dtype=torch.float64).pin_memory().numpy()} for _ in evs]))
            inflight = []
            ksteps = max(ksteps, min(args.steps, 30))
            note = ("cg_set_system (pinned host inputs) + cg_eval_multi_begin / cg_eval_multi_end (both schemes in one pass, both force "
                    "arrays into pinned host buffers); %d configurations in flight, one set of handles and one stream each" % depth)

        def finish_oldest():
            lane = inflight.pop(0)
            last["pairs"] = int(sum(r["pairs"] for r in cg.eval_multi_end(lane["evs"])))

        def step_host():
            if fused:  # one upload, one sort, one pass over the pairs for both schemes, two downloads -- per configuration
                lane = lanes[state_h["s"] % depth]
                state_h["s"] += 1
                if len(inflight) == depth:
                    finish_oldest()
                lane["evs"][0].set_system(hx[0], hx[1], hx[2], hx[3], None, box, True)
                cg.eval_multi_begin(lane["evs"], cg.MODE_ION, cg.FORCE, lane["outs"])
                inflight.append(lane)
                return
            # one handle (and stream) per scheme: the first scheme's download overlaps the second one's kernel;
            # cg_eval_end blocks until that scheme's forces are in its (pinned) host buffer
            evh["wolf"].set_system(hx[0], hx[1], hx[2], hx[3], None, box, True)  # one upload + sort for both schemes
            for k, ev in evh.items():
                ev.eval_begin(cg.MODE_ION, cg.FORCE, out=outs[k])
            last["pairs"] = sum(ev.eval_end()["pairs"] for ev in evh.values())

        state_h = {"s": 0}

        def drain():
            while fused and inflight:
                finish_oldest()
            return last["pairs"]

    for _ in range(6):  # every set of handles sizes its arrays once (the first evaluation of a shape waits for the device)
        step_host()
        drain()
    barrier()
    t0 = time.perf_counter()
    for _ in range(ksteps):
        step_host()
    t_enq = time.perf_counter() - t0  # host time to enqueue the steps (a step period below this is the host's, not the GPU's)
    pairs_last = drain()  # every step's results have reached the host buffers
    torch.cuda.synchronize()
    dt = allmax(time.perf_counter() - t0)
    pe = allsum(float(pairs_last)) * ksteps  # every step evaluates the same configuration
    out = {"value": pe / dt, "unit": "pairs/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes, "steps": ksteps, "note": note,
           "host_enqueue_ms_per_step": 1e3 * allmax(t_enq) / ksteps}
    if halo_h is not None:
        got, fine = delivered()
        total = allsum(float(got))
        out["delivered"] = {"particles_all_ranks": int(total), "expected": N, "distinct_and_finite_on_every_rank": allsum(0.0 if fine else 1.0) == 0.0}
        out["delivered"]["ok"] = int(total) == N and out["delivered"]["distinct_and_finite_on_every_rank"]
        if not out["delivered"]["ok"]:  # reported, not fatal: the line still carries the device-resident numbers
            sys.stderr.write("bench.py: e2e: the forces that came back do not cover every particle once: %r\n" % (out["delivered"],))
            out["value_is_valid"] = False
        ph = phase_ev[-ksteps:]
        mean = lambda f: statistics.mean(f(i) for i in range(len(ph)))  # noqa: E731
        out["phases_ms_rank0"] = {
            "wait_for_upload": mean(lambda i: ph[i][0].elapsed_time(ph[i][1])), "exchange": mean(lambda i: ph[i][1].elapsed_time(ph[i][2])),
            "sort_pair_reduce": mean(lambda i: ph[i][2].elapsed_time(ph[i][3])), "kernels_done_to_downloaded": mean(lambda i: ph[i][3].elapsed_time(ph[i][4])),
            "step_period": statistics.mean(ph[i][0].elapsed_time(ph[i + 1][0]) for i in range(len(ph) - 1)) if len(ph) > 1 else None}
    if halo_h is not None:
        halo_h.close()
    for lane in (lanes[1:] if halo_h is None and fused else []):
        for ev in lane["evs"]:
            ev.close()
    for ev in evh.values():
        ev.close()
    return out


if __name__ == "__main__":
    sys.exit(main())
