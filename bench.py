#!/usr/bin/env python
"""bench.py -- pair interactions/s (fp64) of the B200 batched CoulombGalore engine.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1]): Wolf(12, 0.25) and Fennell(12, 0.25) ion-ion forces, N = 100^3 = 1 000 000 charges
on the jittered lattice of SURVEY.md 8d (a = 3, L = 300, periodic, seed 1002), cell list under R_c = 12, fp64.
One STEP = one Wolf evaluation + one Fennell evaluation of all per-particle forces: spatial sort, pair kernel and
reductions each time (positions are treated as new every step).  The unit of work is one ordered in-cutoff pair
(i <- j); `value` = pairs evaluated by all ranks per second with the inputs resident in HBM; `e2e` is the same metric
through the host-buffer C ABI (cg_set_system + cg_eval: H2D of positions/charges and D2H of forces inside the timing).

N > 1: every rank owns N/R particles, the step all-gathers the SoA inputs over NCCL, then each rank
sorts and evaluates its contiguous block of i-groups.  Total work is fixed -> "scaling": "strong".
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

CFG = dict(n_side=100, a=3.0, seed=1002, cutoff=12.0, alpha=0.25)
F_ALG = {"wolf": 33.0, "fennell": 39.0}  # algorithmic flops per ordered in-cutoff pair, SURVEY.md 8d / appendix D
NOMINAL_FP64_TFLOPS = 148 * 64 * 2 * 1.965e9 / 1e12  # 37.2: 148 SMs x 64 DFMA/clk x 2 x 1.965 GHz


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n-side", type=int, default=CFG["n_side"], help="lattice sites per edge (default: the BASELINE config)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
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


def cpu_reference_rate(kind, n_side, target_seconds, nthreads=0):
    """The reference's own classes looped over the same pairs on the host cores (oracle/harness.hpp), on a bounded
    sample of i-particles; -> (pairs/s, cores, sample description, kind string)."""
    from oracle import oracle as O
    x, y, z, q, _ = O.generate(n_side, CFG["a"], CFG["seed"], kind)
    L = n_side * CFG["a"]
    cores = nthreads or (os.cpu_count() or 1)
    schemes = [O.Scheme(O.params(O.WOLF, cutoff=CFG["cutoff"], alpha=CFG["alpha"]), kind),
               O.Scheme(O.params(O.FENNELL, cutoff=CFG["cutoff"], alpha=CFG["alpha"]), kind)]
    N = x.size

    def run(m):
        t0 = time.perf_counter()
        pairs = 0
        for s in schemes:
            r = s.batched(x, y, z, q, None, [L] * 3, True, O.MODE_ION, O.FORCE, 0, m, cores, scales=False)
            pairs += r["pairs"]
        return pairs, time.perf_counter() - t0

    m = min(N, 16384)
    pairs, dt = run(m)  # probe (includes the serial CPU cell-list build for all N particles)
    m2 = int(min(N, max(m, m * target_seconds / max(dt, 1e-3))))
    if m2 > m:
        pairs, dt = run(m2)
        m = m2
    cpu_reference_rate.last_seconds = dt
    return pairs / dt, cores, "first %d of %d i-particles, Wolf + Fennell force, %.1f s incl. cell-list build" % (m, N, dt), schemes[0].kind


# dram__bytes_read.sum + dram__bytes_write.sum of one Wolf pair_kernel launch (profiles/ncu_r1_pair_kernel_wolf_v5.txt)
NCU_DRAM_BYTES_PER_LAUNCH = 65406976 + 7085568


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    n_side = args.n_side
    N = n_side ** 3
    workload = ("configs[1]: Wolf(12,0.25) + Fennell(12,0.25) ion-ion forces, cell list, N=%d^3=%d charges, a=3, periodic, "
                "R_c=12, fp64; step = spatial sort of the configuration + one Wolf + one Fennell evaluation on it" % (n_side, N))
    config = {"workload": workload, "n_particles": N, "cutoff": CFG["cutoff"], "l2_policy": "L2 flushed (512 MiB write) before every timed step",
              "parallelism": ("i-group sharding x%d; exchange: every rank pulls the particles of its window over NVLink peer memory "
                              "(CG_EXCHANGE=peer: all particles; =nccl or no CUDA IPC: NCCL all-gather)" % world) if world > 1 else "single GPU"}

    from oracle import oracle as O
    okind = "ref" if O.available("ref") else "port"

    # ------------------------------------------------------------------------------------------------------------
    if args.impl == "reference":
        if rank != 0:
            return 0
        per_step, step_ms = [], []
        sample = kind = ""
        cores = 0
        for s in range(args.warmup + args.steps):
            rate, cores, sample, kind = cpu_reference_rate(okind, n_side, 4.0)
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
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    pots = {"wolf": cg.Wolf(CFG["cutoff"], CFG["alpha"]), "fennell": cg.Fennell(CFG["cutoff"], CFG["alpha"])}
    evs = {k: cg.BatchedEvaluator(p, local) for k, p in pots.items()}
    # a dedicated (non-default) stream: everything timed below is enqueued on it, and the CUDA events are recorded on it
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    for ev in evs.values():
        ev.set_stream(stream.cuda_stream)
    evs["wolf"].set_shard(rank, world)
    evs["fennell"].share_system(evs["wolf"])  # same configuration, same cut-off: one binning + sort per step serves both schemes
    L = n_side * CFG["a"]
    box = [L, L, L]

    # synthetic inputs generated on the device by the SURVEY.md 8d rule (bit-identical to the oracle generator)
    full = [torch.empty(N, dtype=torch.float64, device=dev) for _ in range(4)]
    evs["wolf"].generate_lattice_device(n_side, CFG["a"], CFG["seed"], full[0], full[1], full[2], full[3])
    torch.cuda.synchronize()
    from coulombgalore_b200.distributed import HaloExchange, PeerExchange, ShardedStep, owned_range
    lo, hi = owned_range(N, rank, world)
    owned = [t[lo:hi].clone() for t in full]  # this rank's particles
    exchange = halo = None
    if world > 1:
        # the exchange step.  halo (default): every rank pulls, over NVLink peer memory, only the particles of its window;
        # peer: it pulls all of them (both are this repo's kernels); nccl: the collective library's all-gather
        xmode = os.environ.get("CG_EXCHANGE", "halo")
        halo_reason = None
        if xmode == "halo":
            halo, halo_reason = HaloExchange.create(evs["wolf"], N, box, 4, dev)  # None where CUDA IPC is not available
            if halo is None:
                xmode = "nccl"
        if halo is None:
            exchange = ShardedStep(N, 4, dev) if xmode == "nccl" else PeerExchange.create(evs["wolf"], N, 4, dev)
        if halo is not None or isinstance(exchange, PeerExchange):
            owned = torch.stack(owned)
    gathered = exchange.full if exchange is not None else full
    force = torch.empty((N, 3), dtype=torch.float64, device=dev)
    scal = {k: torch.zeros(2, dtype=torch.float64, device=dev) for k in evs}
    flush_buf = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def step_device():
        launches = 0
        fo = force
        if halo is not None:
            sub, ids, n_out = halo.gather(owned)  # positions + charges of the particles this rank's window holds
            halo.set_system(evs["wolf"], sub, has_q=True, has_mu=False)
            fo = force[: sub[0].numel()]
            launches += 4  # cg_halo_publish (3 kernels) + cg_halo_pull
        else:
            if world > 1:
                exchange.gather(owned)  # positions + charges of every rank
                launches += 2 if isinstance(exchange, PeerExchange) else 0  # publish copy + cg_peer_gather (NCCL's kernels are not ours)
            evs["wolf"].set_system_device(gathered[0], gathered[1], gathered[2], gathered[3], None, box, True)
        for name, ev in evs.items():  # wolf first: it owns (bins, sorts) the system fennell borrows
            ev.eval_device(cg.MODE_ION, cg.FORCE, force=fo, scalars=scal[name])
            launches += ev.counters()["launches"]
        return launches

    # ---- device-resident timing ----
    pair_ms = {k: [] for k in evs}
    sort_ms = {k: [] for k in evs}
    for _ in range(args.warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    t_steps = []
    launches = 0
    e0 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    e1 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    barrier()
    for s in range(args.steps):
        flush_buf.fill_(s & 0xFF)  # evict L2 between timed iterations (untimed)
        e0[s].record()
        launches += step_device()
        e1[s].record()
        e1[s].synchronize()
        for k, ev in evs.items():
            t = ev.timings()
            pair_ms[k].append(t["pair_ms"])
            sort_ms[k].append(t["sort_ms"])
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    t_steps = [e0[s].elapsed_time(e1[s]) for s in range(args.steps)]
    total_ms = torch.tensor([sum(t_steps)], dtype=torch.float64, device=dev)
    pairs_rank = torch.tensor([float(sum(scal[k][1].item() for k in evs))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(pairs_rank, op=dist.ReduceOp.SUM)
    ms_per_step = total_ms.item() / args.steps
    pairs_per_step = pairs_rank.item()
    value = pairs_per_step / (ms_per_step * 1e-3)

    # ---- end-to-end through the host-buffer C ABI ----
    e2e = None
    if not args.no_e2e:
        host = [t.cpu().numpy() for t in full]  # cg_set_system copies from these host buffers every step
        hx = [torch.from_numpy(a).pin_memory().numpy() for a in host]
        evh = {k: cg.BatchedEvaluator(p, local) for k, p in pots.items()}
        evh["wolf"].set_shard(rank, world)
        evh["fennell"].share_system(evh["wolf"])

        halo_h = None
        if world > 1 and halo is not None:  # multi-GPU: every rank uploads only its own particles, the rest moves over NVLink
            halo_h, _ = HaloExchange.create(evh["wolf"], N, box, 4, dev)
        if halo_h is not None:
            for ev in evh.values():
                ev.set_stream(stream.cuda_stream)
            own_host = torch.from_numpy(np.stack([a[lo:hi] for a in host])).pin_memory()  # this rank's slice, pinned
            own_dev = torch.empty_like(own_host, device=dev)
            sub, _, _ = halo_h.gather(own_dev.copy_(own_host))  # sizes the receive bound
            nb = sub[0].numel()
            f_dev = {k: torch.zeros((nb, 3), dtype=torch.float64, device=dev) for k in evh}
            f_host = {k: torch.empty((nb, 3), dtype=torch.float64).pin_memory() for k in evh}
            sc_dev = {k: torch.zeros(2, dtype=torch.float64, device=dev) for k in evh}
            sc_host = {k: torch.empty(2, dtype=torch.float64).pin_memory() for k in evh}
            h2d_bytes, d2h_bytes = own_host.numel() * 8, 2 * (nb * 3 * 8 + 16)
            note = ("pinned host buffers; every rank uploads its own N/R particles, the halo exchange brings the rest over NVLink, "
                    "and downloads the forces of the particles it received (those of its shard are non-zero)")

            def step_host():
                own_dev.copy_(own_host, non_blocking=True)
                sub, _, _ = halo_h.gather(own_dev)
                halo_h.set_system(evh["wolf"], sub, has_q=True, has_mu=False)
                for k, ev in evh.items():
                    ev.eval_device(cg.MODE_ION, cg.FORCE, force=f_dev[k], scalars=sc_dev[k])
                    f_host[k].copy_(f_dev[k], non_blocking=True)
                    sc_host[k].copy_(sc_dev[k], non_blocking=True)
                stream.synchronize()
                return int(sum(v[1].item() for v in sc_host.values()))
        else:
            outs = {k: {"force": torch.empty((N, 3), dtype=torch.float64).pin_memory().numpy()} for k in evh}
            h2d_bytes, d2h_bytes = 4 * 8 * N, 2 * (3 * 8 * N + 16)
            note = "cg_set_system + cg_eval_begin/_end with pinned host buffers, one handle per scheme; every rank uploads all N particles"

            def step_host():
                # one handle (and stream) per scheme: the first scheme's download overlaps the second one's kernel;
                # cg_eval_end blocks until that scheme's forces are in its (pinned) host buffer
                evh["wolf"].set_system(hx[0], hx[1], hx[2], hx[3], None, box, True)  # one upload + sort for both schemes
                for k, ev in evh.items():
                    ev.eval_begin(cg.MODE_ION, cg.FORCE, out=outs[k])
                return sum(ev.eval_end()["pairs"] for ev in evh.values())

        step_host()
        barrier()
        t0 = time.perf_counter()
        pe = 0
        ksteps = max(1, min(args.steps, 5))
        for _ in range(ksteps):
            pe += step_host()
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        pe_t = torch.tensor([float(pe)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            dist.all_reduce(pe_t, op=dist.ReduceOp.SUM)
        e2e = {"value": pe_t.item() / dt.item(), "unit": "pairs/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
               "steps": ksteps, "note": note}
        if halo_h is not None:
            halo_h.close()
        for ev in evh.values():
            ev.close()

    exchange_kind = None
    if world > 1:
        if halo is not None:
            exchange_kind = "halo pull over peer memory (cg_halo_publish + cg_halo_pull): %d of %d particles received by rank %d" % (
                halo.check_overflow(), N, rank)
            halo.close()
        else:
            exchange_kind = "peer-memory pull kernel (cg_peer_gather)" if isinstance(exchange, PeerExchange) else \
                "NCCL all-gather (%s)" % (getattr(exchange, "fallback_reason", None) or halo_reason or "CG_EXCHANGE=nccl")
            if isinstance(exchange, PeerExchange):
                exchange.close()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (pair_kernel): FP64 pipe, measured live with CUDA events ----
    peak_tf, _ = evs["wolf"].fp64_peak_probe()
    pk = {}
    for k in evs:
        ms = statistics.mean(pair_ms[k])
        pr = scal[k][1].item()
        pk[k] = {"pair_kernel_ms": ms, "sort_ms": statistics.mean(sort_ms[k]), "pairs_per_launch": pr,
                 "pairs_per_s_kernel": pr / (ms * 1e-3), "tflops_alg": pr * F_ALG[k] / (ms * 1e-3) / 1e12}
    ach = sum(scal[k][1].item() * F_ALG[k] for k in evs) / (sum(statistics.mean(pair_ms[k]) for k in evs) * 1e-3) / 1e12
    alg_bytes = (N // world) * (32 + 24) + N * 32  # per launch: own i read + force write, all j read once
    hbm_peak = None
    try:
        hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    kms = statistics.mean(pair_ms["wolf"])
    roofline = {"bound": "fp64", "kernel": "pair_kernel (Wolf + Fennell launches)", "achieved": ach, "peak": peak_tf,
                "unit": "TFLOP/s", "frac": ach / peak_tf if peak_tf else None,
                "peak_source": "measured here: register-resident DFMA chain (cg_fp64_peak_probe); MEASURED_PEAKS.json has no fp64 entry",
                "peak_nominal": NOMINAL_FP64_TFLOPS, "frac_of_nominal": ach / NOMINAL_FP64_TFLOPS,
                "flops_per_pair": F_ALG, "traffic": NCU_DRAM_BYTES_PER_LAUNCH if world == 1 else None,
                "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one Wolf pair_kernel launch at N=10^6, ncu --set full, "
                                  "profiles/ncu_r1_pair_kernel_wolf_v5.txt (65.41 MB + 7.09 MB); algorithmic bytes per launch: 88 MB",
                "hbm": {"algorithmic_bytes_per_launch": alg_bytes, "achieved_gbs": alg_bytes / (kms * 1e-3) / 1e9,
                        "peak_gbs": hbm_peak, "frac": (alg_bytes / (kms * 1e-3) / 1e9 / hbm_peak) if hbm_peak else None,
                        "note": "shown only to establish that the path is not HBM-bound"},
                "per_scheme": pk}

    cpu = None
    if not args.no_cpu_baseline:
        rate, cores, sample, kind = cpu_reference_rate(okind, n_side, 12.0)
        cpu = {"value": rate, "unit": "pairs/s", "cores": cores, "kind": "reference" if kind == "reference" else "port", "sample": sample}

    line = {"metric": "pair interactions/s (fp64)", "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": config, "pairs_per_step": pairs_per_step, "exchange": exchange_kind, "e2e": e2e,
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
            "counters": evs["wolf"].counters()}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
