#!/usr/bin/env python
"""tools/e2e_probe.py -- where the host-buffer (e2e) step of config 2 spends its time: PCIe copy rates of this box, and the
pipelined cg_set_system + cg_eval_multi_begin/_end loop at several depths (configurations in flight)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import coulombgalore_b200 as cg


def main():
    dev = torch.device("cuda", 0)
    n_side, a = 100, 3.0
    N, L = n_side ** 3, n_side * a
    h_in = torch.empty(4 * N, dtype=torch.float64).pin_memory()
    h_out = torch.empty(6 * N, dtype=torch.float64).pin_memory()
    d_in, d_out = torch.empty(4 * N, dtype=torch.float64, device=dev), torch.empty(6 * N, dtype=torch.float64, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def timed(fn, reps=10):
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps * 1e3

    def up():
        with torch.cuda.stream(s1):
            d_in.copy_(h_in, non_blocking=True)

    def down():
        with torch.cuda.stream(s2):
            h_out.copy_(d_out, non_blocking=True)
    t_up, t_dn = timed(up), timed(down)
    t_both = timed(lambda: (up(), down()))
    print("H2D 32 MB %.3f ms (%.1f GB/s)   D2H 48 MB %.3f ms (%.1f GB/s)   both at once %.3f ms" % (
        t_up, 32e-3 * N / 1e6 / t_up, t_dn, 48e-3 * N / 1e6 / t_dn, t_both))

    lead = cg.BatchedEvaluator(cg.Wolf(12.0, 0.25))
    full = [torch.empty(N, dtype=torch.float64, device=dev) for _ in range(7)]
    lead.generate_lattice_device(n_side, a, 20260101, full[0], full[1], full[2], full[3], (full[4], full[5], full[6]))
    torch.cuda.synchronize()
    hx = [t.cpu().pin_memory().numpy() for t in full[:4]]
    lead.close()
    box = [L] * 3
    for depth in (1, 2, 3, 4):
        lanes = []
        for _ in range(depth):
            evs = [cg.BatchedEvaluator(cg.Wolf(12.0, 0.25)), cg.BatchedEvaluator(cg.Fennell(12.0, 0.25))]
            evs[1].share_system(evs[0])
            lanes.append((evs, [{"force": torch.empty((N, 3), dtype=torch.float64).pin_memory().numpy()} for _ in evs]))
        inflight = []

        def step(s):
            evs, outs = lanes[s % depth]
            if len(inflight) == depth:
                cg.eval_multi_end(inflight.pop(0))
            evs[0].set_system(hx[0], hx[1], hx[2], hx[3], None, box, True)
            cg.eval_multi_begin(evs, cg.MODE_ION, cg.FORCE, outs)
            inflight.append(evs)

        def drain():
            while inflight:
                pairs = sum(r["pairs"] for r in cg.eval_multi_end(inflight.pop(0)))
            return pairs
        for s in range(2 * depth):
            step(s)
            drain()
        K = 24
        t0 = time.perf_counter()
        for s in range(K):
            step(s)
        pairs = drain()
        dt = (time.perf_counter() - t0) / K
        tm = lanes[0][0][0].timings_history(4)
        print("depth %d: %.3f ms per step -> %.3e pairs/s   (pair kernel %.3f ms, sort %.3f ms)" % (
            depth, dt * 1e3, pairs / dt, tm[0]["pair_ms"], tm[0]["sort_ms"]))
        for evs, _ in lanes:
            for ev in evs:
                ev.close()


if __name__ == "__main__":
    main()
