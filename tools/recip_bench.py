#!/usr/bin/env python
"""Timing of the reciprocal-space Ewald kernels (cg_recip_*), device-resident inputs, CUDA events.
usage: python tools/recip_bench.py [n_side ...]     (particles = n_side^3, a = 3, kc = 11, alpha = 0.3, Rc = 10)"""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import coulombgalore_b200 as cg


def main():
    sides = [int(a) for a in sys.argv[1:]] or [32, 64, 100]
    dev = torch.device("cuda", 0)
    st = torch.cuda.Stream()
    torch.cuda.set_stream(st)
    gen = cg.BatchedEvaluator(cg.Ewald(10.0, 0.3))
    gen.set_stream(st.cuda_stream)
    for n in sides:
        N, L = n ** 3, 3.0 * n
        a = [torch.empty(N, dtype=torch.float64, device=dev) for _ in range(7)]
        gen.generate_lattice_device(n, 3.0, 1004, a[0], a[1], a[2], a[3], (a[4], a[5], a[6]))
        pot = cg.ReciprocalEwaldGaussian(cg.ReciprocalEwaldState([L] * 3, 10.0, 11, 0.3, float("inf"), 1.0 / 30.0))
        pot.set_stream(st.cuda_stream)
        K = pot.numKVectors()
        f = torch.empty((N, 3), dtype=torch.float64, device=dev)
        e = torch.empty((N, 3), dtype=torch.float64, device=dev)
        en = torch.zeros(1, dtype=torch.float64, device=dev)
        tu, tf = [], []
        for it in range(6):
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            ev[0].record()
            pot.update_device(a[0], a[1], a[2], a[3], (a[4], a[5], a[6]), sign=0)
            pot.energy_device(en)
            ev[1].record()
            pot.force_field_device(a[0], a[1], a[2], a[3], (a[4], a[5], a[6]), force=f, field=e)
            ev[2].record()
            ev[2].synchronize()
            if it >= 2:
                tu.append(ev[0].elapsed_time(ev[1]))
                tf.append(ev[1].elapsed_time(ev[2]))
        u, ff = statistics.median(tu), statistics.median(tf)
        # complex multiply-adds actually issued per (particle, k): ~15 FP64 instructions (Q), ~21 (force + field)
        print("N=%8d K=%d  update+energy %.3f ms (%.3e particle-k/s)  force+field %.3f ms (%.3e particle-k/s)  U_rec=%.6f" %
              (N, K, u, N * K / u * 1e3, ff, N * K / ff * 1e3, en.item()), flush=True)
        pot.close()


if __name__ == "__main__":
    main()
