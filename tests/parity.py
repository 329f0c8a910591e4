"""Parity helpers shared by the GPU tests, smoke() and bench.py: scheme zoo, tolerance norms.

fp64 bar: per-particle vectors  |delta|_2 <= TOL * max(|ref|_2, sum_j |term_ij|_2);  scalars |delta| <= TOL * sum |terms|;
whole arrays |delta|_F / |ref|_F <= TOL;  in-cutoff ordered pair counts identical.  TOL = 1e-12 (BASELINE.json north_star).
"""
import numpy as np

import coulombgalore_b200 as cg
from oracle import oracle as O

TOL = 1e-12
TOL_FP32 = 1e-5  # same norms, fp32 pair arithmetic (SURVEY.md 8d "Parity tolerance")
INF = float("inf")


def zoo():
    """name -> (product scheme factory, oracle params): every scheme of the reference, incl. the core-Yukawa variants."""
    from zoo import ZOO, oracle_params
    return {k: ((lambda cls=v[1], args=v[2]: getattr(cg, cls)(*args)), oracle_params(k)) for k, v in ZOO.items()}


def splined_pair(name, kind):
    """Splined wrapper of a zoo scheme: the oracle generates the tables, both sides evaluate the same tables."""
    mk, p = zoo()[name]
    ps = O.Params.from_buffer_copy(p)
    ps.splined = 1
    osch = O.Scheme(ps, kind)
    pot = cg.Splined().spline_from_tables(mk(), osch.spline_tables())
    return pot, osch


def compare(gpu, ref, what, tol=TOL):
    """-> dict of worst normalised errors (each must be <= 1 to pass at `tol`)."""
    out = {}
    for key, skey, bit in (("force", "force_scale", cg.FORCE), ("field", "field_scale", cg.FIELD)):
        if what & bit:
            g, r, s = gpu[key], ref[key], ref[skey]
            dn = np.linalg.norm(g - r, axis=1)
            den = np.maximum(np.linalg.norm(r, axis=1), s)
            den = np.where(den > 0, den, 1.0)
            out[key + "_particle"] = float(np.max(dn / den) / tol)
            nr = np.linalg.norm(r)
            out[key + "_array"] = float(np.linalg.norm(g - r) / (nr if nr > 0 else 1.0) / tol)
    if what & cg.POTENTIAL:
        g, r, s = gpu["potential"], ref["potential"], ref["potential_scale"]
        den = np.where(s > 0, s, 1.0)
        out["potential_particle"] = float(np.max(np.abs(g - r) / den) / tol)
    if what & cg.ENERGY:
        den = ref["energy_abs"] if ref["energy_abs"] > 0 else 1.0
        out["energy"] = float(abs(gpu["energy"] - ref["energy"]) / den / tol)
    out["pairs_equal"] = 0.0 if gpu["pairs"] == ref["pairs"] else float("inf")
    return out


def run_case(pot, osch, x, y, z, q, mu, box, pbc, mode, what, shard=None, precision=None, quad=None):
    ev = cg.BatchedEvaluator(pot)
    if shard:
        ev.set_shard(*shard)
    if precision:
        ev.set_precision(precision)
    ev.set_system(x, y, z, q, mu, box, pbc)
    if quad is not None:
        ev.set_quadrupoles(quad)
    g = ev.eval(mode, what)
    g["counters"] = ev.counters()
    g["timings"] = ev.timings()
    ev.close()
    r = osch.batched(x, y, z, q, mu, box, pbc, cg.MODE_MULTIPOLE if quad is not None else mode, what, quad=quad)
    return g, r
