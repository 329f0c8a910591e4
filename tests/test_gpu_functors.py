"""Per-functor parity on a q-grid (SURVEY.md 8d extras): the device functors' S, S', S'', S''' against the oracle.

Tolerance: |delta| <= 1e-13 * max_q |S^(k)| per derivative order (absolute, relative to the function's scale: S -> 0 at
the cut-off by construction, so a purely relative bound is not meaningful there), plus a relative 1e-12 check on S itself
wherever |S| > 1e-3 * max|S| (closer to a zero of S the reference itself loses relative accuracy to cancellation).
"""
import numpy as np
import pytest

import coulombgalore_b200 as cg
from oracle import oracle as O

from parity import splined_pair, zoo

pytestmark = pytest.mark.gpu


def device_srf(pot, q):
    import torch
    ev = cg.BatchedEvaluator(pot)
    tq = torch.tensor(q, dtype=torch.float64, device="cuda")
    out = torch.empty((4, q.size), dtype=torch.float64, device="cuda")
    ev.srf_device(tq, out)
    torch.cuda.synchronize()
    res = out.cpu().numpy()
    ev.close()
    return res


def check(dev, ref):
    for k in range(4):
        scale = max(np.max(np.abs(ref[k])), 1e-300)
        err = np.max(np.abs(dev[k] - ref[k])) / scale
        assert err <= 1e-13, "derivative %d: scaled error %.3e" % (k, err)
    big = np.abs(ref[0]) > 1e-3 * np.max(np.abs(ref[0]))
    rel = np.max(np.abs(dev[0][big] - ref[0][big]) / np.abs(ref[0][big]))
    assert rel <= 1e-12, "S relative error %.3e" % rel


@pytest.mark.parametrize("name", sorted(zoo().keys()))
def test_functor_grid(name, oracle_kind):
    q = np.arange(1, 4096) / 4096.0
    mk, p = zoo()[name]
    check(device_srf(mk(), q), O.Scheme(p, oracle_kind).srf(q))


@pytest.mark.parametrize("name", ["poisson43", "wolf", "ewald_screened"])
def test_splined_functor_grid(name, oracle_kind):
    """Same tables on both sides -> the quintic Horner evaluation must agree to rounding, incl. exactly on the knots."""
    pot, osch = splined_pair(name, oracle_kind)
    knots = np.concatenate([t[0] for t in osch.spline_tables()])
    q = np.unique(np.concatenate([np.arange(1, 4096) / 4096.0, knots[(knots > 0) & (knots <= 1)]]))
    check(device_srf(pot, q), osch.srf(q))


def test_erfc_family_large_eta(oracle_kind):
    """eta = alpha * cutoff = 7.5: arguments run to the end of the device erfc table and beyond its fast range."""
    q = np.arange(1, 2048) / 2048.0
    dev = device_srf(cg.Wolf(30.0, 0.25), q)
    ref = O.Scheme(O.params(O.WOLF, cutoff=30.0, alpha=0.25), oracle_kind).srf(q)
    check(dev, ref)
    dev = device_srf(cg.Ewald(30.0, 0.3, float("inf"), 10.0), q)  # zeta = 3: negative arguments at small q
    ref = O.Scheme(O.params(O.EWALD, cutoff=30.0, alpha=0.3, debye_length=10.0), oracle_kind).srf(q)
    check(dev, ref)
