"""Accuracy of the device special functions (include/coulombgalore/detail/specfun.hpp) against mpmath.

The kernels evaluate erfc(x) and exp(-x^2) from a 1024-entry table plus a short polynomial / Hermite-rule correction
(erfc_gauss_table).  They are reached here through the per-functor probe cg_srf_device with an unscreened Ewald scheme of
cut-off 1 and alpha = 8:  S(q) = erfc(8 q)  and  S'(q) = -(16 / sqrt(pi)) exp(-64 q^2)  (reference ewald.hpp:175-195, zeta = 0),
so q in (0, 1) sweeps x over the whole table range (0, 8).
"""
import numpy as np
import pytest

mp = pytest.importorskip("mpmath")


@pytest.mark.gpu
def test_device_erfc_and_gaussian_against_mpmath():
    import torch
    import coulombgalore_b200 as cg
    mp.mp.dps = 40
    eta = 8.0
    ev = cg.BatchedEvaluator(cg.Ewald(1.0, eta))
    rng = np.random.default_rng(5)
    # table nodes, their neighbourhoods (one ulp away), interval midpoints and random points
    nodes = np.arange(1, 1024) / 128.0
    x = np.concatenate([nodes, np.nextafter(nodes, 0.0), np.nextafter(nodes, 9.0), nodes - 1.0 / 256.0, rng.uniform(1e-3, 8.0, 6000)])
    x = np.unique(x[(x > 0.0) & (x < 8.0)])
    q = torch.tensor(x / eta, dtype=torch.float64, device="cuda")
    out = torch.empty(4 * q.numel(), dtype=torch.float64, device="cuda")
    ev.srf_device(q, out)
    torch.cuda.synchronize()
    res = out.cpu().numpy().reshape(4, -1)
    xq = q.cpu().numpy() * eta  # the argument the device saw (x / eta * eta rounds once more)
    erfc_dev = res[0]
    gauss_dev = -res[1] * np.sqrt(np.pi) / (2.0 * eta)
    erfc_ref = np.array([float(mp.erfc(mp.mpf(float(v)))) for v in xq])
    gauss_ref = np.array([float(mp.exp(-mp.mpf(float(v)) ** 2)) for v in xq])
    rel_e = np.abs(erfc_dev - erfc_ref) / erfc_ref
    rel_g = np.abs(gauss_dev - gauss_ref) / gauss_ref
    ulp = 2.0 ** -52
    lo, mid = xq < 4.0, (xq >= 4.0) & (xq < 6.0)
    # x = eta q carries the rounding of the product (relative 2^-53): an argument error dx shows as 2 x dx / 1 in both functions
    slack = 2.0 * xq * xq * ulp
    assert np.all(rel_e[lo] <= 4 * ulp + slack[lo]), rel_e[lo].max() / ulp
    assert np.all(rel_g[lo] <= 6 * ulp + slack[lo]), rel_g[lo].max() / ulp   # + the sqrt(pi) / (2 eta) rescaling above
    assert np.all(rel_e[mid] <= 1e-14 + slack[mid]) and np.all(rel_g[mid] <= 1e-14 + slack[mid])
    hi = (xq >= 6.0) & (xq < 7.99)
    assert np.all(rel_e[hi] <= 1e-13 + slack[hi]) and np.all(rel_g[hi] <= 1e-13 + slack[hi])
    # the last half step of the table (x >= 8 - 1/256) rounds to the node past its end: only the magnitude is right there
    top = xq >= 7.99
    assert np.all(erfc_dev[top] < 2e-28) and np.all(gauss_dev[top] < 2e-27) and np.all(erfc_dev[top] > 0)
    ev.close()
