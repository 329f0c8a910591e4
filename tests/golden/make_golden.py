#!/usr/bin/env python
"""Regenerates tests/golden/reference_vectors.npz from the UNMODIFIED reference (oracle/_ref/libcg_ref.so, i.e.
/root/reference/include compiled by oracle/Makefile).  Run in the build container, where /root/reference exists:

    make -C oracle ref && python tests/golden/make_golden.py

The fixture lets machines without the reference (the GPU box, CI) pin the oracle restatement and the product to the
reference's own double-precision outputs:
  srf/<scheme>          S, S', S'', S''' on q = k/256, k = 1..255 (+ q = 1)
  pair/<scheme>         all 35 per-pair outputs (cgo_api.h) for 24 seeded pairs incl. beyond the cut-off
  info/<scheme>         cutoff, debye length, self-energy / self-field prefactor probes, dielectric, neutralisation, scheme id
  tab/<scheme>/<k>/..   Andrea tables of the splined schemes (knots, coefficients)
  batch/<case>/..       batched loop (oracle/harness.hpp) on a 6^3 lattice: force, field, potential, energy, pairs
  lattice/..            generator output (splitmix64 rule of SURVEY.md 8d)
  batch_quad/<case>/..  the same loop in multipole mode with seeded, non-symmetric quadrupole tensors (quad_inputs)
  recip/..              reciprocal-space Ewald (ReciprocalEwaldGaussian) on the lattice: k-vectors, Aks, Q, energy, force,
                        field, force scale, surface terms for recip_params
  self/<scheme>/..      O(N) terms on the lattice with a non-neutral charge set (self_inputs): scalars = sum of
                        self_energy, its absolute sum, neutralization_energy (V = 18^3), charge sum; self_field per
                        particle; self/torque = dipole_torque(mu_i, E_i) for the seeded field of self_inputs
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import oracle as O  # noqa: E402
from zoo import ZOO, oracle_params  # noqa: E402

SPLINED = ["poisson43", "wolf", "ewald_screened", "plain", "qpotential3", "reactionfield", "plain_yukawa"]


def pair_inputs(cutoff):
    rng = np.random.default_rng(20240611)
    rc = cutoff if cutoff < 1e10 else 20.0
    out = []
    for t in range(24):
        r = rng.normal(size=3)
        r *= rng.uniform(0.05, 1.08) * rc / np.linalg.norm(r)
        out.append((float(rng.uniform(-2, 2)), float(rng.uniform(-2, 2)), rng.normal(size=3), rng.normal(size=3), r,
                    rng.normal(size=(3, 3)), rng.normal(size=(3, 3))))
    return out


def quad_inputs(n):
    """Seeded general (non-symmetric, not traceless) 3x3 quadrupole tensors."""
    return np.random.default_rng(4242).normal(size=(n, 3, 3))


# cutoff, alpha, eps_sur, kappa, reciprocal_cutoff, box
RECIP_CASES = {"plain": (9.0, 0.35, 80.0, 0.0, 7, [18.0, 18.0, 18.0]), "screened_ortho": (8.0, 0.4, 1.0, 0.05, 5, [18.0, 20.0, 22.0])}


def self_inputs(q):
    """Charges with a net charge (so that the neutralisation term is exercised) and a seeded field for the torque."""
    qn = q.copy()
    qn[:10] = 1.7
    field = np.random.default_rng(77).normal(size=(q.size, 3))
    return qn, field


def main():
    assert O.available("ref"), "oracle/_ref/libcg_ref.so missing: run `make -C oracle ref` where /root/reference exists"
    d = {}
    q = np.concatenate([np.arange(1, 256) / 256.0, [1.0]])
    d["srf_q"] = q
    for name in ZOO:
        for spl in ([False, True] if name in SPLINED else [False]):
            key = ("splined_" if spl else "") + name
            s = O.Scheme(oracle_params(name, splined=spl), "ref")
            d["srf/" + key] = s.srf(q)
            info = s.info()
            d["info/" + key] = np.array([info[k] for k in ("cutoff", "debye_length", "self_energy_ion", "self_energy_dipole",
                                                            "self_field_dipole", "dielectric_m2v1", "neutralization_q1_v1", "scheme")], dtype=float)
            rows = [s.pair(zA, zB, a, b, r, qa, qb)["raw"] for (zA, zB, a, b, r, qa, qb) in pair_inputs(info["cutoff"])]
            d["pair/" + key] = np.array(rows)
            if spl:
                for k, (r2, c) in enumerate(s.spline_tables()):
                    d["tab/%s/%d/r2" % (name, k)] = r2
                    d["tab/%s/%d/c" % (name, k)] = c
    x, y, z, ch, mu = O.generate(6, 3.0, 31337, "ref")
    d["lattice/n6_a3_seed31337"] = np.stack([x, y, z, ch, mu[0], mu[1], mu[2]])
    ALL = O.ENERGY | O.FORCE | O.FIELD | O.POTENTIAL
    for name, mode, pbc in (("wolf", O.MODE_ION, True), ("qpotential3", O.MODE_ION, False), ("reactionfield", O.MODE_DIPOLE, True),
                            ("ewald_screened", O.MODE_MULTIPOLE, True), ("poisson33_yukawa", O.MODE_MULTIPOLE, True),
                            ("plain", O.MODE_MULTIPOLE, False)):
        s = O.Scheme(oracle_params(name), "ref")
        L = 18.0
        if pbc and s.info()["cutoff"] > L:
            continue
        # cut-offs of the zoo (10-14) exceed L/2 = 9: the loop then sums the nearest images, which is what the product does too
        r = s.batched(x, y, z, ch, mu, [L] * 3, pbc, mode, ALL)
        k = "batch/%s_mode%d_pbc%d/" % (name, mode, int(pbc))
        for f in ("force", "field", "potential", "force_scale", "field_scale", "potential_scale"):
            d[k + f] = r[f]
        d[k + "scalars"] = np.array([r["energy"], r["energy_abs"], float(r["pairs"])])
    quad = quad_inputs(x.size)
    for name, pbc in (("wolf", True), ("poisson33_yukawa", True), ("ewald_screened", True), ("qpotential3", False), ("plain", False)):
        r = O.Scheme(oracle_params(name), "ref").batched(x, y, z, ch, mu, [18.0] * 3, pbc, O.MODE_MULTIPOLE, ALL, quad=quad)
        k = "batch_quad/%s_pbc%d/" % (name, int(pbc))
        for f in ("force", "field", "potential", "force_scale", "field_scale", "potential_scale"):
            d[k + f] = r[f]
        d[k + "scalars"] = np.array([r["energy"], r["energy_abs"], float(r["pairs"])])
    for name, (rc, al, eps, kap, kc, box) in RECIP_CASES.items():
        r = O.Reciprocal(rc, al, eps, kap, kc, box, "ref")
        r.update(x, y, z, ch, mu, 0)
        kv, aks = r.kvectors()
        ff = r.force_field(x, y, z, ch, mu)
        se, sf = r.surface(x, y, z, ch, mu)
        d["recip/%s/k" % name], d["recip/%s/aks" % name] = kv, aks
        qq = r.q()
        d["recip/%s/q" % name] = np.stack([qq.real, qq.imag])
        d["recip/%s/force" % name], d["recip/%s/field" % name], d["recip/%s/force_scale" % name] = ff["force"], ff["field"], ff["force_scale"]
        d["recip/%s/scalars" % name] = np.array([r.energy(), se, sf[0], sf[1], sf[2]])
    qn, fld = self_inputs(ch)
    for name in ZOO:
        r = O.Scheme(oracle_params(name), "ref").self_terms(qn, mu, 18.0 ** 3, field=fld)
        d["self/%s/scalars" % name] = np.array([r["self_energy"], r["self_energy_abs"], r["neutralization_energy"], r["charge_sum"]])
        d["self/%s/self_field" % name] = r["self_field"]
        d["self/torque"] = r["torque"]
    np.savez_compressed(os.path.join(HERE, "reference_vectors.npz"), **d)
    print("wrote %d arrays, %.1f KB" % (len(d), os.path.getsize(os.path.join(HERE, "reference_vectors.npz")) / 1024))


if __name__ == "__main__":
    main()
