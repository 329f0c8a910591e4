"""The all-pairs path (allpairs_kernel.cuh: small open systems, no sort, j-tiles staged by TMA bulk copies) against the CPU
oracle and against the cell-list path on the same inputs."""
import os

import numpy as np
import pytest

import coulombgalore_b200 as cg
from oracle import oracle as O

from parity import TOL, compare, run_case, splined_pair, zoo

pytestmark = pytest.mark.gpu

ALL = cg.ENERGY | cg.FORCE | cg.FIELD | cg.POTENTIAL


def assert_parity(errs):
    bad = {k: v for k, v in errs.items() if not (v <= 1.0)}
    assert not bad, "parity beyond %.0e (normalised errors, must be <= 1): %r" % (TOL, errs)


class env:
    def __init__(self, **kv):
        self.kv = kv

    def __enter__(self):
        self.old = {k: os.environ.get(k) for k in self.kv}
        for k, v in self.kv.items():
            os.environ[k] = str(v)

    def __exit__(self, *a):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


@pytest.mark.parametrize("name", sorted(zoo().keys()))
@pytest.mark.parametrize("mode", [cg.MODE_ION, cg.MODE_DIPOLE, cg.MODE_MULTIPOLE])
def test_every_scheme_every_mode_all_pairs(name, mode, oracle_kind):
    """Every scheme x {ion, dipole, multipole} x all outputs on an open 11^3 system (odd count: the tail of the bulk copies)."""
    mk, p = zoo()[name]
    x, y, z, q, mu = O.generate(11, 2.6, 300 + mode, oracle_kind)
    g, r = run_case(mk(), O.Scheme(p, oracle_kind), x, y, z, q, mu, [30.0] * 3, False, mode, ALL)
    assert g["counters"]["launches"] == 2 and g["counters"]["cells"] == 0, g["counters"]  # pair kernel + epilogue, no sort
    assert_parity(compare(g, r, ALL))


@pytest.mark.parametrize("n,slices", [(2, 0), (33, 0), (63, 0), (500, 1), (1500, 1), (1501, 1), (2197, 2), (4096, 0), (8192, 0)])
def test_sizes_tiles_and_slices(n, slices, oracle_kind):
    """Particle counts around the tile / slice / granule boundaries; CG_AP_SLICES forces slices of several tiles, i.e. the
    re-use of the two stages (1500 particles in one slice = three tiles)."""
    rng = np.random.default_rng(n)
    L = 14.0 * max(1.0, (n / 500.0) ** (1.0 / 3.0))
    pos = rng.random((n, 3)) * L
    x, y, z = (np.ascontiguousarray(pos[:, k]) for k in range(3))
    q = rng.choice([-1.0, 1.0], n)
    mk, p = zoo()["qpotential3"]
    kv = {"CG_AP_SLICES": slices} if slices else {}
    with env(**kv):
        g, r = run_case(mk(), O.Scheme(p, oracle_kind), x, y, z, q, None, [L] * 3, False, cg.MODE_ION, cg.ENERGY | cg.FORCE)
    assert g["counters"]["cells"] == 0
    if slices:
        assert g["counters"]["split"] == slices
    assert_parity(compare(g, r, cg.ENERGY | cg.FORCE))


def test_all_pairs_equals_cell_list(oracle_kind):
    """Both paths on one system: identical pair counts, outputs equal to rounding (different summation order)."""
    x, y, z, q, mu = O.generate(14, 2.2, 515, oracle_kind)
    w = cg.ENERGY | cg.FORCE | cg.FIELD
    res = {}
    for key, kv in (("ap", {}), ("cells", {"CG_ALLPAIRS_MAX": 0})):
        with env(**kv):
            ev = cg.BatchedEvaluator(cg.ReactionField(12, 80, 1, False))
            ev.set_system(x, y, z, None, mu, [31.0] * 3, False)
            res[key] = ev.eval(cg.MODE_DIPOLE, w)
            res[key]["counters"] = ev.counters()
            ev.close()
    assert res["ap"]["counters"]["cells"] == 0 and res["cells"]["counters"]["cells"] > 0
    assert res["ap"]["pairs"] == res["cells"]["pairs"]
    scale = np.abs(res["cells"]["force"]).max()
    assert np.abs(res["ap"]["force"] - res["cells"]["force"]).max() <= 1e-11 * scale
    assert abs(res["ap"]["energy"] - res["cells"]["energy"]) <= 1e-11 * abs(res["cells"]["energy"])


def test_two_schemes_one_pass_all_pairs(oracle_kind):
    """cg_eval_multi_device on a small open system: Wolf and Fennell in one all-pairs launch, each against the oracle."""
    import torch
    x, y, z, q, mu = O.generate(12, 2.8, 909, oracle_kind)
    dev = torch.device("cuda", 0)
    evs = [cg.BatchedEvaluator(cg.Wolf(12, 0.25)), cg.BatchedEvaluator(cg.Fennell(12, 0.25))]
    evs[1].share_system(evs[0])
    evs[0].set_system(x, y, z, q, None, [34.0] * 3, False)
    n = len(x)
    f = [torch.zeros((n, 3), dtype=torch.float64, device=dev) for _ in evs]
    sc = [torch.zeros(2, dtype=torch.float64, device=dev) for _ in evs]
    cg.eval_multi_device(evs, cg.MODE_ION, cg.ENERGY | cg.FORCE, force=f, scalars=sc)
    torch.cuda.synchronize()
    assert evs[0].counters()["cells"] == 0
    for k, name in enumerate(("wolf", "fennell")):
        r = O.Scheme(zoo()[name][1], oracle_kind).batched(x, y, z, q, None, [34.0] * 3, False, cg.MODE_ION, cg.ENERGY | cg.FORCE)
        g = {"force": f[k].cpu().numpy(), "energy": float(sc[k][0]), "pairs": int(sc[k][1].item() + 0.5)}
        assert_parity(compare(g, r, cg.ENERGY | cg.FORCE))
    for ev in evs:
        ev.close()


def test_splined_and_unbounded_cutoff_all_pairs(oracle_kind):
    x, y, z, q, mu = O.generate(9, 3.0, 4141, oracle_kind)
    pot, osch = splined_pair("poisson43", oracle_kind)
    w = cg.ENERGY | cg.FORCE
    g, r = run_case(pot, osch, x, y, z, q, mu, [27.0] * 3, False, cg.MODE_MULTIPOLE, w)
    assert g["counters"]["cells"] == 0
    assert_parity(compare(g, r, w))
    mk, p = zoo()["plain"]  # no cut-off at all: every pair passes, the dummies are told apart by their index
    g, r = run_case(mk(), O.Scheme(p, oracle_kind), x, y, z, q, mu, [27.0] * 3, False, cg.MODE_ION, ALL)
    assert g["pairs"] == len(x) * (len(x) - 1)
    assert_parity(compare(g, r, ALL))
