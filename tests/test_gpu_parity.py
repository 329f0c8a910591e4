"""GPU parity: the CUDA path through the C ABI against the CPU oracle on identical seeded inputs."""
import numpy as np
import pytest

import coulombgalore_b200 as cg
from oracle import oracle as O

from parity import TOL, TOL_FP32, compare, run_case, splined_pair, zoo

pytestmark = pytest.mark.gpu

ALL = cg.ENERGY | cg.FORCE | cg.FIELD | cg.POTENTIAL


def lattice(n, a, seed, kind):
    return O.generate(n, a, seed, kind)


def assert_parity(errs):
    bad = {k: v for k, v in errs.items() if not (v <= 1.0)}
    assert not bad, "parity beyond %.0e (normalised errors, must be <= 1): %r" % (TOL, errs)


def test_config1_qpotential_all_pairs_open(oracle_kind):
    """BASELINE config 1: qPotential(14, 3), N = 16^3, a = 1.75, open box, ion-ion energy + force."""
    x, y, z, q, mu = lattice(16, 1.75, 1001, oracle_kind)
    mk, p = zoo()["qpotential3"]
    g, r = run_case(mk(), O.Scheme(p, oracle_kind), x, y, z, q, None, [28.0] * 3, False, cg.MODE_ION, cg.ENERGY | cg.FORCE)
    assert r["pairs"] == 4631758
    assert_parity(compare(g, r, cg.ENERGY | cg.FORCE))


@pytest.mark.parametrize("name", ["wolf", "fennell"])
def test_config2_small_cell_list_pbc(name, oracle_kind):
    """BASELINE config 2 at a size the oracle finishes in seconds: Wolf/Fennell(12, 0.25), a = 3, periodic, forces."""
    x, y, z, q, mu = lattice(24, 3.0, 1002, oracle_kind)
    mk, p = zoo()[name]
    g, r = run_case(mk(), O.Scheme(p, oracle_kind), x, y, z, q, None, [72.0] * 3, True, cg.MODE_ION, cg.FORCE)
    assert_parity(compare(g, r, cg.FORCE))


def test_config3_small_reactionfield_dipoles(oracle_kind):
    x, y, z, q, mu = lattice(20, 3.0, 1003, oracle_kind)
    mk, p = zoo()["reactionfield"]
    w = cg.ENERGY | cg.FIELD | cg.FORCE
    g, r = run_case(mk(), O.Scheme(p, oracle_kind), x, y, z, None, mu, [60.0] * 3, True, cg.MODE_DIPOLE, w)
    assert_parity(compare(g, r, w))


def test_config4_small_ewald_screened_multipole(oracle_kind):
    x, y, z, q, mu = lattice(20, 3.0, 1004, oracle_kind)
    mk, p = zoo()["ewald_screened"]
    w = cg.ENERGY | cg.FORCE
    g, r = run_case(mk(), O.Scheme(p, oracle_kind), x, y, z, q, mu, [60.0] * 3, True, cg.MODE_MULTIPOLE, w)
    assert_parity(compare(g, r, w))


def test_config5_small_splined_poisson_multipole(oracle_kind):
    x, y, z, q, mu = lattice(20, 3.0, 1005, oracle_kind)
    pot, osch = splined_pair("poisson43", oracle_kind)
    w = cg.ENERGY | cg.FORCE
    g, r = run_case(pot, osch, x, y, z, q, mu, [60.0] * 3, True, cg.MODE_MULTIPOLE, w)
    assert_parity(compare(g, r, w))


@pytest.mark.parametrize("name", sorted(zoo().keys()))
@pytest.mark.parametrize("mode", [cg.MODE_ION, cg.MODE_DIPOLE, cg.MODE_MULTIPOLE])
def test_every_scheme_every_mode(name, mode, oracle_kind):
    """Every scheme x {ion, dipole, multipole} x {energy, force, field, potential} at n = 16 (SURVEY.md 8d extras)."""
    mk, p = zoo()[name]
    pot = mk()
    pbc = pot.cutoff < 1e10  # Plain has no cut-off: open boundaries only
    x, y, z, q, mu = lattice(16 if pbc else 10, 3.0, 77 + mode, oracle_kind)
    L = 48.0
    g, r = run_case(pot, O.Scheme(p, oracle_kind), x, y, z, q, mu, [L] * 3, pbc, mode, ALL)
    assert_parity(compare(g, r, ALL))


@pytest.mark.parametrize("name", sorted(zoo().keys()))
@pytest.mark.parametrize("what", [cg.ENERGY | cg.FORCE, ALL])
def test_every_scheme_with_quadrupoles(name, what, oracle_kind):
    """quadrupole_potential, the quadrupole branches of multipole_field and of
    multipole_multipole_energy / _force (core.hpp:320, 477-481, 603-609, 767-773) with general 3x3 tensors."""
    x, y, z, q, mu = lattice(10, 3.0, 77, oracle_kind)
    quad = np.random.default_rng(5).normal(size=(x.size, 3, 3))
    mk, p = zoo()[name]
    pot = mk()
    pbc = pot.cutoff <= 30.0
    g, r = run_case(pot, O.Scheme(p, oracle_kind), x, y, z, q, mu, [30.0] * 3, pbc, cg.MODE_MULTIPOLE_QUAD, what, quad=quad)
    assert_parity(compare(g, r, what))


def test_quadrupole_mode_properties_and_golden(oracle_kind):
    """Zero tensors reproduce CG_MODE_MULTIPOLE, only the symmetric part and the trace matter, quadrupoles are cleared
    by set_system, borrowers see the lender's tensors; and the committed reference run."""
    import os
    from golden.make_golden import quad_inputs
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_vectors.npz"))
    lat = gold["lattice/n6_a3_seed31337"]
    x, y, z, q, mu = lat[0].copy(), lat[1].copy(), lat[2].copy(), lat[3].copy(), (lat[4].copy(), lat[5].copy(), lat[6].copy())
    quad = quad_inputs(x.size)
    for key in [k for k in gold.files if k.startswith("batch_quad/") and k.endswith("/scalars")]:
        case = key.split("/")[1]
        name, pbc = case.rsplit("_pbc", 1)[0], case.endswith("pbc1")
        ev = cg.BatchedEvaluator(zoo()[name][0]())
        ev.set_system(x, y, z, q, mu, [18.0] * 3, pbc)
        ev.set_quadrupoles(quad)
        g = ev.eval(cg.MODE_MULTIPOLE_QUAD, ALL)
        sc = gold[key]
        ref = dict(energy=sc[0], energy_abs=sc[1], pairs=int(sc[2]))
        for f in ("force", "field", "potential", "force_scale", "field_scale", "potential_scale"):
            ref[f] = gold["batch_quad/%s/%s" % (case, f)]
        assert_parity(compare(g, ref, ALL))
        ev.close()
    ev = cg.BatchedEvaluator(cg.Wolf(12.0, 0.25))
    ev.set_system(x, y, z, q, mu, [18.0] * 3, True)
    base = ev.eval(cg.MODE_MULTIPOLE, ALL)
    with pytest.raises(cg.CgError):
        ev.eval(cg.MODE_MULTIPOLE_QUAD, ALL)  # no tensors given
    ev.set_quadrupoles(np.zeros((x.size, 9)))
    zero = ev.eval(cg.MODE_MULTIPOLE_QUAD, ALL)
    assert zero["pairs"] == base["pairs"] and abs(zero["energy"] - base["energy"]) <= 1e-13 * abs(base["energy"])
    assert np.max(np.abs(zero["force"] - base["force"])) <= 1e-13 * np.max(np.abs(base["force"]))
    ev.set_quadrupoles(quad)
    full = ev.eval(cg.MODE_MULTIPOLE_QUAD, ALL)
    bor = cg.BatchedEvaluator(cg.Fennell(12.0, 0.25))
    bor.share_system(ev)
    own = cg.BatchedEvaluator(cg.Fennell(12.0, 0.25))
    own.set_system(x, y, z, q, mu, [18.0] * 3, True)
    own.set_quadrupoles(quad)
    assert np.array_equal(bor.eval(cg.MODE_MULTIPOLE_QUAD, ALL)["force"], own.eval(cg.MODE_MULTIPOLE_QUAD, ALL)["force"])
    ev.set_quadrupoles(0.5 * (quad + quad.transpose(0, 2, 1)))
    sym = ev.eval(cg.MODE_MULTIPOLE_QUAD, ALL)
    assert np.max(np.abs(sym["force"] - full["force"])) <= 1e-12 * np.max(np.abs(full["force"]))
    ev.set_system(x, y, z, q, mu, [18.0] * 3, True)
    with pytest.raises(cg.CgError):
        ev.eval(cg.MODE_MULTIPOLE_QUAD, ALL)  # set_system cleared them
    for e in (ev, bor, own):
        e.close()


@pytest.mark.parametrize("name", ["wolf", "ewald_screened", "qpotential3", "plain_yukawa", "reactionfield"])
def test_splined_schemes(name, oracle_kind):
    pot, osch = splined_pair(name, oracle_kind)
    pbc = pot.cutoff < 1e10
    x, y, z, q, mu = lattice(16 if pbc else 10, 3.0, 99, oracle_kind)
    g, r = run_case(pot, osch, x, y, z, q, mu, [48.0] * 3, pbc, cg.MODE_MULTIPOLE, ALL)
    assert_parity(compare(g, r, ALL))


@pytest.mark.parametrize("name", sorted(zoo().keys()))
def test_fp32_variant(name, oracle_kind):
    """CG_FP32: fp32 pair arithmetic (fp64 positions / distance vector / gate / accumulators) against the fp64 oracle,
    same norms at 1e-5; the in-cutoff pair count stays exact because the gate is evaluated in fp64."""
    mk, p = zoo()[name]
    pot = mk()
    pbc = pot.cutoff < 1e10
    x, y, z, q, mu = lattice(16 if pbc else 10, 3.0, 4711, oracle_kind)
    for mode in (cg.MODE_ION, cg.MODE_MULTIPOLE):
        g, r = run_case(pot, O.Scheme(p, oracle_kind), x, y, z, q, mu, [48.0] * 3, pbc, mode, ALL, precision="fp32")
        errs = compare(g, r, ALL, tol=TOL_FP32)
        bad = {k: v for k, v in errs.items() if not (v <= 1.0)}
        assert not bad, errs
        assert max(errs.values()) > 1e-6  # and it really is single precision


@pytest.mark.parametrize("seed", range(64))
def test_random_small_systems(seed, oracle_kind):
    """Randomised configurations: odd particle counts, non-cubic boxes down to one cut-off per side (every particle then
    meets images of itself and of its neighbours), particles on the box faces, clustered open systems, random scheme, mode
    and output mask."""
    rng = np.random.default_rng(1000 + seed)
    names = sorted(zoo().keys())
    name = names[int(rng.integers(len(names)))]
    mk, p = zoo()[name]
    pot = mk()
    rc = pot.cutoff
    pbc = bool(rng.random() < 0.6) and rc < 1e10
    n = int(rng.integers(1, 400))
    if pbc:
        box = [float(rc * rng.uniform(1.0, 3.5)) for _ in range(3)]
        pos = [rng.uniform(0.0, b, n) for b in box]
        for k in range(3):  # some particles exactly on the lower face and one ulp below the upper one
            if n > 4:
                pos[k][0] = 0.0
                pos[k][1] = np.nextafter(box[k], 0.0)
    else:
        box = [100.0, 100.0, 100.0]
        rcl = min(rc, 14.0)
        centres = rng.uniform(-50.0, 50.0, (3, 3))
        pick = rng.integers(0, 3, n)
        pos = [centres[pick, k] + rng.normal(scale=0.6 * rcl, size=n) for k in range(3)]
    # keep pairs apart (the reference diverges at r -> 0): reject near-duplicates
    P = np.stack(pos, axis=1)
    keep = np.ones(n, dtype=bool)
    for i in range(n):
        if keep[i]:
            d = P[i + 1:] - P[i]
            if pbc:
                d -= np.round(d / np.array(box)) * np.array(box)
            keep[i + 1:] &= np.einsum("ij,ij->i", d, d) > 0.25
    x, y, z = (np.ascontiguousarray(a[keep]) for a in pos)
    n = x.size
    q = rng.choice([-1.0, 1.0, 0.5, 2.0], n)
    mu = tuple(rng.normal(size=n) for _ in range(3))
    mode = int(rng.integers(0, 3))
    what = int(rng.integers(1, 16))
    g, r = run_case(pot, O.Scheme(p, oracle_kind), x, y, z, q, mu, box, pbc, mode, what)
    assert_parity(compare(g, r, what))


@pytest.mark.parametrize("seed", range(8))
def test_random_shard_counts_tile_the_system(seed, oracle_kind):
    """Any number of shards (also more shards than cell rows have work) tiles the evaluation: pair counts and energies add up,
    every particle's outputs come from exactly one shard."""
    rng = np.random.default_rng(500 + seed)
    name = ["wolf", "reactionfield", "poisson43", "qpotential3"][seed % 4]
    mk, p = zoo()[name]
    rc = mk().cutoff
    n = int(rng.integers(50, 3000))
    box = [float(rc * rng.uniform(1.5, 4.0)) for _ in range(3)]
    x, y, z = (rng.uniform(0.0, b, n) for b in box)
    q = rng.choice([-1.0, 1.0], n)
    mu = tuple(rng.normal(size=n) for _ in range(3))
    mode = [cg.MODE_ION, cg.MODE_DIPOLE, cg.MODE_MULTIPOLE][seed % 3]
    w = cg.ENERGY | cg.FORCE
    full, ref = run_case(mk(), O.Scheme(p, oracle_kind), x, y, z, q, mu, box, True, mode, w)
    R = int(rng.choice([2, 3, 5, 7, 16]))
    parts = [run_case(mk(), O.Scheme(p, oracle_kind), x, y, z, q, mu, box, True, mode, w, shard=(r, R))[0] for r in range(R)]
    assert sum(pt["pairs"] for pt in parts) == full["pairs"] == ref["pairs"]
    assert abs(sum(pt["energy"] for pt in parts) - ref["energy"]) <= TOL * max(ref["energy_abs"], 1e-300)
    fsum = sum(pt["force"] for pt in parts)
    assert np.max(np.linalg.norm(fsum - ref["force"], axis=1) / np.maximum(ref["force_scale"], 1e-300)) <= TOL
    owners = sum((np.abs(pt["force"]).sum(axis=1) > 0).astype(int) for pt in parts)
    assert owners.max() <= 1
