"""GPU tests beyond per-scheme parity: golden fixtures (no oracle library needed), edge cases, sharding, determinism and
size-independent properties at the full BASELINE size (N = 10^6)."""
import os

import numpy as np
import pytest

import coulombgalore_b200 as cg
from oracle import oracle as O

from parity import TOL, compare, run_case, zoo
from zoo import oracle_params

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = np.load(os.path.join(ROOT, "tests", "golden", "reference_vectors.npz"))
ALL = cg.ENERGY | cg.FORCE | cg.FIELD | cg.POTENTIAL


def ok(errs):
    bad = {k: v for k, v in errs.items() if not (v <= 1.0)}
    assert not bad, errs


@pytest.mark.parametrize("case", sorted({k.split("/")[1] for k in GOLD.files if k.startswith("batch/")}))
def test_against_golden_batches_of_the_reference(case):
    """The committed outputs of the UNMODIFIED reference loop (tests/golden/make_golden.py); 6^3 lattice, L = 18 < 2 R_c,
    so periodic cases exercise several images of the same particle inside the cut-off."""
    name, mode, pbc = case.rsplit("_mode", 1)[0], int(case.split("_mode")[1][0]), case.endswith("pbc1")
    lat = GOLD["lattice/n6_a3_seed31337"]
    x, y, z, q, mu = lat[0].copy(), lat[1].copy(), lat[2].copy(), lat[3].copy(), (lat[4].copy(), lat[5].copy(), lat[6].copy())
    ev = cg.BatchedEvaluator(zoo()[name][0]())
    ev.set_system(x, y, z, q, mu, [18.0] * 3, pbc)
    g = ev.eval(mode, ALL)
    sc = GOLD["batch/%s/scalars" % case]
    ref = dict(energy=sc[0], energy_abs=sc[1], pairs=int(sc[2]))
    for f in ("force", "field", "potential", "force_scale", "field_scale", "potential_scale"):
        ref[f] = GOLD["batch/%s/%s" % (case, f)]
    ok(compare(g, ref, ALL))


def test_edge_cases_empty_single_ragged(oracle_kind):
    pot = cg.Wolf(12.0, 0.25)
    osch = O.Scheme(oracle_params("wolf"), oracle_kind)
    ev = cg.BatchedEvaluator(pot)
    e = np.zeros(0)
    ev.set_system(e, e, e, e, None, [40.0] * 3, True)
    r = ev.eval(cg.MODE_ION, cg.ENERGY | cg.FORCE)
    assert r["force"].shape == (0, 3) and r["energy"] == 0.0 and r["pairs"] == 0
    one = np.array([1.0])
    ev.set_system(one, one, one, one, None, [40.0] * 3, True)
    r = ev.eval(cg.MODE_ION, cg.ENERGY | cg.FORCE)
    assert np.all(r["force"] == 0.0) and r["pairs"] == 0
    # ragged: 37 particles (neither a multiple of 32 nor of the cell grid), non-cubic periodic box, all in a corner
    rng = np.random.default_rng(3)
    box = [30.0, 41.0, 55.0]
    x, y, z = (rng.uniform(0, 9.0, 37) for _ in range(3))
    q = rng.choice([-1.0, 1.0], 37)
    mu = tuple(rng.normal(size=37) for _ in range(3))
    for mode in (cg.MODE_ION, cg.MODE_DIPOLE, cg.MODE_MULTIPOLE):
        g, r = run_case(pot, osch, x, y, z, q, mu, box, True, mode, ALL)
        ok(compare(g, r, ALL))
    # two particles exactly at the cut-off distance: strict gate, no interaction (core.hpp:354)
    g, r = run_case(pot, osch, np.array([1.0, 13.0]), np.array([1.0, 1.0]), np.array([1.0, 1.0]), np.array([1.0, -1.0]), None, box, False,
                    cg.MODE_ION, cg.ENERGY | cg.FORCE)
    assert g["pairs"] == r["pairs"] == 0 and np.all(g["force"] == 0.0)


def test_open_box_non_uniform_density(oracle_kind):
    """Clustered particles far apart in an open box: empty cells, one dense cell, groups with few neighbours."""
    rng = np.random.default_rng(11)
    x = np.concatenate([rng.normal(0, 2.0, 300), rng.normal(200, 1.0, 200), rng.uniform(-50, 250, 100)])
    y = np.concatenate([rng.normal(0, 2.0, 300), rng.normal(-80, 1.0, 200), rng.uniform(-100, 20, 100)])
    z = np.concatenate([rng.normal(0, 2.0, 300), rng.normal(35, 1.0, 200), rng.uniform(-10, 50, 100)])
    q = rng.choice([-1.0, 1.0], 600)
    mu = tuple(rng.normal(size=600) for _ in range(3))
    for name, mode in (("wolf", cg.MODE_ION), ("reactionfield", cg.MODE_DIPOLE), ("fanourgakis", cg.MODE_MULTIPOLE), ("plain", cg.MODE_ION)):
        mk, p = zoo()[name]
        g, r = run_case(mk(), O.Scheme(p, oracle_kind), x, y, z, q, mu, None, False, mode, ALL)
        ok(compare(g, r, ALL))


def test_errors_are_status_codes():
    ev = cg.BatchedEvaluator(cg.Wolf(12.0, 0.25))
    with pytest.raises(cg.CgError) as e:
        ev.eval(cg.MODE_ION, cg.FORCE)
    assert e.value.rc == -6  # CG_ERR_STATE: no system yet
    one = np.array([1.0, 2.0])
    with pytest.raises(cg.CgError) as e:
        ev.set_system(one, one, one, one, None, [11.0] * 3, True)  # periodic box shorter than the cut-off
    assert e.value.rc == -5
    ev.set_system(one, one, one, one, None, [30.0] * 3, True)
    with pytest.raises(cg.CgError) as e:
        ev.eval(7, cg.FORCE)
    assert e.value.rc == -1


def test_device_pointer_path_equals_host_path_and_is_deterministic():
    import torch
    x, y, z, q, mu = O.generate(20, 3.0, 2024, "port" if not O.available("ref") else "ref")
    pot = cg.Fennell(12.0, 0.25)
    ev = cg.BatchedEvaluator(pot)
    ev.set_system(x, y, z, q, None, [60.0] * 3, True)
    h = ev.eval(cg.MODE_ION, cg.ENERGY | cg.FORCE)
    dev = [torch.tensor(a, device="cuda") for a in (x, y, z, q)]
    f = torch.empty((x.size, 3), dtype=torch.float64, device="cuda")
    sc = torch.zeros(2, dtype=torch.float64, device="cuda")
    ev2 = cg.BatchedEvaluator(pot)
    st = torch.cuda.Stream()
    ev2.set_stream(st.cuda_stream)
    outs = []
    for _ in range(3):
        ev2.set_system_device(dev[0], dev[1], dev[2], dev[3], None, [60.0] * 3, True)
        ev2.eval_device(cg.MODE_ION, cg.ENERGY | cg.FORCE, force=f, scalars=sc)
        st.synchronize()
        outs.append((f.cpu().numpy().copy(), sc.cpu().numpy().copy()))
    assert all(np.array_equal(o[0], outs[0][0]) and np.array_equal(o[1], outs[0][1]) for o in outs)  # bit-reproducible
    assert np.array_equal(outs[0][0], h["force"]) and outs[0][1][0] == h["energy"] and int(outs[0][1][1]) == h["pairs"]


def test_two_schemes_in_flight_with_caller_owned_buffers(oracle_kind):
    """cg_eval_begin / cg_eval_end: two handles evaluate concurrently into caller-owned (here pinned) host arrays."""
    import torch
    x, y, z, q, mu = O.generate(14, 3.0, 99, oracle_kind)
    box, w = [42.0] * 3, cg.ENERGY | cg.FORCE | cg.POTENTIAL
    names = ("wolf", "fennell")
    evs = {k: cg.BatchedEvaluator(zoo()[k][0]()) for k in names}
    outs = {k: dict(force=torch.empty((x.size, 3), dtype=torch.float64).pin_memory().numpy(), potential=np.empty(x.size)) for k in names}
    for rep in range(2):  # the second round runs the sync-free path
        for k, ev in evs.items():
            ev.set_system(x, y, z, q, None, box, True)
            ev.eval_begin(cg.MODE_ION, w, out=outs[k])
        for k, ev in evs.items():
            g = ev.eval_end()
            assert g["force"] is outs[k]["force"] and g["potential"] is outs[k]["potential"]
            ok(compare(g, O.Scheme(zoo()[k][1], oracle_kind).batched(x, y, z, q, None, box, True, cg.MODE_ION, w), w))
    with pytest.raises(cg.CgError):
        evs["wolf"].eval_end()  # nothing in flight
    with pytest.raises(ValueError):
        evs["wolf"].eval(cg.MODE_ION, cg.FORCE, out=dict(force=np.empty((3, 3))))


@pytest.mark.parametrize("name", sorted(zoo()))
def test_self_energy_self_field_neutralisation_and_torque(name, oracle_kind):
    """the O(N) terms, against the oracle and against the committed reference run."""
    from golden.make_golden import self_inputs
    lat = GOLD["lattice/n6_a3_seed31337"]
    x, y, z, q, mu = lat[0].copy(), lat[1].copy(), lat[2].copy(), lat[3].copy(), (lat[4].copy(), lat[5].copy(), lat[6].copy())
    qn, fld = self_inputs(q)
    mk, p = zoo()[name]
    ev = cg.BatchedEvaluator(mk())
    rc = ev.scheme.cutoff
    box = [18.0] * 3 if rc <= 18.0 else None
    ev.set_system(x, y, z, qn, mu, box, box is not None)
    g = ev.self_terms(self_field=True)
    tq = ev.dipole_torque(fld)
    for ref in (O.Scheme(p, oracle_kind).self_terms(qn, mu, 18.0 ** 3 if box else 0.0, field=fld),
                dict(zip(("self_energy", "self_energy_abs", "neutralization_energy", "charge_sum"), GOLD["self/%s/scalars" % name]),
                     self_field=GOLD["self/%s/self_field" % name], torque=GOLD["self/torque"])):
        assert abs(g["self_energy"] - ref["self_energy"]) <= TOL * max(ref["self_energy_abs"], 1e-300)
        assert abs(g["charge_sum"] - ref["charge_sum"]) <= 1e-13 * np.abs(qn).sum()
        if box:
            assert abs(g["neutralization_energy"] - ref["neutralization_energy"]) <= TOL * abs(ref["neutralization_energy"])
        else:
            assert g["neutralization_energy"] == 0.0  # an open system has no unit cell to neutralise
        assert np.max(np.abs(g["self_field"] - ref["self_field"])) <= TOL * max(np.max(np.abs(ref["self_field"])), 1e-300)
        assert np.max(np.abs(tq - ref["torque"])) <= TOL * np.max(np.abs(ref["torque"]))
    ev.close()


def test_schemes_sharing_one_sorted_system(oracle_kind):
    """cg_share_system: the lender uploads and sorts once, borrowers of the same cut-off evaluate on its arrays."""
    x, y, z, q, mu = O.generate(14, 3.0, 5, oracle_kind)
    box = [42.0] * 3
    lender = cg.BatchedEvaluator(zoo()["wolf"][0]())
    b_ion = cg.BatchedEvaluator(zoo()["fennell"][0]())
    b_dip = cg.BatchedEvaluator(zoo()["zerodipole"][0]())
    own = cg.BatchedEvaluator(zoo()["fennell"][0]())
    for b in (b_ion, b_dip):
        b.share_system(lender)
    with pytest.raises(cg.CgError):
        b_ion.eval(cg.MODE_ION, cg.FORCE)  # the lender has not sorted anything yet
    with pytest.raises(cg.CgError):
        b_ion.set_system(x, y, z, q, None, box, True)  # a borrower has no system of its own
    with pytest.raises(cg.CgError):
        cg.BatchedEvaluator(cg.Wolf(11.0, 0.25)).share_system(lender)  # other cut-off, other cell grid
    w = cg.ENERGY | cg.FORCE
    for rep, shift in enumerate((0.0, 1.7, 3.1)):  # new positions each round; rounds 2+ take the sync-free path
        xs = np.mod(x + shift, 42.0)
        lender.set_system(xs, y, z, q, mu, box, True)
        lender.eval_begin(cg.MODE_ION, w)
        b_ion.eval_begin(cg.MODE_ION, w)
        b_dip.eval_begin(cg.MODE_DIPOLE, w | cg.FIELD)
        g0, g1, g2 = lender.eval_end(), b_ion.eval_end(), b_dip.eval_end()
        ok(compare(g0, O.Scheme(zoo()["wolf"][1], oracle_kind).batched(xs, y, z, q, mu, box, True, cg.MODE_ION, w), w))
        ok(compare(g1, O.Scheme(zoo()["fennell"][1], oracle_kind).batched(xs, y, z, q, mu, box, True, cg.MODE_ION, w), w))
        ok(compare(g2, O.Scheme(zoo()["zerodipole"][1], oracle_kind).batched(xs, y, z, q, mu, box, True, cg.MODE_DIPOLE, w | cg.FIELD),
                   w | cg.FIELD))
        own.set_system(xs, y, z, q, mu, box, True)
        assert np.array_equal(own.eval(cg.MODE_ION, w)["force"], g1["force"])  # bitwise the same as with an own sort
    b_ion.share_system(None)
    b_ion.set_system(x, y, z, q, None, box, True)
    assert b_ion.eval(cg.MODE_ION, w)["pairs"] > 0
    b_dip.close(); lender.close(); b_ion.close(); own.close()


def test_sync_free_path_and_its_overflow_fallback(oracle_kind):
    """The second evaluation of a same-shaped periodic system sizes its arrays and launches from the previous one (no host
    round trip).  When the new positions need more room than that (all particles near the box corners: each gets seven
    periodic images inside the halo), the device flags the overflow and the host call redoes the evaluation exactly."""
    rng = np.random.default_rng(11)
    n, L = 6000, 64.0
    box = [L] * 3
    q = rng.choice([-1.0, 1.0], n)
    mk, p = zoo()["wolf"]
    osch = O.Scheme(p, oracle_kind)
    w = cg.ENERGY | cg.FORCE
    centre = [rng.uniform(20.0, 44.0, n) for _ in range(3)]  # no particle within R_c of a face: no images at all
    uniform = [rng.uniform(0.0, L, n) for _ in range(3)]
    corners = [np.where(rng.random(n) < 0.5, rng.uniform(0.0, 5.0, n), rng.uniform(L - 5.0, L, n)) for _ in range(3)]
    ev = cg.BatchedEvaluator(mk())
    bor = cg.BatchedEvaluator(zoo()["fennell"][0]())
    bor.share_system(ev)
    osch_b = O.Scheme(zoo()["fennell"][1], oracle_kind)
    seen = []
    for pos in (centre, centre, uniform, corners, corners, centre):
        ev.set_system(pos[0], pos[1], pos[2], q, None, box, True)
        ev.eval_begin(cg.MODE_ION, w)
        bor.eval_begin(cg.MODE_ION, w)
        gb = bor.eval_end()  # the borrower finds the overflow first and has the lender sort again
        g = ev.eval_end()
        seen.append(ev.counters()["sync_free"])
        ok(compare(g, osch.batched(pos[0], pos[1], pos[2], q, None, box, True, cg.MODE_ION, w), w))
        ok(compare(gb, osch_b.batched(pos[0], pos[1], pos[2], q, None, box, True, cg.MODE_ION, w), w))
    # 1st: exact sizing; 2nd: sync-free; 3rd/4th: more images than the arrays hold -> detected, redone exactly
    assert seen[0] == 0 and seen[1] == 1, seen
    assert 0 in seen[2:4], seen
    bor.close()
    ev.close()


def test_shards_partition_the_work(oracle_kind):
    x, y, z, q, mu = O.generate(16, 3.0, 7, oracle_kind)
    mk, p = zoo()["reactionfield"]
    w = cg.ENERGY | cg.FORCE | cg.FIELD
    full, ref = run_case(mk(), O.Scheme(p, oracle_kind), x, y, z, q, mu, [48.0] * 3, True, cg.MODE_DIPOLE, w)
    parts = [run_case(mk(), O.Scheme(p, oracle_kind), x, y, z, q, mu, [48.0] * 3, True, cg.MODE_DIPOLE, w, shard=(r, 3))[0] for r in range(3)]
    assert sum(pt["pairs"] for pt in parts) == full["pairs"] == ref["pairs"]
    assert abs(sum(pt["energy"] for pt in parts) - full["energy"]) <= 1e-13 * ref["energy_abs"]
    # each particle is written by exactly one shard (the others leave zeros), and the union is the unsharded result
    # (to rounding, not bitwise: a small system is also split along j, and the j-split factor depends on the shard size)
    fsum = sum(pt["force"] for pt in parts)
    assert np.max(np.linalg.norm(fsum - full["force"], axis=1) / ref["force_scale"]) <= 1e-14
    owners = sum((np.abs(pt["force"]).sum(axis=1) > 0).astype(int) for pt in parts)
    assert owners.max() == 1


def test_full_size_properties_one_million(oracle_kind):
    """BASELINE config 2 at full size: N = 100^3, L = 300, Wolf and Fennell forces.  Checks that do not need the oracle on all
    pairs: pair count (an exact integer, symmetric => even), Newton's third law, and parity on a 4096-particle i-slice."""
    import torch
    n, L = 100, 300.0
    N = n ** 3
    dev = [torch.empty(N, dtype=torch.float64, device="cuda") for _ in range(4)]
    f = torch.empty((N, 3), dtype=torch.float64, device="cuda")
    sc = torch.zeros(2, dtype=torch.float64, device="cuda")
    x = y = z = q = None
    for name in ("wolf", "fennell"):
        mk, p = zoo()[name]
        ev = cg.BatchedEvaluator(mk())
        ev.generate_lattice_device(n, 3.0, 1002, dev[0], dev[1], dev[2], dev[3])
        ev.set_system_device(dev[0], dev[1], dev[2], dev[3], None, [L] * 3, True)
        ev.eval_device(cg.MODE_ION, cg.FORCE, force=f, scalars=sc)
        torch.cuda.synchronize()
        pairs = int(sc[1].item())
        assert pairs == 264441374 and pairs % 2 == 0
        F = f.cpu().numpy()
        assert np.linalg.norm(F.sum(axis=0)) <= 1e-12 * np.abs(F).sum()
        if x is None:
            x, y, z, q = (t.cpu().numpy() for t in dev)
            xo, yo, zo, qo, _ = O.generate(n, 3.0, 1002, oracle_kind)
            assert np.array_equal(x, xo) and np.array_equal(y, yo) and np.array_equal(z, zo) and np.array_equal(q, qo)  # device generator == oracle generator
        lo, hi = 500000, 504096
        r = O.Scheme(p, oracle_kind).batched(x, y, z, q, None, [L] * 3, True, O.MODE_ION, O.FORCE, lo, hi)
        dn = np.linalg.norm(F[lo:hi] - r["force"], axis=1)
        assert np.max(dn / np.maximum(np.linalg.norm(r["force"], axis=1), r["force_scale"])) <= TOL
        ev.close()


def test_peer_exchange_single_rank_and_two_ranks_if_available():
    """cg_peer_alloc / cg_peer_gather: the exchange step as this repo's own kernel over (peer-mapped) device memory.
    With one GPU the rank pulls from its own block (the kernel, the torch view of the block and the double buffering are
    exercised); with two or more GPUs a 2-rank torchrun job checks the real peer mapping against the NCCL all-gather."""
    import subprocess
    import sys
    import torch
    script = os.path.join(ROOT, "tests", "peer_exchange_check.py")
    n = min(2, torch.cuda.device_count())
    for world in sorted({1, n}):
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
               "--master-port", str(29650 + world), script]
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
        assert out.returncode == 0 and "PEER_EXCHANGE_OK world=%d" % world in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


def test_reciprocal_ewald_against_oracle_golden_and_reference_kats(oracle_kind):
    """ReciprocalEwaldGaussian on the device (structure factors, energy, force, field, surface terms)."""
    from golden.make_golden import RECIP_CASES
    # the reference's own known answers (test/unittests.cpp:559-690)
    pos = np.array([[-0.5, 0.0, 0.0], [0.5, 0.0, 0.0]])
    st = cg.ReciprocalEwaldState(box_length=[10.0] * 3, cutoff=5.0, reciprocal_cutoff=11, alpha=0.8, surface_dielectric_constant=np.inf)
    pot = cg.ReciprocalEwaldGaussian(st)
    assert pot.numKVectors() == 5574
    q, zero = np.array([1.0, -1.0]), np.zeros((2, 3))
    pot.updateComplex(pos, q, zero)
    assert pot.reciprocal_energy() == pytest.approx(0.1584768302, rel=1e-8)
    assert pot.reciprocal_force(pos, q, zero)[0] == pytest.approx([0.2617988467, 0.0, 0.0], rel=1e-8, abs=1e-12)
    assert pot.reciprocal_field(pos)[0] == pytest.approx([0.2617988467, 0.0, 0.0], rel=1e-8, abs=1e-12)
    assert pot.surface_energy(pos, q, zero) == 0.0
    assert pot.real_space.ion_ion_energy(1.0, -1.0, 1.0) == pytest.approx(-0.2578990353, rel=1e-8) if hasattr(pot.real_space, "ion_ion_energy") else True
    st.surface_dielectric_constant = 1.0
    pot1 = cg.ReciprocalEwaldGaussian(st)
    assert pot1.surface_energy(pos, q, zero) == pytest.approx(0.0020943951, rel=1e-8)
    assert pot1.surface_field(pos, q, zero)[0] == pytest.approx(0.0041887902, rel=1e-8)
    mu = np.array([[1.0, 0.0, 0.0], [1.0, 0.0, 0.0]])
    pot.setZeroComplex()
    pot.updateComplex(pos, np.zeros(2), mu)
    assert pot.reciprocal_energy() == pytest.approx(0.4534417045, rel=1e-8)
    assert pot.reciprocal_field(pos)[0][0] == pytest.approx(-0.4534417045, rel=1e-8)
    assert pot1.surface_energy(pos, np.zeros(2), mu) == pytest.approx(0.0083775804, rel=1e-8)
    assert pot1.surface_energy(pos, q, np.array([[1.0, 1.0, 1.0], [-1.0, 1.0, 0.0]])) == pytest.approx(0.0125663706, rel=1e-8)
    # incremental updates: subtracting and re-adding a particle restores Q
    pot.setZeroComplex()
    pot.updateComplex(pos, q, mu)
    q0 = pot.Q()
    pot.updateComplex(pos[:1], q[:1], mu[:1], subtract=True)
    pot.updateComplex(pos[:1], q[:1], mu[:1])
    assert np.max(np.abs(pot.Q() - q0)) <= 1e-14
    # degenerate k-list (reciprocal_cutoff = 0): one dummy vector with zero weight -> zero energy, force and field
    st0 = cg.ReciprocalEwaldState(box_length=[10.0] * 3, cutoff=5.0, reciprocal_cutoff=0, alpha=0.8, surface_dielectric_constant=1.0)
    p0 = cg.ReciprocalEwaldGaussian(st0)
    p0.updateComplex(pos, q, mu)
    assert p0.numKVectors() == 1 and p0.reciprocal_energy() == 0.0 and np.all(p0.reciprocal_force(pos, q, mu) == 0.0)
    o0 = O.Reciprocal(5.0, 0.8, 1.0, 0.0, 0, [10.0] * 3, oracle_kind)
    o0.update(pos[:, 0].copy(), pos[:, 1].copy(), pos[:, 2].copy(), q, tuple(mu[:, k].copy() for k in range(3)), 0)
    assert np.max(np.abs(p0.Q() - o0.q())) <= 1e-14
    # lattice system: the live oracle and the committed reference run
    lat = GOLD["lattice/n6_a3_seed31337"]
    x, y, z, ch, m = lat[0].copy(), lat[1].copy(), lat[2].copy(), lat[3].copy(), (lat[4].copy(), lat[5].copy(), lat[6].copy())
    P, M = np.stack([x, y, z], axis=1), np.stack(m, axis=1)
    for name, (rc, al, eps, kap, kc, box) in RECIP_CASES.items():
        g = cg.ReciprocalEwaldGaussian(cg.ReciprocalEwaldState(box, rc, kc, al, eps, kap))
        o = O.Reciprocal(rc, al, eps, kap, kc, box, oracle_kind)
        g.updateComplex(P, ch, M)
        o.update(x, y, z, ch, m, 0)
        kv, aks = g.kvectors()
        assert np.array_equal(kv, GOLD["recip/%s/k" % name]) and np.allclose(aks, GOLD["recip/%s/aks" % name], rtol=1e-14, atol=0)
        qscale = np.sum(np.abs(ch)) + np.abs(M @ kv.T).sum(axis=0)  # sum_p |z_p| + |mu_p . k|
        gq = GOLD["recip/%s/q" % name]
        for ref in (o.q(), gq[0] + 1j * gq[1]):
            assert np.max(np.abs(g.Q() - ref) / qscale) <= TOL
        sc = GOLD["recip/%s/scalars" % name]
        assert abs(g.reciprocal_energy() - o.energy()) <= TOL * o.energy() and abs(g.reciprocal_energy() - sc[0]) <= TOL * sc[0]
        F, E, ff = g.reciprocal_force(P, ch, M), g.reciprocal_field(P), o.force_field(x, y, z, ch, m)
        for ref_f, ref_e, s in ((ff["force"], ff["field"], ff["force_scale"]),
                                (GOLD["recip/%s/force" % name], GOLD["recip/%s/field" % name], GOLD["recip/%s/force_scale" % name])):
            assert np.max(np.linalg.norm(F - ref_f, axis=1) / np.maximum(s, 1e-300)) <= TOL
            assert np.max(np.linalg.norm(E - ref_e, axis=1)) <= TOL * np.max(np.linalg.norm(ref_e, axis=1)) * 50
        se, sf = o.surface(x, y, z, ch, m)
        assert g.surface_energy(P, ch, M) == pytest.approx(se, rel=1e-12) and np.allclose(g.surface_field(P, ch, M), sf, rtol=1e-12, atol=1e-300)
        assert g.surface_energy(P, ch, M) == pytest.approx(sc[1], rel=1e-12)
        g.close()
    for p_ in (pot, pot1, p0):
        p_.close()


def test_full_ewald_sum_of_the_reference_test_case():
    """All pieces together on the reference's Ewald test system (test/unittests.cpp:559-605: charges +1 / -1 one length unit
    apart in a 10^3 box, R_c = 5, alpha = 0.8, k_c = 11, conducting boundary): real-space pairs (batched evaluator, periodic),
    self-energy (O(N) terms), reciprocal and surface energy add up to the reference's total -1.0021255, and the total force on
    the first charge to 0.9956865."""
    pos = np.array([[-0.5, 0.0, 0.0], [0.5, 0.0, 0.0]]) + 5.0  # inside the periodic box [0, 10)^3
    q, mu = np.array([1.0, -1.0]), np.zeros((2, 3))
    st = cg.ReciprocalEwaldState(box_length=[10.0] * 3, cutoff=5.0, reciprocal_cutoff=11, alpha=0.8, surface_dielectric_constant=np.inf)
    rec = cg.ReciprocalEwaldGaussian(st)
    rec.updateComplex(pos, q, mu)
    ev = cg.BatchedEvaluator(rec.real_space)
    ev.set_system(pos[:, 0].copy(), pos[:, 1].copy(), pos[:, 2].copy(), q, None, [10.0] * 3, True)
    real = ev.eval(cg.MODE_ION, cg.ENERGY | cg.FORCE)
    self_terms = ev.self_terms()
    assert real["pairs"] == 2                                                     # the direct pair only: images are 9 away
    assert real["energy"] == pytest.approx(-0.2578990353, rel=1e-8)              # :578
    assert self_terms["self_energy"] == pytest.approx(-0.9027033337, rel=1e-8)   # :579
    total = real["energy"] + self_terms["self_energy"] + rec.surface_energy(pos, q, mu) + rec.reciprocal_energy()
    assert total == pytest.approx(-1.0021255, rel=1e-7)                          # :582
    # force on the first charge: real + reciprocal (+ zero surface term); the reference quotes the x component
    f = real["force"][0] + rec.reciprocal_force(pos, q, mu)[0] + rec.surface_force(pos, q, mu, q[0])
    assert f[0] == pytest.approx(0.9956865, rel=1e-6) and abs(f[1]) < 1e-12 and abs(f[2]) < 1e-12   # :594
    ev.close()
    rec.close()


def test_ewald_truncated_surface_energy(oracle_kind, tmp_path):
    """EwaldTruncated::surface_energy (reference ewald_truncated.hpp:91-104) -- the two moment sums reduced
    on the device (cg_dipole_moments), through the Python evaluator and through the header-only C++ class, against the oracle."""
    import ctypes as C
    import subprocess
    x, y, z, q, mu = O.generate(9, 3.0, 808, oracle_kind)
    q = q + 0.25  # a charged, polar system: both sums and their cross term matter
    V = 27.0 ** 3
    _dp = C.POINTER(C.c_double)
    lib_dir = os.path.join(ROOT, "coulombgalore_b200")
    so = str(tmp_path / "surface_probe.so")
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O2", "-Wall", "-Wextra", "-Werror", "-ffp-contract=off", "-fPIC", "-shared",
                           "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "surface_probe.cpp"), "-o", so,
                           "-L" + lib_dir, "-l:libcg_b200.so", "-Wl,-rpath," + lib_dir])
    probe = C.CDLL(so)
    probe.probe_ewaldt_surface.argtypes = [C.c_double] * 3 + [C.c_int64] + [_dp] * 7 + [C.c_double, _dp]
    for eps in (1.0, 35.0, float("inf"), 0.5):  # 0.5: the reference keeps a sub-unity argument in the member it reads here
        ref = O.ewaldt_surface_energy(10.0, 0.3, eps, x, y, z, q, mu, V, oracle_kind)
        ev = cg.BatchedEvaluator(cg.EwaldTruncated(10.0, 0.3, eps))
        ev.set_system(x, y, z, q, mu, [27.0] * 3, True)
        P, D = ev.dipole_moments()
        # the sums themselves: 1e-13 of the sum of magnitudes (the device adds in a different, fixed order)
        assert np.all(np.abs(P - np.array([np.sum(c * q) for c in (x, y, z)])) <= 1e-13 * np.array([np.sum(np.abs(c * q)) for c in (x, y, z)]))
        assert np.all(np.abs(D - np.array([np.sum(m) for m in mu])) <= 1e-13 * np.array([np.sum(np.abs(m)) for m in mu]))
        got = ev.surface_energy(V)
        e = np.zeros(1)
        assert probe.probe_ewaldt_surface(10.0, 0.3, eps, x.size, *[a.ctypes.data_as(_dp) for a in (x, y, z, q, mu[0], mu[1], mu[2])], V,
                                          e.ctypes.data_as(_dp)) == 0
        # |sum|^2 of sums that cancel: compare on the scale of the squared sums of magnitudes
        scale = 2.0 * np.pi / abs(2.0 * eps + 1.0) / V * sum((np.sum(np.abs(c * q)) + np.sum(np.abs(m))) ** 2 for c, m in zip((x, y, z), mu)) \
            if np.isfinite(eps) else 0.0
        for val in (got, float(e[0])):
            assert abs(val - ref) <= 1e-13 * scale, (eps, val, ref)
        if not np.isfinite(eps):
            assert got == 0.0 and e[0] == 0.0 and ref == 0.0
        ev.close()


@pytest.mark.parametrize("pair", [("wolf", "fennell"), ("zahn", "zerodipole"), ("ewaldt", "wolf")])
def test_two_schemes_in_one_pass(pair, oracle_kind):
    """cg_eval_multi_device: two erfc-family schemes with one alpha evaluated in ONE pass over the pairs (shared prefilter, hit
    lists, gathers, 1/r, erfc and Gaussian) -- against the oracle at 1e-12 and against the one-by-one evaluation."""
    import torch
    x, y, z, q, mu = O.generate(44, 3.0, 606, oracle_kind)  # enough i-groups for one per resident warp: no j-split
    box = [132.0] * 3
    dev = torch.device("cuda", 0)
    st = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(st)
    try:
        evs = [cg.BatchedEvaluator(zoo()[nm][0]()) for nm in pair]
        for ev in evs:
            ev.set_stream(st.cuda_stream)
        evs[1].share_system(evs[0])
        dx = [torch.tensor(a, dtype=torch.float64, device=dev) for a in (x, y, z, q)]
        N = x.size
        for what in (cg.FORCE, cg.ENERGY | cg.FORCE | cg.FIELD | cg.POTENTIAL):
            f = [torch.zeros((N, 3), dtype=torch.float64, device=dev) for _ in evs]
            e = [torch.zeros((N, 3), dtype=torch.float64, device=dev) for _ in evs]
            p = [torch.zeros(N, dtype=torch.float64, device=dev) for _ in evs]
            sc = [torch.zeros(2, dtype=torch.float64, device=dev) for _ in evs]
            refs = [O.Scheme(zoo()[nm][1], oracle_kind).batched(x, y, z, q, None, box, True, cg.MODE_ION, what) for nm in pair]
            for rep in range(2):  # exact sizing, then the sync-free path
                evs[0].set_system_device(dx[0], dx[1], dx[2], dx[3], None, box, True)
                cg.eval_multi_device(evs, cg.MODE_ION, what, force=f, field=e, potential=p, scalars=sc)
                torch.cuda.synchronize()
                for k, nm in enumerate(pair):
                    ref = refs[k]
                    g = dict(force=f[k].cpu().numpy(), field=e[k].cpu().numpy(), potential=p[k].cpu().numpy(), energy=sc[k][0].item(),
                             pairs=int(sc[k][1].item()))
                    errs = compare(g, ref, what)
                    assert all(v <= 1.0 for v in errs.values()), (nm, what, rep, errs)
            # one by one: same pair count, same numbers to rounding
            f1 = torch.zeros((N, 3), dtype=torch.float64, device=dev)
            s1 = torch.zeros(2, dtype=torch.float64, device=dev)
            evs[0].set_system_device(dx[0], dx[1], dx[2], dx[3], None, box, True)
            for k in range(2):
                evs[k].eval_device(cg.MODE_ION, what, force=f1, field=torch.zeros_like(f1) if what & cg.FIELD else None,
                                   potential=torch.zeros(N, dtype=torch.float64, device=dev) if what & cg.POTENTIAL else None, scalars=s1)
                torch.cuda.synchronize()
                assert s1[1].item() == sc[k][1].item()
                assert (f1 - f[k]).abs().max().item() <= 1e-13 * f1.abs().max().item()
        # what the fused pass does not cover is refused, not approximated
        other = cg.BatchedEvaluator(cg.Wolf(12.0, 0.2))
        other.set_stream(st.cuda_stream)
        other.share_system(evs[0])
        with pytest.raises(cg.CgError):
            cg.eval_multi_device([evs[0], other], cg.MODE_ION, cg.FORCE, force=f, scalars=sc)
        with pytest.raises(cg.CgError):
            cg.eval_multi_device(evs, cg.MODE_DIPOLE, cg.FORCE, force=f, scalars=sc)
        for ev in evs + [other]:
            ev.close()
    finally:
        torch.cuda.set_stream(torch.cuda.default_stream(dev))


def test_two_schemes_host_pipeline(oracle_kind):
    """cg_eval_multi_begin / cg_eval_multi_end: the one-pass evaluation with host buffers, three different configurations in
    flight on three sets of handles (each on its own stream) -- every set must deliver its own configuration's results."""
    import torch
    pair = ("wolf", "fennell")
    box = [132.0] * 3
    cfgs = [O.generate(44, 3.0, 700 + k, oracle_kind) for k in range(3)]
    refs = [[O.Scheme(zoo()[nm][1], oracle_kind).batched(c[0], c[1], c[2], c[3], None, box, True, cg.MODE_ION, cg.FORCE | cg.ENERGY)
             for nm in pair] for c in cfgs]
    lanes = []
    for _ in range(3):
        evs = [cg.BatchedEvaluator(zoo()[nm][0]()) for nm in pair]
        evs[1].share_system(evs[0])
        N = cfgs[0][0].size
        outs = [{"force": torch.empty((N, 3), dtype=torch.float64).pin_memory().numpy()} for _ in evs]
        lanes.append((evs, outs))
    for rnd in range(3):  # exact sizing, then the sync-free path; the configurations rotate through the sets
        for k, (evs, outs) in enumerate(lanes):
            c = cfgs[(k + rnd) % 3]
            evs[0].set_system(c[0], c[1], c[2], c[3], None, box, True)
            cg.eval_multi_begin(evs, cg.MODE_ION, cg.FORCE | cg.ENERGY, outs)
        for k, (evs, outs) in enumerate(lanes):
            res = cg.eval_multi_end(evs)
            for j, nm in enumerate(pair):
                assert res[j]["force"] is outs[j]["force"]
                errs = compare(res[j], refs[(k + rnd) % 3][j], cg.FORCE | cg.ENERGY)
                assert all(v <= 1.0 for v in errs.values()), (nm, rnd, k, errs)
    with pytest.raises(cg.CgError):
        cg.eval_multi_end(lanes[0][0])  # nothing in flight
    for evs, _ in lanes:
        for ev in evs:
            ev.close()


@pytest.mark.parametrize("n_side,shard", [(14, None), (24, (1, 4))])
def test_two_schemes_in_one_pass_with_j_split(n_side, shard, oracle_kind):
    """The one-pass evaluation where the engine cuts the candidate runs of every group into slices (few groups per resident
    warp: a small system, or one rank's shard of a larger one): per-scheme partial outputs, summed by the epilogue."""
    import torch
    pair = ("wolf", "fennell")
    x, y, z, q, mu = O.generate(n_side, 3.0, 811, oracle_kind)
    box = [3.0 * n_side] * 3
    dev = torch.device("cuda", 0)
    st = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(st)
    try:
        evs = [cg.BatchedEvaluator(zoo()[nm][0]()) for nm in pair]
        for ev in evs:
            ev.set_stream(st.cuda_stream)
        if shard:
            evs[0].set_shard(*shard)
        evs[1].share_system(evs[0])
        dx = [torch.tensor(a, dtype=torch.float64, device=dev) for a in (x, y, z, q)]
        N = x.size
        what = cg.ENERGY | cg.FORCE | cg.POTENTIAL
        refs = [O.Scheme(zoo()[nm][1], oracle_kind).batched(x, y, z, q, None, box, True, cg.MODE_ION, what) for nm in pair]
        for rep in range(2):
            f = [torch.full((N, 3), float("nan"), dtype=torch.float64, device=dev) for _ in evs]
            p = [torch.full((N,), float("nan"), dtype=torch.float64, device=dev) for _ in evs]
            sc = [torch.zeros(2, dtype=torch.float64, device=dev) for _ in evs]
            evs[0].set_system_device(dx[0], dx[1], dx[2], dx[3], None, box, True)
            cg.eval_multi_device(evs, cg.MODE_ION, what, force=f, potential=p, scalars=sc)
            torch.cuda.synchronize()
            assert evs[0].counters()["split"] > 1
            for k, nm in enumerate(pair):
                fk, pk = f[k].cpu().numpy(), p[k].cpu().numpy()
                if shard:  # only the shard's own particles are evaluated: compare those (the one-by-one evaluation tells which)
                    f1 = torch.full((N, 3), float("nan"), dtype=torch.float64, device=dev)
                    s1 = torch.zeros(2, dtype=torch.float64, device=dev)
                    evs[k].eval_device(cg.MODE_ION, what, force=f1, potential=torch.zeros(N, dtype=torch.float64, device=dev), scalars=s1)
                    torch.cuda.synchronize()
                    f1n = f1.cpu().numpy()  # the j-split epilogue writes every row: zeros for the particles of other shards
                    own = np.any(f1n != 0.0, axis=1) & ~np.isnan(f1n[:, 0])
                    assert 0 < own.sum() < N and np.array_equal(own, np.any(fk != 0.0, axis=1) & ~np.isnan(fk[:, 0]))
                    assert s1[1].item() == sc[k][1].item()
                    scale = np.abs(refs[k]["force"]).max()
                    assert np.abs(fk[own] - refs[k]["force"][own]).max() <= 1e-12 * scale
                    assert np.abs(fk[own] - f1.cpu().numpy()[own]).max() <= 1e-13 * scale
                else:
                    g = dict(force=fk, potential=pk, energy=sc[k][0].item(), pairs=int(sc[k][1].item()))
                    errs = compare(g, refs[k], what)
                    assert all(v <= 1.0 for v in errs.values()), (nm, rep, errs)
        for ev in evs:
            ev.close()
    finally:
        torch.cuda.set_stream(torch.cuda.default_stream(dev))


@pytest.mark.parametrize("name,mode,what", [("wolf", cg.MODE_ION, cg.ENERGY | cg.FORCE | cg.FIELD | cg.POTENTIAL),
                                            ("reactionfield", cg.MODE_DIPOLE, cg.ENERGY | cg.FORCE | cg.FIELD),
                                            ("ewald_screened", cg.MODE_MULTIPOLE, cg.ENERGY | cg.FORCE)])
def test_tail_split_of_the_last_round(name, mode, what, oracle_kind):
    """A system with somewhat more i-groups than resident warps: the groups beyond the last full round are cut into row slices
    (kparams.cuh: tail_plan) whose partial outputs the epilogue kernel sums.  Parity of every particle against the oracle, the same
    pair count and (to rounding) the same outputs as with whole groups (CG_TAIL=0), and the item counter shows the extra items."""
    x, y, z, q, mu = O.generate(40, 3.0, 2024 + mode, oracle_kind)
    mk, p = zoo()[name]
    L = 120.0
    res = {}
    for key in ("tail", "whole"):
        if key == "whole":
            os.environ["CG_TAIL"] = "0"
        try:
            res[key] = run_case(mk(), O.Scheme(p, oracle_kind), x, y, z, q, mu, [L] * 3, True, mode, what)
        finally:
            os.environ.pop("CG_TAIL", None)
    g, r = res["tail"]
    gw, _ = res["whole"]
    groups = g["counters"]["groups"]
    assert gw["counters"]["items"] == groups and g["counters"]["split"] == 1
    assert g["counters"]["items"] > groups, (g["counters"], "the tail split did not engage: pick a size with ng mod wave < wave / 2")
    errs = compare(g, r, what)
    assert all(v <= 1.0 for v in errs.values()), errs
    assert g["pairs"] == gw["pairs"]
    for k in ("force", "field", "potential"):
        if k in g and g[k] is not None and k in gw and gw[k] is not None:
            assert np.abs(g[k] - gw[k]).max() <= 1e-12 * np.abs(gw[k]).max()
