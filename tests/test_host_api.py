"""The header-only C++ scheme classes (include/coulombgalore/*.hpp) against the oracle, per call.

tests/cpp/host_api_probe.cpp is user-style code written against the reference's API (class names, constructors,
virtual SchemeBase interface); it is compiled here twice -- with the built-in vec3/mat33 and with an Eigen on the
include path (the oracle's stand-in) -- and every per-pair function is compared with the oracle to 1e-13."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import kat_reference as K
from oracle import oracle as O
from zoo import ZOO, oracle_params

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_dp = C.POINTER(C.c_double)


def _f(a):
    return None if a is None else a.ctypes.data_as(_dp)


@pytest.fixture(scope="module", params=["builtin", "eigen"])
def probe(request, tmp_path_factory):
    out = str(tmp_path_factory.mktemp("probe") / ("probe_%s.so" % request.param))
    cmd = ["/usr/bin/g++", "-std=c++17", "-O2", "-Wall", "-Wextra", "-Werror", "-ffp-contract=off", "-fPIC", "-shared",
           "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "host_api_probe.cpp"), "-o", out]
    if request.param == "eigen":
        cmd.insert(-3, "-I" + os.path.join(ROOT, "oracle", "eigen_shim"))
    subprocess.check_call(cmd)
    lib = C.CDLL(out)
    lib.probe_vector_backend.restype = C.c_char_p
    assert lib.probe_vector_backend().decode() == request.param
    lib.probe_create.argtypes = [C.c_int, C.c_int] + [C.c_double] * 6 + [C.c_int] * 4 + [C.c_double, C.POINTER(C.c_void_p)]
    lib.probe_destroy.argtypes = [C.c_void_p]
    lib.probe_info.argtypes = [C.c_void_p, _dp]
    lib.probe_name.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
    lib.probe_srf.argtypes = [C.c_void_p, C.c_int64, _dp, _dp]
    lib.probe_pair.argtypes = [C.c_void_p, C.c_double, C.c_double] + [_dp] * 6
    lib.probe_static_wolf_energy.restype = C.c_double
    lib.probe_static_wolf_energy.argtypes = [C.c_double] * 5
    return lib


class Host:
    def __init__(self, lib, p):
        self.lib, self.h = lib, C.c_void_p()
        rc = lib.probe_create(p.scheme, p.splined, p.cutoff, p.alpha, p.eps_sur, p.debye_length, p.eps_rf, p.eps_r, p.shifted, p.C, p.D,
                              p.order, p.spline_utol, C.byref(self.h))
        if rc != 0:
            raise ValueError(rc)

    def __del__(self):
        if self.h.value:
            self.lib.probe_destroy(self.h)

    def info(self):
        o = np.zeros(10)
        self.lib.probe_info(self.h, _f(o))
        return o

    def name(self):
        b = C.create_string_buffer(64)
        self.lib.probe_name(self.h, b, 64)
        return b.value.decode()

    def srf(self, q):
        q = np.ascontiguousarray(q, dtype=float)
        o = np.zeros((4, q.size))
        self.lib.probe_srf(self.h, q.size, _f(q), _f(o))
        return o

    def pair(self, zA, zB, a, b, r, qa=None, qb=None):
        arr = [np.ascontiguousarray(v, dtype=float).ravel() if v is not None else None for v in (a, b, qa, qb, r)]
        o = np.zeros(35)
        self.lib.probe_pair(self.h, zA, zB, *[_f(v) for v in arr], _f(o))
        return o


SLOTS = dict(ion_potential=0, dipole_potential=1, quadrupole_potential=2, ion_field=3, dipole_field=6, quadrupole_field=9, multipole_field=12,
             ion_ion_energy=15, ion_dipole_energy=16, dipole_dipole_energy=17, ion_quadrupole_energy=18, multipole_multipole_energy=19,
             ion_ion_force=20, ion_dipole_force=23, dipole_dipole_force=26, ion_quadrupole_force=29, multipole_multipole_force=32)


def test_reference_kats_through_the_header_classes(probe):
    cache = {}
    for entry in K.PAIR:
        name, args, slot, comp, v = entry[:5]
        tol = entry[5] if len(entry) > 5 else K.RTOL
        h = cache.setdefault(name, Host(probe, K.SCHEMES[name]))
        o = h.pair(args.get("zA", K.ZA), args.get("zB", K.ZB), args.get("muA", K.MUA), args.get("muB", K.MUB), args["r"], args.get("quadA"),
                   args.get("quadB"))
        got = o[SLOTS[slot] + (comp or 0)]
        assert abs(got - v) <= tol * abs(v), (name, slot, comp, got, v)
    for name, q, vals in K.SRF:
        s = Host(probe, K.SCHEMES[name]).srf([q])[:, 0]
        tol = K.SRF_RTOL.get((name, q), K.RTOL)
        for k, v in enumerate(vals):
            if v is not None:
                assert abs(s[k] - v) <= tol * abs(v) + 1e-9 * (v == 0.0), (name, q, k, s[k], v)


@pytest.mark.parametrize("name", sorted(ZOO))
def test_every_pair_function_against_the_oracle(probe, name, oracle_kind):
    p = oracle_params(name)
    h, o = Host(probe, p), O.Scheme(p, oracle_kind)
    io, ih = o.info(), h.info()
    ref_info = np.array([io[k] for k in ("cutoff", "debye_length", "self_energy_ion", "self_energy_dipole", "self_field_dipole",
                                         "dielectric_m2v1", "neutralization_q1_v1", "scheme")], dtype=float)
    fin = np.isfinite(ref_info)
    assert np.array_equal(np.isfinite(ih[:8]), fin) and np.allclose(ih[:8][fin], ref_info[fin], rtol=1e-13, atol=0)
    q = np.arange(1, 512) / 512.0
    so, sh = o.srf(q), h.srf(q)
    assert np.max(np.abs(so - sh) / np.maximum(np.max(np.abs(so), axis=1, keepdims=True), 1e-300)) <= 1e-13
    rng = np.random.default_rng(7)
    rc = io["cutoff"] if io["cutoff"] < 1e10 else 20.0
    plain = O.Scheme(O.params(O.PLAIN), oracle_kind)
    for t in range(64):
        r = rng.normal(size=3)
        r *= rng.uniform(0.05, 1.1) * rc / np.linalg.norm(r)
        a, b, qa, qb = rng.normal(size=3), rng.normal(size=3), rng.normal(size=(3, 3)), rng.normal(size=(3, 3))
        ro = o.pair(1.3, -0.7, a, b, r, qa, qb)["raw"]
        rh = h.pair(1.3, -0.7, a, b, r, qa, qb)
        # |delta| <= 1e-12 |ref| + 3e-14 * (magnitude of the bare Coulomb terms of this pair): S -> 0 at the cut-off and
        # u_mm passes through zero, where the reference's own relative accuracy degrades ("tolerance definition")
        bare = np.max(np.abs(plain.pair(1.3, -0.7, a, b, r, qa, qb)["raw"]))
        assert np.all(np.abs(ro - rh) <= 1e-12 * np.abs(ro) + 3e-14 * bare), (name, t, np.max(np.abs(ro - rh)), bare)
        if io["cutoff"] < 1e10 and np.linalg.norm(r) >= rc:
            assert np.all(rh == 0.0)  # beyond the cut-off every function returns exactly zero


@pytest.mark.parametrize("name", ["poisson43", "wolf", "ewald_screened", "plain", "qpotential3"])
def test_splined_wrapper(probe, name, oracle_kind):
    """Splined().spline<T>(...): same knots as the reference's Andrea tables, S..S''' within the table tolerance of the analytic
    scheme, base data copied from the wrapped scheme (splined.hpp:329; reference test :1037-1038, :1061-1062)."""
    p = oracle_params(name, splined=True)
    h, o = Host(probe, p), O.Scheme(p, oracle_kind)
    assert int(h.info()[7]) == o.info()["scheme"] == p.scheme and int(h.info()[8]) == 12
    assert h.name() == {"poisson43": "poisson", "wolf": "Wolf", "ewald_screened": "Ewald real-space", "plain": "plain", "qpotential3": "qpotential"}[name]
    q = np.arange(1, 256) / 256.0
    so, sh = o.srf(q), h.srf(q)
    scale = np.maximum(np.max(np.abs(so), axis=1, keepdims=True), 1.0)
    assert np.max(np.abs(so - sh) / scale) <= 1e-5  # table-vs-table: generation is finite-difference sensitive
    analytic = O.Scheme(oracle_params(name), oracle_kind).srf(q)
    assert np.max(np.abs(sh - analytic)) <= 2.1e-3


def test_qpochhammer_free_functions(probe, oracle_kind):
    """qPochhammerSymbol and its three derivative functions (reference qpotential.hpp:50-192): the reference's own known
    answers (test/unittests.cpp:66-85, default doctest tolerance) and the live oracle on a (q, l, P) grid."""
    probe.probe_qpochhammer.argtypes = [C.c_double, C.c_int, C.c_int, _dp]
    probe.probe_qpochhammer.restype = None

    def f(q, l, P):
        o = np.zeros(4)
        probe.probe_qpochhammer(q, l, P, _f(o))
        return o
    kat = [  # (q, l, P, derivative order, value)  unittests.cpp:68-84
        (0.5, 0, 0, 0, 1.0), (0.0, 0, 1, 0, 1.0), (1.0, 0, 1, 0, 0.0), (1.0, 1, 2, 0, 0.0), (0.75, 0, 2, 0, 0.109375),
        (2.0 / 3.0, 2, 5, 0, 0.4211104676), (0.125, 1, 1, 0, 0.984375), (0.75, 0, 2, 1, -0.8125), (2.0 / 3.0, 2, 5, 1, -2.538458169),
        (0.125, 1, 1, 1, -0.25), (0.75, 0, 2, 2, 2.5), (2.0 / 3.0, 2, 5, 2, -1.444601767), (0.125, 1, 1, 2, -2.0), (0.75, 0, 2, 3, 6.0),
        (2.0 / 3.0, 2, 5, 3, 92.48631425), (0.125, 1, 1, 3, 0.0), (0.4, 3, 7, 3, -32.80472205)]
    eps = 100 * np.finfo(np.float32).eps  # doctest::Approx default
    for q, l, P, k, v in kat:
        got = f(q, l, P)[k]
        assert abs(got - v) < eps * (1.0 + max(abs(got), abs(v))), (q, l, P, k, got, v)
    for l in (0, 1, 2, 5):
        for P in (1, 2, 3, 6, 12):
            for q in (0.05, 0.3, 0.5, 0.77, 0.93):
                ref, got = O.qpochhammer(q, l, P, oracle_kind), f(q, l, P)
                # the reference forms the derivatives from log-derivative sums; 1e-12 of the largest term of each sum
                scale = np.maximum(np.abs(ref), np.abs(O.qpochhammer(q, l, P, oracle_kind)).max())
                assert np.all(np.abs(got - ref) <= 1e-12 * scale * (1 + P * (P + l)) ** 2), (q, l, P, got, ref)


def test_constructor_errors_and_static_use(probe):
    for C_, D in ((0, 3), (3, -2), (3, 0)):
        with pytest.raises(ValueError):
            Host(probe, O.params(O.POISSON, cutoff=12.0, C_=C_, D=D))
    w = O.Scheme(O.params(O.WOLF, cutoff=12.0, alpha=0.25), "port")
    want = w.pair(1.0, -1.0, (0, 0, 0), (0, 0, 0), (5.0, 0, 0))["ion_ion_energy"] + (-1.0) * w.info()["self_energy_ion"]
    assert probe.probe_static_wolf_energy(12.0, 0.25, 1.0, -1.0, 5.0) == pytest.approx(want, rel=1e-13)


def test_runtime_factory_and_serialisation():
    """createScheme / to_json of the reference (all.hpp:42-104, core.hpp:223-243), incl. its quirks: unknown keyword
    throws, "spline" yields nothing, Wolf-like schemes serialise the REDUCED alpha under "alpha", Fennell under
    "alpha_red", Ewald writes an infinite surface dielectric as the string "inf"."""
    import ctypes as C
    import coulombgalore_b200 as cg
    from coulombgalore_b200 import _capi
    cases = [
        (dict(type="plain"), cg.Plain()),
        (dict(type="plain", debyelength=23.0), cg.Plain(23.0)),
        (dict(type="wolf", cutoff=12.0, alpha=0.25), cg.Wolf(12.0, 0.25)),
        (dict(type="zahn", cutoff=12.0, alpha=0.25), cg.Zahn(12.0, 0.25)),
        (dict(type="fennell", cutoff=12.0, alpha=0.25), cg.Fennell(12.0, 0.25)),
        (dict(type="zerodipole", cutoff=12.0, alpha=0.25), cg.ZeroDipole(12.0, 0.25)),
        (dict(type="fanourgakis", cutoff=11.0), cg.Fanourgakis(11.0)),
        (dict(type="qpotential", cutoff=14.0, order=3), cg.qPotential(14.0, 3)),
        (dict(type="ewald", cutoff=10.0, alpha=0.3), cg.Ewald(10.0, 0.3)),
        (dict(type="ewald", cutoff=10.0, alpha=0.3, epss=80.0, debyelength=30.0), cg.Ewald(10.0, 0.3, 80.0, 30.0)),
        (dict(type="ewaldt", cutoff=10.0, alpha=0.3, epss="inf"), cg.EwaldTruncated(10.0, 0.3)),
        (dict(type="poisson", cutoff=12.0, C=4, D=3), cg.Poisson(12.0, 4, 3)),
        (dict(type="poisson", cutoff=12.0, C=3, D=3, debyelength=30.0), cg.Poisson(12.0, 3, 3, 30.0)),
        (dict(type="reactionfield", cutoff=12.0, epsRF=80.0, epsr=1.0, shifted=True), cg.ReactionField(12.0, 80.0, 1.0, True)),
    ]
    for j, direct in cases:
        made = cg.createScheme(j)
        assert type(made) is type(direct) and bytes(made.desc) == bytes(direct.desc), j
        # the dependency-free C-ABI factory builds the same descriptor
        keys = [k for k in j if k != "type"]
        vals = (C.c_double * max(len(keys), 1))(*[float("inf") if j[k] == "inf" else float(j[k]) for k in keys])
        karr = (C.c_char_p * max(len(keys), 1))(*[k.encode() for k in keys])
        d = _capi.SchemeDesc()
        assert _capi.lib.cg_scheme_create(j["type"].encode(), len(keys), karr, vals, C.byref(d)) == 0, j
        assert bytes(d) == bytes(direct.desc), j
    with pytest.raises(RuntimeError, match="unknown coulomb scheme"):
        cg.createScheme(dict(type="coulomb"))
    with pytest.raises(KeyError):
        cg.createScheme(dict(type="wolf", cutoff=12.0))
    with pytest.raises(ValueError):
        cg.createScheme(dict(type="poisson", cutoff=12.0, C=0, D=3))  # the constructor's own check (poisson.hpp:81-86)
    assert cg.createScheme(dict(type="spline")) is None
    d = _capi.SchemeDesc()
    assert _capi.lib.cg_scheme_create(b"wolf", 0, None, None, C.byref(d)) == -1
    assert _capi.lib.cg_scheme_create(b"nonsense", 0, None, None, C.byref(d)) == -1
    assert _capi.lib.cg_scheme_create(b"spline", 0, None, None, C.byref(d)) == -5
    # serialisation
    assert cg.Wolf(12.0, 0.25).to_json() == {"alpha": 3.0, "cutoff": 12.0, "doi": "10.1063/1.478738", "type": "Wolf"}
    assert cg.Fennell(12.0, 0.25).to_json()["alpha_red"] == 3.0
    assert cg.Poisson(12.0, 3, 3, 30.0).to_json() == {"C": 3, "D": 3, "cutoff": 12.0, "doi": "10/c5fr", "type": "poisson", "debyelength": 30.0}
    je = cg.Ewald(10.0, 0.3, cg.infinity, 30.0).to_json()
    assert je["epss"] == "inf" and abs(je["alpha"] - 0.3) < 1e-15 and "debyelength" not in je  # Ewald hides kappa from the base
    assert cg.Ewald(10.0, 0.3, 80.0).to_json()["epss"] == 80.0
    jr = cg.ReactionField(12.0, 80.0, 1.0, False).to_json()
    assert (jr["epsr"], jr["epsRF"], jr["shifted"], jr["type"]) == (1.0, 80.0, False, "Reaction-field")
    assert cg.qPotential(14.0, 3).to_json()["order"] == 3 and cg.qPotentialFixedOrder5(14.0).to_json()["order"] == 5
    jp = cg.Plain().to_json()
    assert jp["type"] == "plain" and jp["cutoff"] > 1e150 and "debyelength" not in jp
    # round trip where the reference's own keys allow it
    back = cg.createScheme(cg.Poisson(12.0, 4, 3).to_json())
    assert bytes(back.desc) == bytes(cg.Poisson(12.0, 4, 3).desc)


def test_cpp_factory(probe):
    """createScheme(type, {key: value}) of include/coulombgalore/all.hpp: same schemes as direct construction."""
    probe.probe_factory.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_char_p), _dp, C.POINTER(C.c_void_p)]

    def make(type_, **kv):
        keys = (C.c_char_p * max(len(kv), 1))(*[k.encode() for k in kv])
        vals = (C.c_double * max(len(kv), 1))(*[float(v) for v in kv.values()])
        h = C.c_void_p()
        return probe.probe_factory(type_.encode(), len(kv), keys, vals, C.byref(h)), h

    q = np.linspace(0.05, 0.95, 7)
    for name, kv in (("wolf", dict(cutoff=12.0, alpha=0.25)), ("poisson43", dict(cutoff=12.0, C=4, D=3)),
                     ("reactionfield", dict(cutoff=12.0, epsRF=80.0, epsr=1.0, shifted=0)),
                     ("ewald_screened", dict(cutoff=10.0, alpha=0.3, debyelength=30.0)), ("qpotential3", dict(cutoff=14.0, order=3))):
        p = oracle_params(name)
        type_ = {"poisson43": "poisson", "ewald_screened": "ewald", "qpotential3": "qpotential"}.get(name, name)
        rc, h = make(type_, **kv)
        assert rc == 0, name
        made, direct = Host.__new__(Host), Host(probe, p)
        made.lib, made.h = probe, h
        assert np.array_equal(made.srf(q), direct.srf(q)) and np.array_equal(made.info(), direct.info()), name
    assert make("spline")[0] == 1 and make("coulomb")[0] == 2 and make("wolf", cutoff=12.0)[0] == 3
