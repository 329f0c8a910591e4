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


def test_constructor_errors_and_static_use(probe):
    for C_, D in ((0, 3), (3, -2), (3, 0)):
        with pytest.raises(ValueError):
            Host(probe, O.params(O.POISSON, cutoff=12.0, C_=C_, D=D))
    w = O.Scheme(O.params(O.WOLF, cutoff=12.0, alpha=0.25), "port")
    want = w.pair(1.0, -1.0, (0, 0, 0), (0, 0, 0), (5.0, 0, 0))["ion_ion_energy"] + (-1.0) * w.info()["self_energy_ion"]
    assert probe.probe_static_wolf_energy(12.0, 0.25, 1.0, -1.0, 5.0) == pytest.approx(want, rel=1e-13)
