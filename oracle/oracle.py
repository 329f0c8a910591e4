"""ctypes front-end of the CPU oracle (oracle/cgo_api.h).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this module;
nothing under coulombgalore_b200/ does.  `load("ref")` is the unmodified reference compiled into
oracle/_ref/libcg_ref.so, `load("port")` the independent restatement oracle/libcg_port.so, `load("best")`
the reference when its prebuilt library is present and the port otherwise.
"""
import ctypes as C
import math
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
INF = math.inf

# cgo_scheme_id (reference enum Scheme, core.hpp:56-70)
PLAIN, EWALD, EWALDT, REACTIONFIELD, WOLF, POISSON, QPOTENTIAL, FANOURGAKIS, ZERODIPOLE, ZAHN, FENNELL, QPOTENTIAL5, SPLINE = range(13)
MODE_ION, MODE_DIPOLE, MODE_MULTIPOLE = 0, 1, 2
ENERGY, FORCE, FIELD, POTENTIAL = 1, 2, 4, 8
PAIR_OUT = 35


class Params(C.Structure):
    _fields_ = [("scheme", C.c_int32), ("splined", C.c_int32), ("order", C.c_int32), ("C", C.c_int32),
                ("D", C.c_int32), ("shifted", C.c_int32), ("cutoff", C.c_double), ("alpha", C.c_double),
                ("eps_sur", C.c_double), ("debye_length", C.c_double), ("eps_rf", C.c_double),
                ("eps_r", C.c_double), ("spline_utol", C.c_double)]


def params(scheme, cutoff=0.0, alpha=0.0, eps_sur=INF, debye_length=INF, eps_rf=0.0, eps_r=0.0, shifted=False,
           C_=0, D=0, order=0, splined=False, spline_utol=0.0):
    return Params(scheme, int(splined), order, C_, D, int(shifted), cutoff, alpha, eps_sur, debye_length, eps_rf,
                  eps_r, spline_utol)


_dp = C.POINTER(C.c_double)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(_dp)


def lib_path(kind):
    return os.path.join(HERE, "_ref", "libcg_ref.so") if kind == "ref" else os.path.join(HERE, "libcg_port.so")


def available(kind):
    return os.path.exists(lib_path(kind))


_libs = {}


def load(kind="best"):
    if kind == "best":
        kind = "ref" if available("ref") else "port"
    if kind in _libs:
        return _libs[kind]
    lib = C.CDLL(lib_path(kind))
    lib.cgo_kind.restype = C.c_char_p
    lib.cgo_create.argtypes = [C.POINTER(Params), C.POINTER(C.c_void_p)]
    lib.cgo_destroy.argtypes = [C.c_void_p]
    lib.cgo_destroy.restype = None
    lib.cgo_info.argtypes = [C.c_void_p, _dp]
    lib.cgo_srf.argtypes = [C.c_void_p, C.c_int64, _dp, _dp]
    lib.cgo_pair.argtypes = [C.c_void_p, C.c_double, C.c_double] + [_dp] * 6
    lib.cgo_spline_table.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int32), _dp, _dp]
    lib.cgo_generate.argtypes = [C.c_int32, C.c_double, C.c_uint64] + [_dp] * 7
    lib.cgo_self_terms.argtypes = [C.c_void_p, C.c_int64] + [_dp] * 4 + [C.c_double] + [_dp] * 4
    lib.cgo_self_terms.restype = C.c_int
    lib.cgo_generate.restype = None
    lib.cgo_batched.argtypes = ([C.c_void_p, C.c_int64] + [_dp] * 7 + [_dp, C.c_int32, C.c_int32, C.c_int32,
                                C.c_int64, C.c_int64, C.c_int32, C.c_int32] + [_dp] * 8)
    lib.cgo_batched_quad.argtypes = ([C.c_void_p, C.c_int64] + [_dp] * 8 + [_dp, C.c_int32, C.c_int32,
                                     C.c_int64, C.c_int64, C.c_int32, C.c_int32] + [_dp] * 8)
    _libs[kind] = lib
    return lib


class Scheme:
    """One scheme object of the oracle (reference class or its restatement)."""

    def __init__(self, p, kind="best"):
        self.lib = load(kind)
        self.kind = self.lib.cgo_kind().decode()
        self.p = p
        self.h = C.c_void_p()
        rc = self.lib.cgo_create(C.byref(p), C.byref(self.h))
        if rc != 0:
            raise ValueError("oracle rejected scheme parameters (rc=%d)" % rc)

    def __del__(self):
        if getattr(self, "h", None) and self.h.value:
            self.lib.cgo_destroy(self.h)
            self.h = None

    def info(self):
        o = np.zeros(10)
        self.lib.cgo_info(self.h, _ptr(o))
        return dict(cutoff=o[0], debye_length=o[1], self_energy_ion=o[2], self_energy_dipole=o[3],
                    self_field_dipole=o[4], dielectric_m2v1=o[5], neutralization_q1_v1=o[6], scheme=int(o[7]))

    def srf(self, q):
        q = np.ascontiguousarray(q, dtype=np.float64)
        out = np.zeros((4, q.size))
        self.lib.cgo_srf(self.h, q.size, _ptr(q), _ptr(out))
        return out

    def pair(self, zA, zB, muA, muB, r, quadA=None, quadB=None):
        a = np.ascontiguousarray(muA, dtype=np.float64)
        b = np.ascontiguousarray(muB, dtype=np.float64)
        rr = np.ascontiguousarray(r, dtype=np.float64)
        qa = None if quadA is None else np.ascontiguousarray(quadA, dtype=np.float64).reshape(9)
        qb = None if quadB is None else np.ascontiguousarray(quadB, dtype=np.float64).reshape(9)
        o = np.zeros(PAIR_OUT)
        self.lib.cgo_pair(self.h, zA, zB, _ptr(a), _ptr(b), _ptr(qa), _ptr(qb), _ptr(rr), _ptr(o))
        return dict(ion_potential=o[0], dipole_potential=o[1], quadrupole_potential=o[2], ion_field=o[3:6],
                    dipole_field=o[6:9], quadrupole_field=o[9:12], multipole_field=o[12:15], ion_ion_energy=o[15],
                    ion_dipole_energy=o[16], dipole_dipole_energy=o[17], ion_quadrupole_energy=o[18],
                    multipole_multipole_energy=o[19], ion_ion_force=o[20:23], ion_dipole_force=o[23:26],
                    dipole_dipole_force=o[26:29], ion_quadrupole_force=o[29:32],
                    multipole_multipole_force=o[32:35], raw=o)

    def spline_tables(self):
        """[(r2, c)] * 4 for a splined scheme: knots and 6 coefficients per interval (splined.hpp:60-66)."""
        out = []
        for k in range(4):
            n = C.c_int32(0)
            if self.lib.cgo_spline_table(self.h, k, C.byref(n), None, None) != 0:
                raise ValueError("not a splined scheme")
            r2 = np.zeros(n.value)
            c = np.zeros(6 * (n.value - 1))
            self.lib.cgo_spline_table(self.h, k, C.byref(n), _ptr(r2), _ptr(c))
            out.append((r2, c))
        return out

    def batched(self, x, y, z, q, mu, box, pbc, mode, what, i_begin=0, i_end=None, nthreads=0, scales=True, quad=None):
        """quad: optional (N, 3, 3) / (N, 9) quadrupole tensors; multipole mode only (core.hpp:320, 457, 581, 737)."""
        N = x.size
        i_end = N if i_end is None else i_end
        n = i_end - i_begin
        res = {}
        f = np.zeros((n, 3)) if what & FORCE else None
        e = np.zeros((n, 3)) if what & FIELD else None
        p = np.zeros(n) if what & POTENTIAL else None
        u = np.zeros(n) if what & ENERGY else None
        fs = np.zeros(n) if (what & FORCE and scales) else None
        es = np.zeros(n) if (what & FIELD and scales) else None
        ps = np.zeros(n) if (what & POTENTIAL and scales) else None
        tot = np.zeros(3)
        boxa = np.ascontiguousarray(box, dtype=np.float64)
        mx, my, mz = (None, None, None) if mu is None else mu
        if quad is not None:
            if mode not in (MODE_MULTIPOLE, 3):
                raise ValueError("quadrupoles need the multipole mode")
            qd = np.ascontiguousarray(quad, dtype=np.float64).reshape(N, 9)
            rc = self.lib.cgo_batched_quad(self.h, N, _ptr(x), _ptr(y), _ptr(z), _ptr(q), _ptr(mx), _ptr(my), _ptr(mz), _ptr(qd),
                                           _ptr(boxa), int(pbc), what, i_begin, i_end, nthreads, int(scales), _ptr(f),
                                           _ptr(e), _ptr(p), _ptr(u), _ptr(fs), _ptr(es), _ptr(ps), _ptr(tot))
        else:
            rc = self.lib.cgo_batched(self.h, N, _ptr(x), _ptr(y), _ptr(z), _ptr(q), _ptr(mx), _ptr(my), _ptr(mz),
                                      _ptr(boxa), int(pbc), mode, what, i_begin, i_end, nthreads, int(scales), _ptr(f),
                                      _ptr(e), _ptr(p), _ptr(u), _ptr(fs), _ptr(es), _ptr(ps), _ptr(tot))
        if rc != 0:
            raise RuntimeError("cgo_batched failed rc=%d" % rc)
        res.update(force=f, field=e, potential=p, energy_i=u, force_scale=fs, field_scale=es, potential_scale=ps,
                   energy=tot[0], energy_abs=tot[1], pairs=int(tot[2]))
        return res


def _self_terms(self, q, mu, volume=0.0, field=None):
    """O(N) terms: dict(self_energy, self_energy_abs, neutralization_energy, charge_sum, self_field (N,3), torque (N,3))."""
    N = (q if q is not None else mu[0]).size
    mx, my, mz = (None, None, None) if mu is None else mu
    sc = np.zeros(4)
    sf = np.zeros((N, 3))
    tq = np.zeros((N, 3)) if field is not None else None
    fld = None if field is None else np.ascontiguousarray(field, dtype=np.float64)
    rc = self.lib.cgo_self_terms(self.h, N, _ptr(q), _ptr(mx), _ptr(my), _ptr(mz), C.c_double(volume), _ptr(fld), _ptr(sc), _ptr(sf),
                                 _ptr(tq))
    if rc != 0:
        raise RuntimeError("cgo_self_terms failed rc=%d" % rc)
    return dict(self_energy=sc[0], self_energy_abs=sc[1], neutralization_energy=sc[2], charge_sum=sc[3], self_field=sf, torque=tq)


Scheme.self_terms = _self_terms


class Reciprocal:
    """Reciprocal-space Ewald of the oracle: the reference's ReciprocalEwaldGaussian (kind "ref") or its restatement."""

    def __init__(self, cutoff, alpha, eps_sur, kappa, kc, box, kind="best"):
        self.lib = load(kind)
        L = self.lib
        L.cgo_recip_create.argtypes = [C.c_double] * 4 + [C.c_int32, _dp, C.POINTER(C.c_void_p)]
        L.cgo_recip_destroy.argtypes = [C.c_void_p]
        L.cgo_recip_destroy.restype = None
        L.cgo_recip_numk.argtypes = [C.c_void_p]
        L.cgo_recip_numk.restype = C.c_int32
        L.cgo_recip_kvectors.argtypes = [C.c_void_p, _dp, _dp]
        L.cgo_recip_update.argtypes = [C.c_void_p, C.c_int64] + [_dp] * 7 + [C.c_int32]
        L.cgo_recip_q.argtypes = [C.c_void_p, _dp]
        L.cgo_recip_energy.argtypes = [C.c_void_p, _dp]
        L.cgo_recip_force_field.argtypes = [C.c_void_p, C.c_int64] + [_dp] * 10
        L.cgo_recip_surface.argtypes = [C.c_void_p, C.c_int64] + [_dp] * 9
        self.h = C.c_void_p()
        b = np.ascontiguousarray(box, dtype=np.float64)
        if L.cgo_recip_create(cutoff, alpha, eps_sur, kappa, kc, _ptr(b), C.byref(self.h)) != 0:
            raise ValueError("cgo_recip_create failed")
        self.K = int(L.cgo_recip_numk(self.h))

    def __del__(self):
        if getattr(self, "h", None) and self.h.value:
            self.lib.cgo_recip_destroy(self.h)
            self.h = None

    def kvectors(self):
        k, a = np.zeros((self.K, 3)), np.zeros(self.K)
        self.lib.cgo_recip_kvectors(self.h, _ptr(k), _ptr(a))
        return k, a

    @staticmethod
    def _mu(mu):
        return (None, None, None) if mu is None else mu

    def update(self, x, y, z, q, mu, sign=0):
        mx, my, mz = self._mu(mu)
        self.lib.cgo_recip_update(self.h, x.size, _ptr(x), _ptr(y), _ptr(z), _ptr(q), _ptr(mx), _ptr(my), _ptr(mz), sign)

    def q(self):
        out = np.zeros(2 * self.K)
        self.lib.cgo_recip_q(self.h, _ptr(out))
        return out[0::2] + 1j * out[1::2]

    def energy(self):
        e = np.zeros(1)
        self.lib.cgo_recip_energy(self.h, _ptr(e))
        return float(e[0])

    def force_field(self, x, y, z, q, mu):
        mx, my, mz = self._mu(mu)
        f, e, s = np.zeros((x.size, 3)), np.zeros((x.size, 3)), np.zeros(x.size)
        self.lib.cgo_recip_force_field(self.h, x.size, _ptr(x), _ptr(y), _ptr(z), _ptr(q), _ptr(mx), _ptr(my), _ptr(mz), _ptr(f), _ptr(e),
                                       _ptr(s))
        return dict(force=f, field=e, force_scale=s)

    def surface(self, x, y, z, q, mu):
        mx, my, mz = self._mu(mu)
        e, f = np.zeros(1), np.zeros(3)
        self.lib.cgo_recip_surface(self.h, x.size, _ptr(x), _ptr(y), _ptr(z), _ptr(q), _ptr(mx), _ptr(my), _ptr(mz), _ptr(e), _ptr(f))
        return float(e[0]), f


def generate(n, a, seed, kind="best"):
    """Jittered-lattice system of SURVEY.md 8d -> x, y, z, q, (mux, muy, muz)."""
    lib = load(kind)
    N = n ** 3
    arrs = [np.zeros(N) for _ in range(7)]
    lib.cgo_generate(n, a, seed, *[_ptr(v) for v in arrs])
    return arrs[0], arrs[1], arrs[2], arrs[3], (arrs[4], arrs[5], arrs[6])


def ewaldt_surface_energy(cutoff, alpha, eps_sur, x, y, z, q, mu, volume, kind="best"):
    """EwaldTruncated(cutoff, alpha, eps_sur).surface_energy(positions, charges, dipoles, volume) (reference ewald_truncated.hpp:91-104)."""
    lib = load(kind)
    lib.cgo_ewaldt_surface.argtypes = [C.c_double] * 3 + [C.c_int64] + [_dp] * 7 + [C.c_double, _dp]
    mx, my, mz = (None, None, None) if mu is None else mu
    e = np.zeros(1)
    rc = lib.cgo_ewaldt_surface(cutoff, alpha, eps_sur, x.size, _ptr(x), _ptr(y), _ptr(z), _ptr(q), _ptr(mx), _ptr(my), _ptr(mz), volume, _ptr(e))
    if rc != 0:
        raise RuntimeError("cgo_ewaldt_surface failed rc=%d" % rc)
    return float(e[0])


def qpochhammer(q, l, P, kind="best"):
    """(qPochhammerSymbol, ...Derivative, ...SecondDerivative, ...ThirdDerivative)(q, l, P) (reference qpotential.hpp:50-192)."""
    lib = load(kind)
    lib.cgo_qpochhammer.argtypes = [C.c_double, C.c_int32, C.c_int32, _dp]
    o = np.zeros(4)
    lib.cgo_qpochhammer(q, l, P, _ptr(o))
    return o
