"""Scheme objects with the reference's class names and constructor arguments, backed by the C ABI.

Each object owns a `cg_scheme_desc` filled by the matching `cg_scheme_*` constructor of libcg_b200.so, i.e. by the
same C++ code the header classes under include/coulombgalore/ use.  The scalar S(q) methods evaluate on the host
through `cg_srf_host`; the batched pair path is `coulombgalore_b200.BatchedEvaluator`.
"""
import ctypes as C
import math

import numpy as np

from . import _capi
from ._capi import SchemeDesc, check, lib

infinity = math.inf


class SchemeBase:
    """Common surface of all schemes: `cutoff`, `debye_length`, `scheme`, `name`, `doi`, S(q) and derivatives."""
    name = ""
    doi = ""

    def __init__(self):
        self.desc = SchemeDesc()
        self._keep = []  # buffers the descriptor points into (spline tables)

    # public data members of the reference's SchemeBase (core.hpp:148-152)
    @property
    def cutoff(self):
        return self.desc.cutoff

    @property
    def debye_length(self):
        return self.desc.debye_length

    @property
    def scheme(self):
        return self.desc.scheme

    def _srf(self, q):
        qa = np.ascontiguousarray(np.atleast_1d(q), dtype=np.float64)
        out = np.empty((4, qa.size))
        check(lib.cg_srf_host(C.byref(self.desc), qa.size, qa.ctypes.data_as(_capi._dp), out.ctypes.data_as(_capi._dp)),
              "cg_srf_host")
        return out

    def srf(self, q):
        """S, S', S'', S''' at q (array of shape (4, len(q)))."""
        return self._srf(q)

    def short_range_function(self, q):
        r = self._srf(q)[0]
        return r if np.ndim(q) else float(r[0])

    def short_range_function_derivative(self, q):
        r = self._srf(q)[1]
        return r if np.ndim(q) else float(r[0])

    def short_range_function_second_derivative(self, q):
        r = self._srf(q)[2]
        return r if np.ndim(q) else float(r[0])

    def short_range_function_third_derivative(self, q):
        r = self._srf(q)[3]
        return r if np.ndim(q) else float(r[0])

    # O(N) terms (core.hpp:124-145, 179, 809-842)
    def self_energy(self, squared_moments):
        ic = self.desc.inverse_cutoff
        return sum(self.desc.self_energy_prefactor[i] * squared_moments[i] * ic ** (2 * i + 1) for i in range(2))

    def self_field(self, moments):
        ic = self.desc.inverse_cutoff
        out = np.zeros(3)
        for i in range(2):
            out += self.desc.self_field_prefactor[i] * np.asarray(moments[i], dtype=float) * ic ** (2 * i + 1)
        return out

    def calc_dielectric(self, M2V):
        T0 = self.desc.T0
        return (M2V * T0 + 2.0 * M2V + 1.0) / (M2V * T0 - M2V + 1.0)

    def neutralization_energy(self, charges, volume):
        s = float(np.sum(charges))
        return self.desc.chi / 2.0 / volume * s * s


    # serialisation (core.hpp:223-243): the scheme's own keys, then cutoff / doi / type / debyelength
    def _to_json(self):
        return {}

    def to_json(self):
        j = dict(self._to_json())
        if math.isfinite(self.cutoff):
            j["cutoff"] = self.cutoff
        if self.doi:
            j["doi"] = self.doi
        if self.name:
            j["type"] = self.name
        if math.isfinite(self.debye_length):
            j["debyelength"] = self.debye_length
        return j


def _ctor(rc, what):
    if rc != 0:
        # the reference constructors throw std::runtime_error on the same conditions (poisson.hpp:81-86)
        raise ValueError("%s: parameters rejected (%s)" % (what, _capi.STATUS.get(rc, rc)))


class Plain(SchemeBase):
    name, doi = "plain", "Premier mémoire sur l’électricité et le magnétisme by Charles-Augustin de Coulomb"

    def __init__(self, debye_length=infinity):
        super().__init__()
        _ctor(lib.cg_scheme_plain(debye_length, C.byref(self.desc)), "Plain")


class ReactionField(SchemeBase):
    name, doi = "Reaction-field", "10.1080/00268977300102101"

    def __init__(self, cutoff, epsRF, epsr, shifted):
        super().__init__()
        _ctor(lib.cg_scheme_reactionfield(cutoff, epsRF, epsr, int(bool(shifted)), C.byref(self.desc)), "ReactionField")

    def _to_json(self):  # reactionfield.hpp:82-84
        return {"epsr": self.desc.eps_r, "epsRF": self.desc.eps_rf, "shifted": bool(self.desc.shifted)}


class Poisson(SchemeBase):
    name, doi = "poisson", "10/c5fr"

    def __init__(self, cutoff, C_, D, debye_length=infinity):
        super().__init__()
        _ctor(lib.cg_scheme_poisson(cutoff, int(C_), int(D), debye_length, C.byref(self.desc)), "Poisson")

    def _to_json(self):  # poisson.hpp:230
        return {"C": self.desc.C, "D": self.desc.D}


class qPotential(SchemeBase):
    name, doi = "qpotential", "10.1039/c9cp03875b"

    def __init__(self, cutoff, order):
        super().__init__()
        _ctor(lib.cg_scheme_qpotential(cutoff, int(order), C.byref(self.desc)), "qPotential")

    def _to_json(self):  # qpotential.hpp:280
        return {"order": self.desc.order}


class qPotentialFixedOrder5(SchemeBase):
    """qPotentialFixedOrder<5> (qpotential.hpp:195-232)."""
    name, doi = "qpotential", "10.1039/c9cp03875b"

    def __init__(self, cutoff):
        super().__init__()
        _ctor(lib.cg_scheme_qpotential5(cutoff, C.byref(self.desc)), "qPotentialFixedOrder<5>")

    def _to_json(self):  # qpotential.hpp:230
        return {"order": self.desc.order}


class Fanourgakis(SchemeBase):
    name, doi = "fanourgakis", "10.1063/1.3216520"

    def __init__(self, cutoff):
        super().__init__()
        _ctor(lib.cg_scheme_fanourgakis(cutoff, C.byref(self.desc)), "Fanourgakis")


class Fennell(SchemeBase):
    name, doi = "Fennell", "10.1063/1.2206581"

    def __init__(self, cutoff, alpha):
        super().__init__()
        _ctor(lib.cg_scheme_fennell(cutoff, alpha, C.byref(self.desc)), "Fennell")

    def _to_json(self):  # fennell.hpp:80 -- the REDUCED alpha, under its own key
        return {"alpha_red": self.desc.eta}


class ZeroDipole(SchemeBase):
    name, doi = "ZeroDipole", "10.1063/1.3582791"

    def __init__(self, cutoff, alpha):
        super().__init__()
        _ctor(lib.cg_scheme_zerodipole(cutoff, alpha, C.byref(self.desc)), "ZeroDipole")

    def _to_json(self):  # zerodipole.hpp:88 -- the REDUCED alpha (alpha * cutoff) under the key "alpha", as the reference writes it
        return {"alpha": self.desc.eta}


class Zahn(SchemeBase):
    name, doi = "Zahn", "10.1021/jp025949h"

    def __init__(self, cutoff, alpha):
        super().__init__()
        _ctor(lib.cg_scheme_zahn(cutoff, alpha, C.byref(self.desc)), "Zahn")

    def _to_json(self):  # zahn.hpp:83 -- the REDUCED alpha (alpha * cutoff) under the key "alpha", as the reference writes it
        return {"alpha": self.desc.eta}


class Wolf(SchemeBase):
    name, doi = "Wolf", "10.1063/1.478738"

    def __init__(self, cutoff, alpha):
        super().__init__()
        _ctor(lib.cg_scheme_wolf(cutoff, alpha, C.byref(self.desc)), "Wolf")

    def _to_json(self):  # wolf.hpp:78 -- the REDUCED alpha (alpha * cutoff) under the key "alpha", as the reference writes it
        return {"alpha": self.desc.eta}


class Ewald(SchemeBase):
    name, doi = "Ewald real-space", "10.1002/andp.19213690304"

    def __init__(self, cutoff, alpha, surface_dielectric_constant=infinity, debye_length=infinity):
        super().__init__()
        _ctor(lib.cg_scheme_ewald(cutoff, alpha, surface_dielectric_constant, debye_length, C.byref(self.desc)), "Ewald")
        # the member keeps the value as passed: the reference's `< 1 -> infinity` fix-up assigns to the constructor
        # parameter that shadows it (ewald.hpp:131-140), so to_json writes e.g. epss = 0.5 back out
        self.surface_dielectric_constant = surface_dielectric_constant

    def _to_json(self):  # ewald.hpp:203-210
        eps = self.surface_dielectric_constant
        return {"alpha": self.desc.eta / self.cutoff, "epss": "inf" if math.isinf(eps) else eps}


class EwaldTruncated(SchemeBase):
    name, doi = "EwaldT real-space", "XYZ"

    def __init__(self, cutoff, alpha, surface_dielectric_constant=infinity, debye_length=infinity):
        super().__init__()
        _ctor(lib.cg_scheme_ewaldt(cutoff, alpha, surface_dielectric_constant, debye_length, C.byref(self.desc)),
              "EwaldTruncated")
        self.surface_dielectric_constant = surface_dielectric_constant

    def _to_json(self):  # ewald_truncated.hpp:112-119
        eps = self.surface_dielectric_constant
        return {"alpha": self.desc.eta / self.cutoff, "epss": "inf" if math.isinf(eps) else eps}


class Splined(SchemeBase):
    """Splined().spline(Poisson, 12, 4, 3)  ==  reference `Splined pot; pot.spline<Poisson>(12, 4, 3);` (splined.hpp:340-367).

    `spline_from_tables` installs externally generated tables instead (parity runs feed the oracle's tables to both
    sides, SURVEY.md 7 "Splined parity").
    """

    def __init__(self):
        super().__init__()
        self.tolerance = 1e-3  # splined.hpp:340
        self.inner = None

    def setTolerance(self, tol):
        self.tolerance = float(tol)

    def numKnots(self):
        return [int(self.desc.spline_knots[k]) for k in range(4)]

    def spline(self, cls, *args, **kw):
        self.inner = cls(*args, **kw)
        self.name, self.doi = self.inner.name, self.inner.doi  # SchemeBase::operator= (splined.hpp:329)
        storage = np.zeros(int(lib.cg_spline_storage_doubles()))
        _ctor(lib.cg_scheme_splined(C.byref(self.inner.desc), self.tolerance, storage.ctypes.data_as(_capi._dp),
                                    C.byref(self.desc)), "Splined")
        self._keep = [storage]
        return self

    def spline_from_tables(self, inner, tables):
        """tables: [(knots, coefficients)] * 4, as oracle.Scheme.spline_tables() returns them."""
        self.inner = inner
        self.name, self.doi = inner.name, inner.doi
        r2 = [np.ascontiguousarray(t[0], dtype=np.float64) for t in tables]
        c = [np.ascontiguousarray(t[1], dtype=np.float64) for t in tables]
        n = (C.c_int32 * 4)(*[a.size for a in r2])
        pr = (_capi._dp * 4)(*[a.ctypes.data_as(_capi._dp) for a in r2])
        pc = (_capi._dp * 4)(*[a.ctypes.data_as(_capi._dp) for a in c])
        _ctor(lib.cg_scheme_splined_tables(C.byref(inner.desc), n, pr, pc, C.byref(self.desc)), "Splined(tables)")
        self._keep = [r2, c]
        return self

    def tables(self):
        out = []
        for k in range(4):
            n = int(self.desc.spline_knots[k])
            out.append((np.ctypeslib.as_array(self.desc.spline_r2[k], (n,)).copy(),
                        np.ctypeslib.as_array(self.desc.spline_c[k], (6 * (n - 1),)).copy()))
        return out


# ---- run-time factory (reference all.hpp:42-104) ----
def _at(j, key):
    if key not in j:
        raise KeyError("key '%s' not found" % key)  # nlohmann::json::at throws out_of_range
    return j[key]


def _num(v):  # the reference's own to_json writes epss = "inf" as a string (ewald.hpp:206)
    return math.inf if isinstance(v, str) and v.strip().lower() in ("inf", "infinity") else float(v)


_FACTORY = {
    "plain": lambda j: Plain(_num(j.get("debyelength", infinity))),
    "wolf": lambda j: Wolf(_at(j, "cutoff"), _at(j, "alpha")),
    "zahn": lambda j: Zahn(_at(j, "cutoff"), _at(j, "alpha")),
    "fennell": lambda j: Fennell(_at(j, "cutoff"), _at(j, "alpha")),
    "zerodipole": lambda j: ZeroDipole(_at(j, "cutoff"), _at(j, "alpha")),
    "fanourgakis": lambda j: Fanourgakis(_at(j, "cutoff")),
    "qpotential": lambda j: qPotential(_at(j, "cutoff"), _at(j, "order")),
    "ewald": lambda j: Ewald(_at(j, "cutoff"), _at(j, "alpha"), _num(j.get("epss", infinity)), _num(j.get("debyelength", infinity))),
    "ewaldt": lambda j: EwaldTruncated(_at(j, "cutoff"), _at(j, "alpha"), _num(j.get("epss", infinity)),
                                       _num(j.get("debyelength", infinity))),
    "poisson": lambda j: Poisson(_at(j, "cutoff"), _at(j, "C"), _at(j, "D"), _num(j.get("debyelength", infinity))),
    "reactionfield": lambda j: ReactionField(_at(j, "cutoff"), _at(j, "epsRF"), _at(j, "epsr"), _at(j, "shifted")),
    "spline": lambda j: None,  # a known keyword for which the reference's factory returns a null pointer (all.hpp:98-100)
}


def createScheme(j):
    """The reference's createScheme(json) for a dict: j["type"] selects the class, the other keys are those of the
    reference's JSON constructors.  Unknown type -> RuntimeError("unknown coulomb scheme ..."), missing key -> KeyError."""
    name = _at(j, "type")
    if name not in _FACTORY:
        raise RuntimeError("unknown coulomb scheme " + str(name))
    return _FACTORY[name](j)


create_scheme = createScheme
