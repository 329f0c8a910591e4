"""The scheme zoo shared by every test: name -> (oracle constructor params, product class name, product ctor args).
Covers all twelve scheme classes of the reference plus the core-Yukawa variants (SURVEY.md 8d "extra parity-only cases")."""
import math

from oracle import oracle as O

INF = math.inf

ZOO = {
    "plain": (dict(scheme=O.PLAIN), "Plain", ()),
    "plain_yukawa": (dict(scheme=O.PLAIN, debye_length=23.0), "Plain", (23.0,)),
    "reactionfield": (dict(scheme=O.REACTIONFIELD, cutoff=12.0, eps_rf=80.0, eps_r=1.0, shifted=False), "ReactionField", (12.0, 80.0, 1.0, False)),
    "reactionfield_shifted": (dict(scheme=O.REACTIONFIELD, cutoff=12.0, eps_rf=80.0, eps_r=1.0, shifted=True), "ReactionField", (12.0, 80.0, 1.0, True)),
    "poisson43": (dict(scheme=O.POISSON, cutoff=12.0, C_=4, D=3), "Poisson", (12.0, 4, 3)),
    "poisson33_yukawa": (dict(scheme=O.POISSON, cutoff=12.0, C_=3, D=3, debye_length=30.0), "Poisson", (12.0, 3, 3, 30.0)),
    "poisson10": (dict(scheme=O.POISSON, cutoff=12.0, C_=1, D=0), "Poisson", (12.0, 1, 0)),
    "poisson1m1": (dict(scheme=O.POISSON, cutoff=12.0, C_=1, D=-1), "Poisson", (12.0, 1, -1)),
    "poisson22": (dict(scheme=O.POISSON, cutoff=12.0, C_=2, D=2), "Poisson", (12.0, 2, 2)),
    "qpotential3": (dict(scheme=O.QPOTENTIAL, cutoff=14.0, order=3), "qPotential", (14.0, 3)),
    "qpotential4": (dict(scheme=O.QPOTENTIAL, cutoff=12.0, order=4), "qPotential", (12.0, 4)),
    "qpotential5_fixed": (dict(scheme=O.QPOTENTIAL5, cutoff=12.0), "qPotentialFixedOrder5", (12.0,)),
    "fanourgakis": (dict(scheme=O.FANOURGAKIS, cutoff=12.0), "Fanourgakis", (12.0,)),
    "wolf": (dict(scheme=O.WOLF, cutoff=12.0, alpha=0.25), "Wolf", (12.0, 0.25)),
    "fennell": (dict(scheme=O.FENNELL, cutoff=12.0, alpha=0.25), "Fennell", (12.0, 0.25)),
    "zerodipole": (dict(scheme=O.ZERODIPOLE, cutoff=12.0, alpha=0.25), "ZeroDipole", (12.0, 0.25)),
    "zahn": (dict(scheme=O.ZAHN, cutoff=12.0, alpha=0.25), "Zahn", (12.0, 0.25)),
    "ewald": (dict(scheme=O.EWALD, cutoff=10.0, alpha=0.3), "Ewald", (10.0, 0.3)),
    "ewald_screened": (dict(scheme=O.EWALD, cutoff=10.0, alpha=0.3, debye_length=30.0), "Ewald", (10.0, 0.3, INF, 30.0)),
    "ewaldt": (dict(scheme=O.EWALDT, cutoff=12.0, alpha=0.25), "EwaldTruncated", (12.0, 0.25)),
}


def oracle_params(name, splined=False, utol=0.0):
    kw = dict(ZOO[name][0])
    return O.params(kw.pop("scheme"), splined=splined, spline_utol=utol, **kw)
