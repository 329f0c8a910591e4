"""Known-answer tests transcribed from the reference's own unit tests (reference test/unittests.cpp, line numbers cited).

Only the literal constants are taken; each entry says which reference call produced it.  SURVEY.md 4 established that the
literals hold to <= 1e-8 relative (two of them, :847 and :906, to <= 1e-6), far tighter than the doctest Approx the reference
uses; they are applied here with a purely relative tolerance RTOL (1e-7) unless an entry carries its own.
The per-pair entries address slots of `cgo_pair` (oracle/cgo_api.h): pair(zA, zB, muA, muB, r, quadA, quadB).
"""
import math

from oracle import oracle as O

RTOL = 1e-7
INF = math.inf

ZA, ZB = 2.0, 3.0
MUA, MUB = (19.0, 7.0, 11.0), (13.0, 17.0, 5.0)
QUADA = ((3.0, 7.0, 8.0), (5.0, 9.0, 6.0), (2.0, 1.0, 4.0))
R = (23.0, 0.0, 0.0)
RQ = (5.75 * math.sqrt(6.0), 5.75 * math.sqrt(2.0), 11.5 * math.sqrt(2.0))


def neg(v):
    return tuple(-x for x in v)


SCHEMES = {
    "plain": O.params(O.PLAIN),                                                        # :137
    "plain_yukawa": O.params(O.PLAIN, debye_length=23.0),                              # :299
    "wolf": O.params(O.WOLF, cutoff=29.0, alpha=0.1),                                  # :330
    "zahn": O.params(O.ZAHN, cutoff=29.0, alpha=0.1),                                  # :348
    "fennell": O.params(O.FENNELL, cutoff=29.0, alpha=0.1),                            # :366
    "zerodipole": O.params(O.ZERODIPOLE, cutoff=29.0, alpha=0.1),                      # :384
    "rf": O.params(O.REACTIONFIELD, cutoff=29.0, eps_rf=80.0, eps_r=1.0, shifted=False),   # :404
    "rf_shifted": O.params(O.REACTIONFIELD, cutoff=29.0, eps_rf=80.0, eps_r=1.0, shifted=True),  # :418
    "qpotential4": O.params(O.QPOTENTIAL, cutoff=29.0, order=4),                       # :436
    "fanourgakis": O.params(O.FANOURGAKIS, cutoff=29.0),                               # :460
    "ewald": O.params(O.EWALD, cutoff=29.0, alpha=0.1),                                # :527
    "ewald_screened": O.params(O.EWALD, cutoff=29.0, alpha=0.1, debye_length=23.0),    # :547
    "ewald_lit": O.params(O.EWALD, cutoff=5.0, alpha=0.8),                             # :562-566 (box 10, cutoff 5, alpha 8/10)
    "ewaldt": O.params(O.EWALDT, cutoff=29.0, alpha=0.1),                              # :699
    "poisson33": O.params(O.POISSON, cutoff=29.0, C_=3, D=3),                          # :722
    "poisson43": O.params(O.POISSON, cutoff=29.0, C_=4, D=3),                          # :757
    "poisson33_yukawa": O.params(O.POISSON, cutoff=29.0, C_=3, D=3, debye_length=23.0),  # :838
    "splined_plain": O.params(O.PLAIN, splined=True),                                  # :965
}

# (scheme, q, (S, S', S'', S''') with None = not pinned by the reference)
SRF = [
    ("plain", 0.5, (1.0, 0.0, 0.0, 0.0)),                                              # :143-146
    ("wolf", 0.5, (0.04028442542, -0.3997546829, 3.361591250, -21.54779991)),          # :336-339
    ("wolf", 1.0, (0.0, None, None, None)),                                            # :340
    ("zahn", 0.5, (0.04049737684, -0.3997135850, 3.360052030, -21.54779991)),          # :354-357
    ("zahn", 1.0, (0.0000410979, None, None, None)),                                   # :358 (7 digits given)
    ("fennell", 0.5, (0.04009202308, -0.3997546830, 3.363130468, -21.54779991)),       # :372-375
    ("zerodipole", 0.5, (0.04014012381, -0.3998508840, 3.362745664, -21.54549108)),    # :390-393
    ("rf", 0.5, (1.061335404, 0.3680124224, 1.472049689, 2.944099379)),                # :410-413
    ("rf", 1.0, (1.490683230, None, None, None)),                                      # :414
    ("rf_shifted", 0.5, (0.3159937888, -1.122670807, 1.472049689, 2.944099379)),       # :424-427
    ("qpotential4", 0.5, (0.3076171875, -1.453125, 1.9140625, 17.25)),                 # :442-445
    ("qpotential4", 0.0, (1.0, -1.0, -2.0, 0.0)),                                      # :450-453
    ("fanourgakis", 0.5, (0.1992187500, -1.1484375, 3.28125, 6.5625)),                 # :466-469
    ("ewald", 0.5, (0.04030497436, -0.399713585, 3.36159125, -21.54779991)),           # :533-536
    ("ewald_screened", 0.5, (0.07306333588, -0.63444119, 4.423133599, -19.85937171)),  # :551-554
    ("ewaldt", 0.5, (0.03993035150, -0.3992923727, 3.364180355, -21.56439607)),        # :705-708
    ("poisson33", 0.5, (0.15625, -1.0, 3.75, 0.0)),                                    # :728-731
    ("poisson33", 0.6, (None, None, None, -5.76)),                                     # :732
    ("poisson33", 0.0, (1.0, -2.0, 0.0, 0.0)),                                         # :737-740
    ("poisson43", 0.5, (0.19921875, -1.1484375, 3.28125, 6.5625)),                     # :760-763
    ("poisson33_yukawa", 0.5, (0.5673222034, -1.437372757, -2.552012334, 4.384434209)),  # :844-847 (:847 to 1e-6)
    ("splined_plain", 0.5, (1.0, 0.0, 0.0, 0.0)),                                      # :971-974
]
SRF_RTOL = {("poisson33_yukawa", 0.5): 1e-6, ("zahn", 1.0): 1e-5}

# self_energy({z2, mu2}) literals: (scheme, z2, mu2, value)
SELF_ENERGY = [
    ("wolf", 4.0, 0.0, -0.2256786677), ("wolf", 0.0, 2.0, -0.0007522527778),           # :332-333
    ("zahn", 4.0, 0.0, -0.2256227568),                                                 # :350
    ("fennell", 4.0, 0.0, -0.2257317442),                                              # :368
    ("zerodipole", 4.0, 0.0, -0.2257052059), ("zerodipole", 0.0, 2.0, -0.0007522843336),  # :386-387
    ("rf", 0.0, 2.0, -0.00004023807698),                                               # :407
    ("rf_shifted", 4.0, 0.0, -0.1028057400), ("rf_shifted", 0.0, 2.0, -0.00004023807698),  # :420-421
    ("qpotential4", 4.0, 0.0, -0.06896551724), ("qpotential4", 0.0, 2.0, -0.000041002091105),  # :438-439
    ("fanourgakis", 4.0, 0.0, -0.120689655172414),                                     # :462
    ("ewald", 4.0, 0.0, -0.2256758334), ("ewald", 0.0, 2.0, -0.0007522527778),         # :529-530
    ("ewald_screened", 4.0, 0.0, -0.1493013040), ("ewald_screened", 0.0, 2.0, -0.0006704901976),  # :549-550
    ("ewaldt", 4.0, 0.0, -0.2257993685), ("ewaldt", 0.0, 2.0, -0.0007528321650),       # :701-702
    ("poisson33", 4.0, 0.0, -0.137931034482759),                                       # :724
    ("poisson33_yukawa", 4.0, 0.0, -0.03037721287),                                    # :840
    ("ewald_lit", 2.0, 0.0, -0.9027033337),                                            # :579 (charges 1, -1)
    ("ewald_lit", 0.0, 2.0, -0.3851534224),                                            # :648 (two unit dipoles)
]
# neutralization_energy({1,-1,2,3}, volume 2) = chi/2/V * 25
NEUTRALIZATION = [("ewald", -1963.341583564), ("ewald_screened", -1917.674014375), ("ewaldt", -1955.4692924725)]  # :537,:555,:712

# per-pair literals: (scheme, dict(call args), slot, component or None, value[, rtol])
_std = dict(zA=ZA, zB=ZB, muA=MUA, muB=MUB, r=R, quadA=QUADA)
_rq = dict(_std, r=RQ)
_r30 = dict(_std, r=(30.0, 0.0, 0.0))  # (cutoff + 1) * rh
PAIR = [
    # Plain :151-195
    ("plain", _r30, "ion_potential", None, 0.06666666667), ("plain", _std, "ion_potential", None, 0.08695652174),
    ("plain", _r30, "dipole_potential", None, 0.02111111111), ("plain", _std, "dipole_potential", None, 0.03591682420),
    ("plain", _rq, "quadrupole_potential", None, 0.00093632817),
    ("plain", _std, "ion_field", 0, 0.003780718336),
    ("plain", _std, "dipole_field", 0, 0.003123202104), ("plain", _std, "dipole_field", 1, -0.0005753267034),
    ("plain", _std, "dipole_field", 2, -0.0009040848196),
    ("plain", _std, "quadrupole_field", 0, -0.00003752130674), ("plain", _std, "quadrupole_field", 1, -0.00006432224013),
    ("plain", _std, "quadrupole_field", 2, -0.00005360186677),
    ("plain", _r30, "ion_ion_energy", None, 0.2), ("plain", _std, "ion_ion_energy", None, 0.2608695652),
    ("plain", _r30, "ion_dipole_energy", None, -0.02888888889), ("plain", _std, "ion_dipole_energy", None, -0.04914933837),
    ("plain", _r30, "dipole_dipole_energy", None, -0.01185185185), ("plain", _std, "dipole_dipole_energy", None, -0.02630064930),
    ("plain", _rq, "ion_quadrupole_energy", None, 0.002808984511),
    ("plain", _std, "ion_ion_force", 0, 0.01134215501),
    ("plain", _std, "ion_dipole_force", 0, 0.009369606312), ("plain", _std, "ion_dipole_force", 1, -0.001725980110),
    ("plain", _std, "ion_dipole_force", 2, -0.002712254459),
    ("plain", _std, "dipole_dipole_force", 0, 0.003430519474), ("plain", _std, "dipole_dipole_force", 1, -0.004438234569),
    ("plain", _std, "dipole_dipole_force", 2, -0.002551448858),
    # Plain with Debye screening :302-323
    ("plain_yukawa", _r30, "ion_potential", None, 0.01808996296), ("plain_yukawa", _std, "ion_potential", None, 0.03198951663),
    ("plain_yukawa", _r30, "dipole_potential", None, 0.01320042949), ("plain_yukawa", _std, "dipole_potential", None, 0.02642612243),
    ("plain_yukawa", _std, "ion_field", 0, 0.002781697098),
    ("plain_yukawa", _std, "dipole_field", 0, 0.002872404612), ("plain_yukawa", _std, "dipole_field", 1, -0.0004233017324),
    ("plain_yukawa", _std, "dipole_field", 2, -0.0006651884364),
    ("plain_yukawa", _std, "dipole_dipole_force", 0, 0.003594120919), ("plain_yukawa", _std, "dipole_dipole_force", 1, -0.003809715590),
    ("plain_yukawa", _std, "dipole_dipole_force", 2, -0.002190126354),
    # Splined Plain reproduces Plain :979-1006
    ("splined_plain", _std, "ion_potential", None, 0.08695652174), ("splined_plain", _std, "dipole_potential", None, 0.03591682420),
    ("splined_plain", _std, "ion_ion_energy", None, 0.2608695652), ("splined_plain", _std, "ion_dipole_energy", None, -0.04914933837),
    ("splined_plain", _std, "dipole_dipole_energy", None, -0.02630064930), ("splined_plain", _std, "ion_ion_force", 0, 0.01134215501),
    ("splined_plain", _std, "dipole_dipole_force", 1, -0.004438234569),
    # Poisson(4,3) :766-831
    ("poisson43", _std, "ion_potential", None, 0.0009430652121), ("poisson43", _std, "dipole_potential", None, 0.005750206554),
    ("poisson43", _rq, "quadrupole_potential", None, 0.000899228165),
    ("poisson43", _std, "ion_field", 0, 0.0006052849004),
    ("poisson43", _std, "dipole_field", 0, 0.002702513754), ("poisson43", _std, "dipole_field", 1, -0.00009210857180),
    ("poisson43", _std, "dipole_field", 2, -0.0001447420414),
    ("poisson43", _std, "quadrupole_field", 0, 0.00001919309993), ("poisson43", _std, "quadrupole_field", 1, -0.00004053806958),
    ("poisson43", _std, "quadrupole_field", 2, -0.00003378172465),
    ("poisson43", _std, "multipole_field", 0, 0.003326991754330), ("poisson43", _std, "multipole_field", 1, -0.00013264664138),
    ("poisson43", _std, "multipole_field", 2, -0.00017852376605),
    ("poisson43", _std, "ion_ion_energy", None, 0.002829195636), ("poisson43", _std, "ion_dipole_energy", None, -0.007868703705),
    ("poisson43", dict(_std, zA=ZB, muB=MUA, r=neg(R)), "ion_dipole_energy", None, 0.01725061966),     # ion_dipole_energy(zB, muA, -r) :797
    ("poisson43", _std, "dipole_dipole_energy", None, -0.03284312288), ("poisson43", _rq, "ion_quadrupole_energy", None, 0.002697684495),
    ("poisson43", _std, "multipole_multipole_energy", None, -0.020248530406300),
    ("poisson43", _std, "ion_ion_force", 0, 0.001815854701),
    ("poisson43", _std, "ion_dipole_force", 0, 0.008107541263), ("poisson43", _std, "ion_dipole_force", 1, -0.0002763257154),
    ("poisson43", _std, "ion_dipole_force", 2, -0.0004342261242),
    ("poisson43", dict(_std, zB=ZA, muA=MUB, r=neg(R)), "ion_dipole_force", 0, 0.003698176716),         # ion_dipole_force(zA, muB, -r) :817
    ("poisson43", dict(_std, zB=ZA, muA=MUB, r=neg(R)), "ion_dipole_force", 1, -0.0004473844916),
    ("poisson43", dict(_std, zB=ZA, muA=MUB, r=neg(R)), "ion_dipole_force", 2, -0.0001315836740),
    ("poisson43", _std, "dipole_dipole_force", 0, 0.009216400961), ("poisson43", _std, "dipole_dipole_force", 1, -0.002797126801),
    ("poisson43", _std, "dipole_dipole_force", 2, -0.001608010094),
    ("poisson43", _std, "multipole_multipole_force", 0, 0.022895552940790), ("poisson43", _std, "multipole_multipole_force", 1, -0.00364245121674),
    ("poisson43", _std, "multipole_multipole_force", 2, -0.00227516506615),
    # Poisson(3,3) with Debye length 23 :851-913
    ("poisson33_yukawa", _std, "ion_potential", None, 0.003344219306), ("poisson33_yukawa", _std, "dipole_potential", None, 0.01614089171),
    ("poisson33_yukawa", _rq, "quadrupole_potential", None, 0.0016294707475),
    ("poisson33_yukawa", _std, "ion_field", 0, 0.001699041230),
    ("poisson33_yukawa", _std, "dipole_field", 0, 0.004956265485), ("poisson33_yukawa", _std, "dipole_field", 1, -0.0002585497523),
    ("poisson33_yukawa", _std, "dipole_field", 2, -0.0004062924688),
    ("poisson33_yukawa", _std, "quadrupole_field", 0, -0.00005233355205), ("poisson33_yukawa", _std, "quadrupole_field", 1, -0.00007768480608),
    ("poisson33_yukawa", _std, "quadrupole_field", 2, -0.00006473733856),
    ("poisson33_yukawa", _std, "multipole_field", 0, 0.006602973162950), ("poisson33_yukawa", _std, "multipole_field", 1, -0.00033623455838),
    ("poisson33_yukawa", _std, "multipole_field", 2, -0.00047102980736),
    ("poisson33_yukawa", _std, "ion_ion_energy", None, 0.01003265793), ("poisson33_yukawa", _std, "ion_dipole_energy", None, -0.02208753604),
    ("poisson33_yukawa", dict(_std, zA=ZB, muB=MUA, r=neg(R)), "ion_dipole_energy", None, 0.04842267505),
    ("poisson33_yukawa", _std, "dipole_dipole_energy", None, -0.05800464321), ("poisson33_yukawa", _rq, "ion_quadrupole_energy", None, 0.004888412229),
    ("poisson33_yukawa", _std, "multipole_multipole_energy", None, -0.0211832396518),
    ("poisson33_yukawa", _std, "ion_ion_force", 0, 0.005097123689),
    ("poisson33_yukawa", _std, "ion_dipole_force", 0, 0.01486879646), ("poisson33_yukawa", _std, "ion_dipole_force", 1, -0.0007756492577),
    ("poisson33_yukawa", _std, "ion_dipole_force", 2, -0.001218877402),
    ("poisson33_yukawa", dict(_std, zB=ZA, muA=MUB, r=neg(R)), "ion_dipole_force", 0, 0.006782258035),
    ("poisson33_yukawa", dict(_std, zB=ZA, muA=MUB, r=neg(R)), "ion_dipole_force", 1, -0.001255813082),
    ("poisson33_yukawa", dict(_std, zB=ZA, muA=MUB, r=neg(R)), "ion_dipole_force", 2, -0.0003693567885),
    ("poisson33_yukawa", _std, "dipole_dipole_force", 0, 0.002987655323, 1e-6), ("poisson33_yukawa", _std, "dipole_dipole_force", 1, -0.005360251624),
    ("poisson33_yukawa", _std, "dipole_dipole_force", 2, -0.003081497314),
    ("poisson33_yukawa", _std, "multipole_multipole_force", 0, 0.02957883285085), ("poisson33_yukawa", _std, "multipole_multipole_force", 1, -0.00762476838194),
    ("poisson33_yukawa", _std, "multipole_multipole_force", 2, -0.00486394352018),
    # Ewald real-space literature check, charges (1, -1) one unit apart, alpha 0.8, cutoff 5: rAB = positions[0]-positions[1] = (-1,0,0)
    # real energy = ion_ion_energy(1,-1,1) + dipole_dipole_energy(0,0,rAB) :578 ; real force = ion_ion_force(1,-1,rAB)[0] :585
    ("ewald_lit", dict(zA=1.0, zB=-1.0, muA=(0.0, 0.0, 0.0), muB=(0.0, 0.0, 0.0), r=(-1.0, 0.0, 0.0)), "ion_ion_energy", None, -0.2578990353),
    ("ewald_lit", dict(zA=1.0, zB=-1.0, muA=(0.0, 0.0, 0.0), muB=(0.0, 0.0, 0.0), r=(-1.0, 0.0, 0.0)), "ion_ion_force", 0, 0.7338876643),
    # two unit dipoles along x :647, :652: dipole_dipole_energy(mu, mu, rAB) and dipole_field(mu, rAB)[0]
    ("ewald_lit", dict(zA=0.0, zB=0.0, muA=(1.0, 0.0, 0.0), muB=(1.0, 0.0, 0.0), r=(-1.0, 0.0, 0.0)), "dipole_dipole_energy", None, -2.0770407737),
    ("ewald_lit", dict(zA=0.0, zB=0.0, muA=(1.0, 0.0, 0.0), muB=(1.0, 0.0, 0.0), r=(-1.0, 0.0, 0.0)), "dipole_field", 0, 2.0770407737),
]
# values the reference checks to be exactly zero at / beyond the cut-off (:766-822, :850-904)
ZERO_AT_CUTOFF = ["poisson43", "poisson33_yukawa"]


def pair_value(scheme_obj, args, slot, comp):
    out = scheme_obj.pair(args.get("zA", ZA), args.get("zB", ZB), args.get("muA", MUA), args.get("muB", MUB), args["r"],
                          args.get("quadA"), args.get("quadB"))
    v = out[slot]
    return float(v if comp is None else v[comp])
