"""coulombgalore_b200 -- B200-native batched pair-interaction engine behind CoulombGalore's scheme API.

    import coulombgalore_b200 as cg
    pot = cg.Wolf(12.0, 0.25)                       # reference: CoulombGalore::Wolf pot(12.0, 0.25)
    ev = cg.BatchedEvaluator(pot)                   # cg_create: CUDA only, no CPU fallback
    ev.set_system(x, y, z, q, box=[L, L, L], pbc=True)
    out = ev.eval(cg.MODE_ION, cg.FORCE)            # sum_j ion_ion_force(z_j, z_i, r_i - r_j) for every i

The C ABI is include/cg_b200.h; the C++ drop-in headers are under include/coulombgalore/.
"""
from ._capi import (ENERGY, FIELD, FORCE, MODE_DIPOLE, MODE_ION, MODE_MULTIPOLE, MODE_MULTIPOLE_QUAD, POTENTIAL, CgError, lib)  # noqa: F401
from .batched import BatchedEvaluator, eval_multi_begin, eval_multi_device, eval_multi_end  # noqa: F401
from .reciprocal import ReciprocalEwaldGaussian, ReciprocalEwaldState  # noqa: F401
from .schemes import (Ewald, EwaldTruncated, Fanourgakis, Fennell, Plain, Poisson, ReactionField, SchemeBase,  # noqa: F401
                      Splined, Wolf, Zahn, ZeroDipole, createScheme, create_scheme, infinity, qPotential,
                      qPotentialFixedOrder5)


def device_count():
    return int(lib.cg_device_count())


def version():
    return lib.cg_version().decode()
