// pair_inst_poisson.cu -- pair-kernel instantiations for this functor family (one translation unit per family
// so the families compile in parallel; see pair_kernel.cuh).
#include "allpairs_kernel.cuh"

using entry_poisson_functor = cgb::cgd::PoissonSrf;
CGB_DEFINE_FUNCTOR_ENTRY(entry_poisson, entry_poisson_functor, false)
using entry_poisson_yukawa_functor = cgb::cgd::PoissonSrf;
CGB_DEFINE_FUNCTOR_ENTRY(entry_poisson_yukawa, entry_poisson_yukawa_functor, true)
