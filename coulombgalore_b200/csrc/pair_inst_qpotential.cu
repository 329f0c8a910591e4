// pair_inst_qpotential.cu -- pair-kernel instantiations for this functor family (one translation unit per family
// so the families compile in parallel; see pair_kernel.cuh).
#include "allpairs_kernel.cuh"

using entry_qpotential_functor = cgb::cgd::QPotentialSrf<0>;
CGB_DEFINE_FUNCTOR_ENTRY(entry_qpotential, entry_qpotential_functor, false)
using entry_qpotential3_functor = cgb::cgd::QPotentialSrf<3>;
CGB_DEFINE_FUNCTOR_ENTRY(entry_qpotential3, entry_qpotential3_functor, false)
