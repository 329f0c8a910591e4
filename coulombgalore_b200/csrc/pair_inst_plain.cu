// pair_inst_plain.cu -- pair-kernel instantiations for this functor family (one translation unit per family
// so the families compile in parallel; see pair_kernel.cuh).
#include "allpairs_kernel.cuh"

using entry_plain_functor = cgb::cgd::PlainSrf;
CGB_DEFINE_FUNCTOR_ENTRY(entry_plain, entry_plain_functor, false)
using entry_plain_yukawa_functor = cgb::cgd::PlainSrf;
CGB_DEFINE_FUNCTOR_ENTRY(entry_plain_yukawa, entry_plain_yukawa_functor, true)
