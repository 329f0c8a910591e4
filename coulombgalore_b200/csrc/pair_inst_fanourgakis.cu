// pair_inst_fanourgakis.cu -- pair-kernel instantiations for this functor family (one translation unit per family
// so the families compile in parallel; see pair_kernel.cuh).
#include "allpairs_kernel.cuh"

using entry_fanourgakis_functor = cgb::cgd::FanourgakisSrf;
CGB_DEFINE_FUNCTOR_ENTRY(entry_fanourgakis, entry_fanourgakis_functor, false)
