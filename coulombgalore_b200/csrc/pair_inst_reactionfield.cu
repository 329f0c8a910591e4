// pair_inst_reactionfield.cu -- pair-kernel instantiations for this functor family (one translation unit per family
// so the families compile in parallel; see pair_kernel.cuh).
#include "allpairs_kernel.cuh"

using entry_reactionfield_functor = cgb::cgd::ReactionFieldSrf;
CGB_DEFINE_FUNCTOR_ENTRY(entry_reactionfield, entry_reactionfield_functor, false)
