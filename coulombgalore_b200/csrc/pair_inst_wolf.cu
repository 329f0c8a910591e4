// pair_inst_wolf.cu -- pair-kernel instantiations for this functor family (one translation unit per family
// so the families compile in parallel; see pair_kernel.cuh).
#include "allpairs_kernel.cuh"

using entry_wolf_functor = cgb::cgd::ErfcFamilySrf<cgb::cgd::kWolf>;
CGB_DEFINE_FUNCTOR_ENTRY(entry_wolf, entry_wolf_functor, false)
