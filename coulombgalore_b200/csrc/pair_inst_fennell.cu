// pair_inst_fennell.cu -- pair-kernel instantiations for this functor family (one translation unit per family
// so the families compile in parallel; see pair_kernel.cuh).
#include "allpairs_kernel.cuh"

using entry_fennell_functor = cgb::cgd::ErfcFamilySrf<cgb::cgd::kFennell>;
CGB_DEFINE_FUNCTOR_ENTRY(entry_fennell, entry_fennell_functor, false)
