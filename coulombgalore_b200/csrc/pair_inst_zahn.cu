// pair_inst_zahn.cu -- pair-kernel instantiations for this functor family (one translation unit per family
// so the families compile in parallel; see pair_kernel.cuh).
#include "allpairs_kernel.cuh"

using entry_zahn_functor = cgb::cgd::ErfcFamilySrf<cgb::cgd::kZahn>;
CGB_DEFINE_FUNCTOR_ENTRY(entry_zahn, entry_zahn_functor, false)
