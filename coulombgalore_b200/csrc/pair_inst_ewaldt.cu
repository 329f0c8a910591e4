// pair_inst_ewaldt.cu -- pair-kernel instantiations for this functor family (one translation unit per family
// so the families compile in parallel; see pair_kernel.cuh).
#include "allpairs_kernel.cuh"

using entry_ewaldt_functor = cgb::cgd::ErfcFamilySrf<cgb::cgd::kEwaldT>;
CGB_DEFINE_FUNCTOR_ENTRY(entry_ewaldt, entry_ewaldt_functor, false)
