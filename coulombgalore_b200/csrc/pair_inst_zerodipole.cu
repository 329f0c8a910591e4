// pair_inst_zerodipole.cu -- pair-kernel instantiations for this functor family (one translation unit per family
// so the families compile in parallel; see pair_kernel.cuh).
#include "allpairs_kernel.cuh"

using entry_zerodipole_functor = cgb::cgd::ErfcFamilySrf<cgb::cgd::kZeroDipole>;
CGB_DEFINE_FUNCTOR_ENTRY(entry_zerodipole, entry_zerodipole_functor, false)
