// pair_inst_ewald.cu -- pair-kernel instantiations for this functor family (one translation unit per family
// so the families compile in parallel; see pair_kernel.cuh).
#include "allpairs_kernel.cuh"

using entry_ewald_functor = cgb::cgd::ErfcFamilySrf<cgb::cgd::kEwaldPlain>;
CGB_DEFINE_FUNCTOR_ENTRY(entry_ewald, entry_ewald_functor, false)
using entry_ewald_screened_functor = cgb::cgd::EwaldScreenedSrf;
CGB_DEFINE_FUNCTOR_ENTRY(entry_ewald_screened, entry_ewald_screened_functor, false)
