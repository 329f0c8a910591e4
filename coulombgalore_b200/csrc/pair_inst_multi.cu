// pair_inst_multi.cu -- pair-kernel instantiations of the multi-scheme erfc functor (ion modes): several schemes that share
// eta and the cut-off evaluated in ONE pass over the pairs (cg_eval_multi_device; see detail/srf.hpp: ErfcMultiSrf).
#include "allpairs_kernel.cuh"

using entry_erfc_multi2_functor = cgb::cgd::ErfcMultiSrf<2>;
CGB_DEFINE_FUNCTOR_ENTRY_ION(entry_erfc_multi2, entry_erfc_multi2_functor)
