// pair_inst_spline.cu -- pair-kernel instantiations for this functor family (one translation unit per family
// so the families compile in parallel; see pair_kernel.cuh).
#include "allpairs_kernel.cuh"

using entry_spline_functor = cgb::cgd::SplineSrf;
CGB_DEFINE_FUNCTOR_ENTRY(entry_spline, entry_spline_functor, false)
using entry_spline_yukawa_functor = cgb::cgd::SplineSrf;
CGB_DEFINE_FUNCTOR_ENTRY(entry_spline_yukawa, entry_spline_yukawa_functor, true)
