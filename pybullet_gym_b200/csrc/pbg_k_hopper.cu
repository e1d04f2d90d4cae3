// Kernel instantiations: Hopper, HopperMJ (see pbg_kcfg.cuh).
#include "pbg_kcfg.cuh"
namespace pbg {
PBG_DEFINE_INFO(Hopper)
PBG_DEFINE_INFO(HopperMJ)
}
