// Kernel instantiations: CheetahMJ (see pbg_kcfg.cuh).
#include "pbg_kcfg.cuh"
namespace pbg {
PBG_DEFINE_INFO(CheetahMJ)
}
