// Kernel instantiations: Walker, WalkerMJ (see pbg_kcfg.cuh).
#include "pbg_kcfg.cuh"
namespace pbg {
PBG_DEFINE_INFO(Walker)
PBG_DEFINE_INFO(WalkerMJ)
}
