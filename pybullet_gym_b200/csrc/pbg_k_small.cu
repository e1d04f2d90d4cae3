// Kernel instantiations: Pendulum, DoublePendulum, DoublePendulumMJ, Reacher (see pbg_kcfg.cuh).
#include "pbg_kcfg.cuh"
namespace pbg {
PBG_DEFINE_INFO(Pendulum)
PBG_DEFINE_INFO(DoublePendulum)
PBG_DEFINE_INFO(DoublePendulumMJ)
PBG_DEFINE_INFO(Reacher)
}
