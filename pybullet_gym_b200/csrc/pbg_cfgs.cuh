// Kernel configurations of libpbg_b200 and the per-configuration launch descriptors.  Every configuration is instantiated in
// its own translation unit (pbg_k_*.cu), so that the library builds in parallel; pbg_abi.cu only sees the descriptors.
#pragma once
#include "pbg_model.cuh"
#include <cuda_runtime.h>

namespace pbg {

struct KernelInfo {
    int nb, nj, floating, nlim, maxc, ncand, npair, nfeet, nact, obs, sstride, canon, epb, threads, hasx, off_feet, nnoise, ysz, off_task, nslot, maxr, tors, q0id;
    size_t smem;
    void (*launch)(const DevModel *, const StepBuffers &, const LaunchArgs &, cudaStream_t);
    cudaError_t (*prepare)();
    void (*phases)(unsigned long long *out32, int reset);   // development builds (-DPBG_PHASE_CLOCKS): per-phase cycle sums
    unsigned low[32];         // the kernel's compile-time tree topology: dofs below k coupled with dof k (KCfg::low)
};

#define PBG_FOR_EACH_CFG(X) \
    X(Pendulum) X(DoublePendulum) X(DoublePendulumMJ) X(Reacher) X(Hopper) X(HopperMJ) X(WalkerMJ) X(Walker) X(Cheetah) \
    X(Ant) X(AntMJ) X(HumanoidMJ) X(Humanoid) X(Harder) X(CheetahMJ)
#define PBG_DECL_INFO(name) KernelInfo info_##name();
PBG_FOR_EACH_CFG(PBG_DECL_INFO)
#undef PBG_DECL_INFO

}  // namespace pbg
