// Included by the pbg_k_*.cu translation units only: configuration typedefs and the descriptor factory.
#pragma once
#include "pbg_cfgs.cuh"
#include "pbg_kernels.cuh"

namespace pbg {

// kernel configurations: NB, NJ, FLOATING, NLIM, MAXC, LPE, NCAND, NPAIR, NFEET, NACT, OBS, WARPS per CTA, CTAs per SM
// 14 warps x 2 envs = 28 envs per CTA = one CTA per SM (7.2 KB shared memory per env): 148 CTAs hold 4144 envs.
using CfgPendulum = KCfg<2, 2, 0, 1, 0, 16, 0, 0, 0, 1, 5, 4, 4, 0, 1, TopoDense, 0, 0, 1>;
using CfgDoublePendulum = KCfg<3, 3, 0, 1, 0, 16, 0, 0, 0, 1, 9, 4, 4, 0, 2, TopoDense, 0, 0, 1>;
using CfgDoublePendulumMJ = KCfg<3, 3, 0, 1, 0, 16, 0, 0, 0, 1, 11, 4, 4, 0, 2, TopoDense, 0, 0, 1>;
using CfgReacher = KCfg<4, 4, 0, 3, 0, 16, 0, 0, 0, 2, 9, 4, 4, 0, 4, TopoReacher, 0, 0, 1>;
using CfgHopper = KCfg<6, 6, 0, 3, 6, 16, 8, 0, 1, 3, 15, 14, 1, 0, 3, TopoDense, 0, 0, 2>;
using CfgHopperMJ = KCfg<6, 6, 0, 3, 6, 16, 8, 0, 1, 3, 11, 14, 1, 0, 6, TopoDense, 0, 0, 2>;
using CfgWalkerMJ = KCfg<9, 9, 0, 6, 6, 16, 14, 0, 2, 6, 17, 14, 1, 0, 9, TopoBiped2D, 0, 0, 2>;
using CfgWalker = KCfg<9, 9, 0, 6, 6, 16, 14, 0, 2, 6, 22, 14, 1, 0, 6, TopoBiped2D, 0, 0, 2>;
using CfgCheetah = KCfg<9, 9, 0, 6, 6, 16, 16, 0, 6, 6, 26, 14, 1, 0, 6, TopoBiped2D, 0, 0, 2>;
// HalfCheetahMuJoCoEnv: the cheetah with torsional friction rows (6 rows per contact: up to 42 rows -> one env per warp)
using CfgCheetahMJ = KCfg<9, 9, 0, 6, 6, 32, 16, 0, 6, 6, 17, 14, 1, 0, 9, TopoBiped2D, 0, 1, 2>;
#ifndef PBG_ANT_WARPS
#define PBG_ANT_WARPS 14
#define PBG_ANT_BLOCKS 1
#endif
using CfgAnt = KCfg<9, 8, 1, 8, 6, 16, 25, 0, 4, 8, 28, PBG_ANT_WARPS, PBG_ANT_BLOCKS, 0, 8, TopoAnt, 0, 0, 1>;
using CfgAntMJ = KCfg<9, 8, 1, 8, 6, 16, 25, 0, 4, 8, 111, PBG_ANT_WARPS, PBG_ANT_BLOCKS, 0, 8, TopoAnt, 0, 0, 1>;
using CfgHumanoidMJ = KCfg<18, 17, 1, 17, 12, 32, 30, 66, 2, 17, 376, 14, 1, 0, 17, TopoHumanoid, 36>;
// The humanoid kinds run 14 envs (warps) per SM -- 2048 envs are one wave of 147 CTAs -- which needs <= 16.2 KB of shared memory
// per env: a row budget (17 possible limit rows + 12 x 3 contact rows would be 53; random-policy rollouts peak at 26 rows, the
// robot lying on the ground in FlagrunHarder reaches 42).  The budget is 36: at 40 the env blocks fill the SM and the model tables
// stay in global memory; at 36 (15.1 KB per env) the 10.8 KB of tables fit behind the env blocks as well, +5 % (Humanoid).
using CfgHumanoid = KCfg<18, 17, 1, 17, 12, 32, 30, 66, 2, 17, 44, 14, 1, 0, 17, TopoHumanoid, 36>;
// HumanoidFlagrunHarder: the humanoid + the cube (one more free body, 8 corner candidates, 17 geom-vs-cube pairs)
// FlagrunHarder (the humanoid + the cube: 29 dofs, 8 corner candidates, 17 geom-vs-cube pairs) also runs 14 envs per SM: a 36-row
// budget (a robot lying on the ground reaches 42 rows; P(rows > 36) = 0.24 % of the random-policy env steps, the shallowest
// contacts are dropped then and pbg_episode_stats.contact_overflow counts it) and L / Y rows without the 4-float padding
// (some bank conflicts on the lane-strided row stores) bring the env block to 16.0 KB.
using CfgHarder = KCfg<18, 17, 1, 17, 12, 32, 30, 66, 2, 17, 44, 14, 1, 17, 17, TopoHumanoid, 36, 0, 0, 0>;

template <class C>
static void launch_cfg(const DevModel *m, const StepBuffers &b, const LaunchArgs &la, cudaStream_t s) {
    const int blocks = (la.E + C::EPB - 1) / C::EPB;
    if (la.mode == MODE_POLICY) env_kernel<C, true><<<blocks, C::THREADS, C::SMEM_BYTES, s>>>(m, b, la);
    else env_kernel<C, false><<<blocks, C::THREADS, C::SMEM_BYTES, s>>>(m, b, la);
}
template <class C>
static cudaError_t prepare_cfg() {
    cudaError_t e = cudaFuncSetAttribute(env_kernel<C, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(env_kernel<C, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_BYTES);
}
static void read_phases(unsigned long long *out32, int reset) {
#ifdef PBG_PHASE_CLOCKS
    cudaMemcpyFromSymbol(out32, g_phase, 32 * sizeof(unsigned long long));
    if (reset == 2) { cudaMemcpyFromSymbol(out32, g_warpclk, 4096 * 3 * sizeof(unsigned)); return; }
    if (reset) { unsigned long long z[32] = {0}; cudaMemcpyToSymbol(g_phase, z, sizeof z); }
#else
    for (int i = 0; i < 32; ++i) out32[i] = 0;
    (void)reset;
#endif
}
template <class C>
static KernelInfo info_of() {
    KernelInfo ki = KernelInfo{C::NB, C::NJ, C::FLOATING, C::NLIM, C::MAXC, C::NCAND, C::NPAIR, C::NFEET, C::NACT, C::OBS,
                      C::SSTRIDE, C::CANON, C::EPB, C::THREADS, C::HASX, C::oF, C::NNOISE, C::HIDCAP, C::oT, C::NSLOT, C::MAXR, C::TORS, C::Q0RAW, C::SMEM_BYTES, &launch_cfg<C>, &prepare_cfg<C>, &read_phases, {0}};
    for (int k = 0; k < C::ND; ++k) ki.low[k] = C::low(k);
    return ki;
}


#define PBG_DEFINE_INFO(name) KernelInfo info_##name() { return info_of<Cfg##name>(); }

}  // namespace pbg
