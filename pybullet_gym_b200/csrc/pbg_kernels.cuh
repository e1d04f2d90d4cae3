// Fused env-step kernel of libpbg_b200 (sm_100a).
//
// One environment per LPE-lane group (LPE = 16: two envs per warp, LPE = 32: one env per warp).
// Replaces, per env step, the reference's apply_action -> stepSimulation -> calc_state -> reward
// chain (/root/reference/pybulletgym/envs/roboschool/gym_locomotion_envs.py:54-114 and the 23..72
// pybullet calls behind it) with one launch:
//
//   load state (coalesced) -> smem
//   frame_skip x { FK (lane = body, level by level) -> contact generation (lane = candidate)
//                  -> composite inertias + bias wrench (lane = body, subtree sums lane = component)
//                  -> joint-space inertia rows + Cholesky in registers (lane = dof)
//                  -> constraint rows J, Y = L^-1 J^T (lane = row) -> Delassus matrix A = Y Y^T
//                  -> projected Gauss-Seidel on lambda (lane = row, one shuffle per row update)
//                  -> du = L^-T (h y_f + Y^T lambda), semi-implicit integration }
//   FK -> calc_state / reward / termination (+ in-kernel reset) -> store
//
// No tensor cores: per-env matrices are 6..23 wide.
#pragma once
#include "pbg_model.cuh"
#include <math_constants.h>

namespace pbg {

// Development builds (-DPBG_PHASE_CLOCKS): per-phase cycle counters summed over warps (lane 0 of each warp adds its clock64
// deltas) and the duration of every warp of the last launch; read through pbg_debug_phases (tools/ab.py, tools/warpclk.py).
#ifdef PBG_PHASE_CLOCKS
__device__ unsigned long long g_phase[32];
__device__ unsigned g_warpclk[4096 * 3];      // per warp of the last launch: cycles, smid, most rows beyond LPE
#define PBG_PHASE(id) do { const long long _c = clock64(); if ((threadIdx.x & 31) == 0) atomicAdd(&g_phase[id], (unsigned long long)(_c - _pc)); _pc = _c; } while (0)
#define PBG_PHASE_BEGIN long long _pc = clock64()
#define PBG_PHASE_RESET _pc = clock64()
#else
#define PBG_PHASE(id) do { } while (0)
#define PBG_PHASE_BEGIN do { } while (0)
#define PBG_PHASE_RESET do { } while (0)
#endif

// Compile-time tree topology of a robot: low(k) = bit mask of the dofs below k that dof k is coupled with in the joint-space
// inertia matrix (its ancestors in the kinematic tree; the six dofs of a floating base count as a chain).  M = L^T L is
// factorised leaves first (Featherstone's branch-induced sparsity), so L[k][i] is structurally zero unless i is in low(k):
// the unrolled factorisation / substitution loops skip those entries at compile time.  pbg_create checks the masks against
// the compiled model (DevModel.up) and refuses a model that does not match.
struct TopoDense {      // serial chains (pendula, Hopper) and the fallback: everything below k
    static constexpr __host__ __device__ unsigned low(int k) { return (1u << k) - 1u; }
};
struct TopoReacher {    // arm (joint0 -> joint1) and the target's two slides, unconnected
    static constexpr __host__ __device__ unsigned low(int k) { return k == 1 ? 0x1u : (k == 3 ? 0x4u : 0u); }
};
struct TopoBiped2D {    // Walker2D / HalfCheetah: root chain 0-2, first leg 3-5, second leg 6-8 (hangs off the root chain)
    static constexpr __host__ __device__ unsigned low(int k) {
        return k < 6 ? (1u << k) - 1u : (k == 6 ? 0x7u : (k == 7 ? 0x47u : 0xc7u));
    }
};
struct TopoAnt {        // base 0-5, four legs of (hip, ankle)
    static constexpr __host__ __device__ unsigned low(int k) {
        return k < 6 ? (1u << k) - 1u : (((k - 6) & 1) ? (0x3fu | (1u << (k - 1))) : 0x3fu);
    }
};
struct TopoHumanoid {   // base 0-5, abdomen 6-8, right leg 9-12, left leg 13-16 (both off the abdomen), right arm 17-19, left arm 20-22
    static constexpr __host__ __device__ unsigned low(int k) {
        return k < 13 ? (1u << k) - 1u
             : k < 17 ? (0x1ffu | (((1u << (k - 13)) - 1u) << 13))
             : k < 20 ? (0x3fu | (((1u << (k - 17)) - 1u) << 17))
                      : (0x3fu | (((1u << (k - 20)) - 1u) << 20));
    }
};

// XP_ > 0: the world also holds HumanoidFlagrunHarder's cube (rs/robot_locomotors.py:236-266), a second free
// body that is the last body / the last six dofs / the last eight ground candidates / the last XP_ pairs.
template <int NB_, int NJ_, int FLOATING_, int NLIM_, int MAXC_, int LPE_, int NCAND_, int NPAIR_, int NFEET_,
          int NACT_, int OBS_, int WARPS_, int MIN_BLOCKS_, int XP_ = 0, int NNOISE_ = NACT_, class TOPO_ = TopoDense, int MAXROWS_ = 0, int TORS_ = 0, int Q0ID_ = 0, int LSTPAD_ = 4>
struct KCfg {
    static constexpr int NNOISE = NNOISE_;                     // injectable reset draws per env (pbg_reset_with)
    static constexpr int HASX = XP_ > 0 ? 1 : 0;
    static constexpr int NB = NB_ + HASX, NJ = NJ_, FLOATING = FLOATING_, NLIM = NLIM_;
    static constexpr int NB_ROBOT = NB_;
    static constexpr int XD0 = NJ_ + 6 * FLOATING_;            // first cube dof
    static constexpr int ND = XD0 + 6 * HASX;
    // structural coupling masks (see Topo* above); the cube's six dofs are a chain of their own
    static constexpr __host__ __device__ unsigned low(int k) {
        return (HASX && k >= XD0) ? (((1u << (k - XD0)) - 1u) << XD0) : TOPO_::low(k);
    }
    static constexpr __host__ __device__ bool coupled(int hi, int lo) { return ((low(hi) >> lo) & 1u) != 0u; }
    static constexpr int MAXC = MAXC_, LPE = LPE_, NCAND = NCAND_ + 8 * HASX, NPAIR = NPAIR_ + XP_, NSLOT = NCAND + NPAIR;
    static constexpr int NFEET = NFEET_, NACT = NACT_, OBS = OBS_;
    // observations wider than the 64-float staging area are MuJoCo-style layouts whose tail is zero padding
    // (pybulletgym/envs/mujoco/robot_locomotors.py:222-319): only the first OBSNZ entries are staged
    static constexpr int OBSNZ = OBS_ <= 64 ? OBS_ : 11 + 2 * NJ_;
    // constraint rows per env.  MAXROWS_ > 0 is a row budget below the worst case NLIM + 3 MAXC (it buys shared memory: the
    // Delassus matrix is MAXR^2): the violated joint limits of a sub-step come first, contacts get (MAXR - nl) / RPC of the rest
    // TORS_: every contact also gets Bullet's torsional friction rows (one spinning about the normal, two rolling about the
    // tangents), placed between the normal rows and the lateral friction rows: RPC rows per contact
    static constexpr int TORS = TORS_ ? 1 : 0;
    // Q0ID_: every body's rest rotation relative to its parent is the identity (true for all the MJCF robots but the Humanoid):
    // forward kinematics skips that matrix product and the lanes do not carry the nine constants (pbg_create checks the model)
    static constexpr int Q0ID = Q0ID_ ? 1 : 0;
    // Q0ID_ == 2: the robot is also PLANAR -- fixed base, every hinge about +-y with a unit axis, every slide axis a unit vector in
    // the xz plane (Hopper, Walker2D, HalfCheetah; pbg_create checks the model).  Orientations are then angles that add along the
    // tree, angular velocities stay parallel to y, and forward kinematics becomes one lane-parallel pass plus two sweeps of
    // parent-to-child additions (fk_planar) instead of a full 3-D frame composition per tree level.
    static constexpr int PLANAR = Q0ID_ == 2 ? 1 : 0;
    static constexpr int Q0RAW = Q0ID_;
    static_assert(!PLANAR || (FLOATING_ == 0 && XP_ == 0), "planar kinds have a fixed base and no cube");
    static constexpr bool FKREG = true;       // link-frame constants of forward kinematics stay in registers (read at every tree level)
    // kinds that are not register-bound (7 warps per SM: 255 registers per thread) also keep a body's kinematics of the pass in
    // registers between forward kinematics and the wrench phase instead of re-reading its kin() record
    static constexpr bool KINREG = WARPS_ <= 8;
    static constexpr int RPC = 3 + 3 * TORS;
    static constexpr int MAXR = MAXROWS_ > 0 ? MAXROWS_ : NLIM + RPC * MAXC;
    static_assert(MAXR <= NLIM + RPC * MAXC && MAXR >= NLIM, "row budget");
    static_assert(!TORS || NPAIR_ + XP_ == 0, "torsional rows are implemented for ground contacts only");
    static constexpr int MAXRP = MAXR > 0 ? MAXR : 1;
    static constexpr int NDP = (ND + 3) / 4 * 4;
    static constexpr int LST = NDP + LSTPAD_;      // row stride of L / Y: float4 rows; + 4 makes lane-strided row accesses conflict-free
    static constexpr int EPW = 32 / LPE;
    // One CTA per SM: all warps of an SM run the same phase of the same substep at about the same
    // time (block barrier per substep), so the long straight-line phases are fetched once per SM
    // instead of once per warp -- instruction fetch was the top stall with small independent CTAs.
    static constexpr int WARPS = WARPS_, THREADS = 32 * WARPS_, EPB = WARPS * EPW;
    static constexpr int MIN_BLOCKS = MIN_BLOCKS_;
    static_assert(MAXR <= 2 * LPE, "at most two row slots per lane");
    static_assert(ND + 1 <= LPE && NB <= LPE && NCAND <= 2 * LPE && NLIM <= LPE, "lane budget");
    // per-env state (floats)
    static constexpr int oQ = 7 * FLOATING, oX = oQ + NJ, oU = oX + 7 * HASX, oW = oU + ND, oT = oW + NSLOT, oF = oT + TASK_FLOATS;
    static constexpr int oP = oF + NFEET;          // feet flags of the last physics step, not yet seen by calc_state
    static constexpr int SSIZE = oP + NFEET, SSTRIDE = (SSIZE + 3) / 4 * 4;
    static constexpr int CANON = 13 * FLOATING + 2 * NJ + 13 * HASX;   // [base 13][q][qd][cube pos3 quat4 omega3 vel3]
    // shared memory per env (floats).  Regions with disjoint lifetimes share storage:
    //   {KIN, ACC, SH, F, COL} (kinematics .. Cholesky, dead once the rows are built)  |  {A} (Delassus matrix)
    //   {Y} (constraint rows, substeps only)                                          |  {OUT} (staged outputs)
    static constexpr int KS = 31;                  // R9 x3 w3 v3 al3 a3 z3 A3 (+1 pad)
    static constexpr int CTS = 16;                 // contact record: bodyA bodyB slot pad pA3 pB3 n3 dist mu pad
    static constexpr int sST = 0;
    static constexpr int sL = (sST + SSTRIDE + 3) / 4 * 4;
    static constexpr int sINV = sL + ND * LST;
    static constexpr int sY = (sINV + NDP + 3) / 4 * 4;
    static constexpr int YSZ = MAXRP * LST > 72 ? MAXRP * LST : 72;
    static constexpr int sOUT = sY;                // staged outputs: obs[64] reward terms[5]
    static constexpr int sLAM = sY + YSZ;
    static constexpr int sCT = sLAM + MAXRP;
    static constexpr int sLIM = sCT + (MAXC > 0 ? MAXC : 1) * CTS;
    static constexpr int sCD = sLIM + 2 * (NLIM > 0 ? NLIM : 1);
    static constexpr int sMISC = sCD + (NSLOT > 0 ? NSLOT : 1);   // this step's feet flags
    static constexpr int sACTN = sMISC + 8;                        // this step's actions, read from global / mapped host memory once
    static constexpr int sU = (sACTN + NACT + 3) / 4 * 4;
    static constexpr int sKIN = sU;
    static constexpr int sACC = sKIN + NB * KS;
    static constexpr int sSH = (sACC + NB * 17 + 3) / 4 * 4;
    static constexpr int sF = sSH + ND * 12;
    static constexpr int sCOL = (sF + NDP + 3) / 4 * 4;
    static constexpr int G1 = sCOL + 2 * (NDP + 4) - sU;
    static constexpr int sA = sU;
    static constexpr int ENV_FLOATS_RAW = (sU + (G1 > MAXRP * MAXRP + 2 * LPE ? G1 : MAXRP * MAXRP + 2 * LPE) + 3) / 4 * 4;
    // two envs per warp: place the second env's block 16 banks away from the first one's (stride = 16 mod 32), so that the two
    // half-warps never collide, neither on env-uniform (broadcast) nor on lane-indexed shared-memory accesses
    static constexpr int ENV_FLOATS = LPE == 16 ? ENV_FLOATS_RAW + ((48 - ENV_FLOATS_RAW % 32) % 32) : ENV_FLOATS_RAW;
    static constexpr size_t ENV_BYTES = size_t(ENV_FLOATS) * EPB * sizeof(float);
    // the model tables (13 KB of joint / geometry / scene constants read all over the step) are copied into shared memory behind
    // the env blocks when they fit: lane-indexed table reads become shared-memory loads instead of L1-cached global loads
    static constexpr bool MODEL_IN_SMEM = ENV_BYTES + sizeof(DevModel) <= 227 * 1024;
    static constexpr size_t SMEM_BYTES = ENV_BYTES + (MODEL_IN_SMEM ? sizeof(DevModel) : 0);
    static constexpr int HIDCAP = ENV_FLOATS - sU;             // room for the fused policy's hidden activations
};

// ---------------------------------------------------------------------------------------------
struct V3 {
    float x, y, z;
};
__device__ __forceinline__ V3 mk(float x, float y, float z) { return V3{x, y, z}; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 operator*(float s, V3 a) { return mk(s * a.x, s * a.y, s * a.z); }
__device__ __forceinline__ float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ V3 cross(V3 a, V3 b) {
    return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ V3 ld3(const float *p) { return mk(p[0], p[1], p[2]); }
__device__ __forceinline__ void st3(float *p, V3 a) { p[0] = a.x; p[1] = a.y; p[2] = a.z; }
__device__ __forceinline__ V3 mulR(const float *R, V3 a) {   // row-major 3x3 times vector
    return mk(R[0] * a.x + R[1] * a.y + R[2] * a.z, R[3] * a.x + R[4] * a.y + R[5] * a.z,
              R[6] * a.x + R[7] * a.y + R[8] * a.z);
}
__device__ __forceinline__ V3 mulRt(const float *R, V3 a) {
    return mk(R[0] * a.x + R[3] * a.y + R[6] * a.z, R[1] * a.x + R[4] * a.y + R[7] * a.z,
              R[2] * a.x + R[5] * a.y + R[8] * a.z);
}
__device__ __forceinline__ void quat2mat(const float *q, float *R) {
    float x = q[0], y = q[1], z = q[2], w = q[3];
    R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - w * z); R[2] = 2 * (x * z + w * y);
    R[3] = 2 * (x * y + w * z); R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - w * x);
    R[6] = 2 * (x * z - w * y); R[7] = 2 * (y * z + w * x); R[8] = 1 - 2 * (x * x + y * y);
}
__device__ __forceinline__ void mat2quat(const float *m, float *q) {
    float t = m[0] + m[4] + m[8];
    if (t > 0) {
        float s = sqrtf(t + 1.0f) * 2;
        q[3] = 0.25f * s; q[0] = (m[7] - m[5]) / s; q[1] = (m[2] - m[6]) / s; q[2] = (m[3] - m[1]) / s;
    } else if (m[0] >= m[4] && m[0] >= m[8]) {
        float s = sqrtf(m[0] - m[4] - m[8] + 1.0f) * 2;
        q[0] = 0.25f * s; q[1] = (m[3] + m[1]) / s; q[2] = (m[6] + m[2]) / s; q[3] = (m[7] - m[5]) / s;
    } else if (m[4] >= m[8]) {
        float s = sqrtf(m[4] - m[8] - m[0] + 1.0f) * 2;
        q[1] = 0.25f * s; q[2] = (m[7] + m[5]) / s; q[0] = (m[1] + m[3]) / s; q[3] = (m[2] - m[6]) / s;
    } else {
        float s = sqrtf(m[8] - m[0] - m[4] + 1.0f) * 2;
        q[2] = 0.25f * s; q[0] = (m[2] + m[6]) / s; q[1] = (m[5] + m[7]) / s; q[3] = (m[3] - m[1]) / s;
    }
}

// Philox4x32-10, bit-identical to oracle/oracle.c rng_uniform_s
// The task layer runs once per env step, by which time the sub-step loop has pushed its code out of the 32 KB instruction cache: its
// cost is instruction fetch (half of its stall samples are no_inst).  The libdevice routines it calls several times are kept
// out of line so that each is fetched once and re-executed from the cache.
static __device__ __noinline__ float atan2_shared(float y, float x) { return atan2f(y, x); }
static __device__ __noinline__ float2 sincos_shared(float x) {
    float s, c;
    sincosf(x, &s, &c);
    return make_float2(s, c);
}

__device__ __forceinline__ float rng_uniform(unsigned long long seed, unsigned long long env, unsigned ep,
                                             unsigned stream, unsigned n, float lo, float hi) {
    unsigned c0 = (unsigned)env, c1 = (unsigned)(env >> 32), c2 = ep, c3 = (stream << 24) | (n >> 2);
    unsigned k0 = (unsigned)seed, k1 = (unsigned)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        unsigned hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        unsigned hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        unsigned n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    unsigned r = (n & 3) == 0 ? c0 : ((n & 3) == 1 ? c1 : ((n & 3) == 2 ? c2 : c3));
    float u = (float)(r >> 8) * (1.0f / 16777216.0f);
    return fmaf(hi - lo, u, lo);
}

// ---------------------------------------------------------------------------------------------
template <class C>
struct Env {
    const DevModel *__restrict__ m;
    float *sm;          // this env's shared-memory block
    int gl;             // lane within the group
    int grp;            // group within the warp
    // lane-as-body constants.  Only the tree indices stay in registers; link frames, masses and inertias are read from the
    // model tables (shared memory for most kinds) where they are used -- once per sub-step each -- and a body's kinematics of
    // the current pass live in its kin() record: the kernel is register-bound (128 per thread at 14 warps per SM)
    int bparent, bjtype, bdepth, bdof;
    float anchor_p[3], com_off[3], axis[3], alen;     // link-frame constants of forward kinematics
    float kR[9]; V3 kx, kw, kv, kal, ka;               // KINREG kinds: this body's kinematics of the current pass
    float tau;          // joint force of this dof for the current env step
    int nc, nl;         // active contacts / limit rows of this env (group-uniform)
    int dbg_nov;        // development: most rows beyond LPE any sub-step of this env step had (warp-wide)
    int ovf;            // this env step had a sub-step with more than MAXC candidates in contact (the solver kept the deepest)
    float *dbg;         // optional debug dump of the constraint rows (development only)
    unsigned long long rng_seed, rng_env;   // counter-RNG key / stream of this env

    static constexpr unsigned FULL = 0xffffffffu;

    __device__ __forceinline__ float shfl(float v, int src) const { return __shfl_sync(FULL, v, src, C::LPE); }
    __device__ __forceinline__ int shfli(int v, int src) const { return __shfl_sync(FULL, v, src, C::LPE); }
    __device__ __forceinline__ float gsum(float v) const {
#pragma unroll
        for (int o = C::LPE / 2; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o, C::LPE);
        return v;
    }
    __device__ __forceinline__ unsigned gballot(bool p) const {
        unsigned b = __ballot_sync(FULL, p);
        if (C::LPE == 32) return b;
        return (b >> (grp * C::LPE)) & ((1u << (C::LPE & 31)) - 1u);
    }
    __device__ __forceinline__ int wmax(int v) const { return __reduce_max_sync(FULL, v); }

    __device__ __forceinline__ float *kin(int b) const { return sm + C::sKIN + b * C::KS; }
    __device__ __forceinline__ float *st() const { return sm + C::sST; }
    __device__ __forceinline__ float *uvec() const { return sm + C::sST + C::oU; }

    __device__ void load_lane_constants() {
        const int b = gl < C::NB ? gl : 0;
        bparent = m->parent[b]; bjtype = m->jtype[b]; bdepth = gl < C::NB ? m->depth[b] : 1000; bdof = m->dof[b];
        if (C::FKREG) {
#pragma unroll
            for (int i = 0; i < 3; ++i) { anchor_p[i] = m->anchor_p[b][i]; com_off[i] = m->com_off[b][i]; axis[i] = m->axis[b][i]; }
            alen = m->axis_len[b];
        }
        tau = 0.f; nc = 0; nl = 0; ovf = 0; dbg = nullptr;
    }

    // ---------------------------------------------------------------- forward kinematics
    // lane = body.  bias=true also propagates the velocity-product accelerations needed for the bias wrench.
    // Planar robots (C::PLANAR): body b's orientation is Ry(theta_b) with theta_b the signed sum of the hinge angles above it, its
    // angular velocity (0, Omega_b, 0) likewise, so every lane first gets (theta, Omega) by one sweep of parent-to-child additions
    // (two shuffles per tree level), then computes ITS segment -- parent COM to anchor to own COM -- and that segment's
    // contributions to position, velocity and velocity-product acceleration on its own, and a second sweep adds the parents'
    // totals (nine shuffles per level).  Same quantities, in the same kin() records, as the general routine below: with all
    // axes parallel the angular velocity-product acceleration vanishes and w x (w x r) = -Omega^2 r_perp.
    __device__ void fk_planar(bool bias) {
        const float *S = st();
        const int maxdepth = m->maxdepth;
        const bool act = gl < C::NB;
        const bool hinge = act && bjtype == 1;
        const float q = act ? S[C::oQ + bdof] : 0.f, qd = act ? S[C::oU + bdof] : 0.f;
        const V3 ax = ld3(axis), cOff = ld3(com_off), aP = ld3(anchor_p);
        const int psrc = bparent < 0 ? 0 : bparent;
        float th = hinge ? ax.y * q : 0.f, om = hinge ? ax.y * qd : 0.f, omp = 0.f;
        for (int lvl = 1; lvl <= maxdepth; ++lvl) {
            const float t = shfl(th, psrc), o = shfl(om, psrc);
            if (bdepth == lvl) { th += t; om += o; omp = o; }
        }
        float s, c;
        sincosf(th, &s, &c);
        float sp = shfl(s, psrc), cp = shfl(c, psrc);
        if (bparent < 0) { sp = 0.f; cp = 1.f; }
        // Ry(theta) v = (c vx + s vz, vy, -s vx + c vz)
        const V3 e = mk(cp * aP.x + sp * aP.z, aP.y, -sp * aP.x + cp * aP.z);              // parent COM -> joint anchor
        V3 zw, f, dx, dv, da = mk(0.f, 0.f, 0.f);
        if (hinge) {
            zw = mk(0.f, ax.y, 0.f);
            f = mk(c * cOff.x + s * cOff.z, cOff.y, -s * cOff.x + c * cOff.z);            // anchor -> own COM
            dx = e + f;
            dv = mk(omp * e.z + om * f.z, 0.f, -omp * e.x - om * f.x);                     // w_p x e + w x f
            if (bias) da = mk(-omp * omp * e.x - om * om * f.x, 0.f, -omp * omp * e.z - om * om * f.z);
        } else {
            zw = mk(cp * ax.x + sp * ax.z, ax.y, -sp * ax.x + cp * ax.z);                  // slide: the link keeps its parent's orientation
            f = mk(cp * cOff.x + sp * cOff.z, cOff.y, -sp * cOff.x + cp * cOff.z);
            dx = e + q * zw + f;
            dv = mk(omp * dx.z, 0.f, -omp * dx.x) + qd * zw;                                // w_p x (x - x_p) + qd z
            if (bias) da = mk(-omp * omp * dx.x, 0.f, -omp * omp * dx.z) + (2.f * qd * omp) * mk(zw.z, 0.f, -zw.x);
        }
        V3 x = dx, v = dv, a = da;
        for (int lvl = 1; lvl <= maxdepth; ++lvl) {
            const V3 px = mk(shfl(x.x, psrc), shfl(x.y, psrc), shfl(x.z, psrc));
            const V3 pv = mk(shfl(v.x, psrc), shfl(v.y, psrc), shfl(v.z, psrc));
            if (bdepth == lvl) { x = x + px; v = v + pv; }
            if (bias) {
                const V3 pa = mk(shfl(a.x, psrc), shfl(a.y, psrc), shfl(a.z, psrc));
                if (bdepth == lvl) a = a + pa;
            }
        }
        if (act) {
            const V3 A = hinge ? x - f : x - f - q * zw;
            const V3 w = mk(0.f, om, 0.f);
            float *k = kin(gl);
            k[0] = c; k[1] = 0.f; k[2] = s; k[3] = 0.f; k[4] = 1.f; k[5] = 0.f; k[6] = -s; k[7] = 0.f; k[8] = c;
            st3(k + 9, x); st3(k + 12, w); st3(k + 15, v);
            if (bias) { st3(k + 18, mk(0.f, 0.f, 0.f)); st3(k + 21, a); }
            st3(k + 24, zw); st3(k + 27, A);
            k[30] = th;                          // the body's accumulated angle: the task layer's pitch without an atan2f
            if (C::KINREG) {
                kR[0] = c; kR[1] = 0.f; kR[2] = s; kR[3] = 0.f; kR[4] = 1.f; kR[5] = 0.f; kR[6] = -s; kR[7] = 0.f; kR[8] = c;
                kx = x; kw = w; kv = v; kal = mk(0.f, 0.f, 0.f); ka = a;
            }
        }
        __syncwarp();
    }

    // General trees: the same idea as fk_planar in 3-D.  Only the orientation is a genuinely multiplicative recursion,
    // R_b = R_p (Q0_b Rj(q_b)); everything else -- position, angular velocity, linear velocity, and the two velocity-product
    // accelerations -- is a SUM along the chain of per-body increments that each lane can compute on its own once the previous
    // quantity is known for itself and its parent:
    //     dx = e + f (+ q z)                      e = R_p anchor, f = R_b com   (parent COM -> anchor -> own COM)
    //     w  = w_p + qd z                         z = |axis| R_b axis
    //     dv = w_p x e + w x f                    (slide: w_p x dx + qd z)
    //     al = al_p + qd (w_p x z)
    //     da = al_p x e + w_p x (w_p x e) + al x f + w x (w x f)        (slide: al_p x dx + w_p x (w_p x dx) + 2 qd (w_p x z))
    // So: every lane builds its local rotation at once (sincos, Rodrigues, rest rotation), one sweep over the tree levels
    // multiplies the rotations down (9 shuffles + 27 FMAs per level), and three more sweeps of pure parent-to-child additions
    // (6, 6 and 3 shuffles per level) carry (x, w), (v, al) and a, with the cross products done by all lanes in parallel in
    // between.  The level-by-level routine this replaces ran the whole per-body computation once per tree level with the one to
    // four lanes of that level active: 27 % of a Humanoid step (six levels).
    __device__ void fk_tree(bool bias) {
        const float *S = st();
        const int maxdepth = m->maxdepth;
        const bool act = gl < C::NB;
        const int b = act ? gl : 0;
        const bool is_cube = C::HASX && act && bjtype == 4;
        const bool rootfree = act && (bjtype == 3 || is_cube);
        const bool hinge = act && bjtype == 1;
        const int psrc = bparent < 0 ? 0 : bparent;
        float L[9] = {1.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 1.f};
        float q = 0.f, qd = 0.f;
        V3 X = mk(0.f, 0.f, 0.f), W = X, V = X;
        const V3 ax = ld3(axis), cOff = ld3(com_off), aP = ld3(anchor_p);
        if (rootfree) {
            const float *P = is_cube ? S + C::oX : S;
            const float *U = S + C::oU + (is_cube ? C::XD0 : 0);
            quat2mat(P + 3, L);
            X = ld3(P); W = ld3(U); V = ld3(U + 3);
        } else if (act) {
            q = S[C::oQ + bdof - 6 * C::FLOATING]; qd = S[C::oU + bdof];
            float Rj[9] = {1.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 1.f};
            if (hinge) {
                float sn, cs;
                sincosf(q, &sn, &cs);
                const float t = 1.f - cs;
                Rj[0] = cs + t * ax.x * ax.x; Rj[1] = t * ax.x * ax.y - sn * ax.z; Rj[2] = t * ax.x * ax.z + sn * ax.y;
                Rj[3] = t * ax.x * ax.y + sn * ax.z; Rj[4] = cs + t * ax.y * ax.y; Rj[5] = t * ax.y * ax.z - sn * ax.x;
                Rj[6] = t * ax.x * ax.z - sn * ax.y; Rj[7] = t * ax.y * ax.z + sn * ax.x; Rj[8] = cs + t * ax.z * ax.z;
            }
            if (C::Q0ID) {
#pragma unroll
                for (int i = 0; i < 9; ++i) L[i] = Rj[i];
            } else {
                const float *Q0 = m->q0m[b];
#pragma unroll
                for (int i = 0; i < 3; ++i)
#pragma unroll
                    for (int j = 0; j < 3; ++j)
                        L[3 * i + j] = Q0[3 * i] * Rj[j] + Q0[3 * i + 1] * Rj[3 + j] + Q0[3 * i + 2] * Rj[6 + j];
            }
        }
        // sweep 1: orientations
        float R[9], Rp[9] = {1.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 1.f};
#pragma unroll
        for (int i = 0; i < 9; ++i) R[i] = L[i];
        for (int lvl = 1; lvl <= maxdepth; ++lvl) {
            float T[9];
#pragma unroll
            for (int i = 0; i < 9; ++i) T[i] = shfl(R[i], psrc);
            if (bdepth == lvl) {
#pragma unroll
                for (int i = 0; i < 9; ++i) Rp[i] = T[i];
#pragma unroll
                for (int i = 0; i < 3; ++i)
#pragma unroll
                    for (int j = 0; j < 3; ++j)
                        R[3 * i + j] = T[3 * i] * L[j] + T[3 * i + 1] * L[3 + j] + T[3 * i + 2] * L[6 + j];
            }
        }
        // segment vectors and the world joint axis (motion subspace: the MJCF axis as written -- Bullet does not normalise it)
        V3 zw = mk(0.f, 0.f, 0.f), e = zw, f = zw, dx = zw;
        if (act && !rootfree) {
            zw = alen * mulR(R, ax);
            e = mulR(Rp, aP);
            f = mulR(R, cOff);
            dx = hinge ? e + f : e + q * zw + f;
            X = dx;
            W = hinge ? qd * zw : mk(0.f, 0.f, 0.f);
        }
        // sweep 2: positions and angular velocities
        V3 Wp = mk(0.f, 0.f, 0.f);
        for (int lvl = 1; lvl <= maxdepth; ++lvl) {
            const V3 px = mk(shfl(X.x, psrc), shfl(X.y, psrc), shfl(X.z, psrc));
            const V3 pw = mk(shfl(W.x, psrc), shfl(W.y, psrc), shfl(W.z, psrc));
            if (bdepth == lvl) { X = X + px; Wp = pw; W = W + pw; }
        }
        V3 AL = mk(0.f, 0.f, 0.f);
        if (act && !rootfree) {
            V = hinge ? cross(Wp, e) + cross(W, f) : cross(Wp, dx) + qd * zw;
            if (bias && hinge) AL = qd * cross(Wp, zw);
        }
        // sweep 3: linear velocities and angular velocity-product accelerations
        V3 ALp = mk(0.f, 0.f, 0.f);
        for (int lvl = 1; lvl <= maxdepth; ++lvl) {
            const V3 pv = mk(shfl(V.x, psrc), shfl(V.y, psrc), shfl(V.z, psrc));
            if (bdepth == lvl) V = V + pv;
            if (bias) {
                const V3 pal = mk(shfl(AL.x, psrc), shfl(AL.y, psrc), shfl(AL.z, psrc));
                if (bdepth == lvl) { ALp = pal; AL = AL + pal; }
            }
        }
        V3 Acc = mk(0.f, 0.f, 0.f);
        if (bias) {
            if (act && !rootfree) {
                Acc = hinge ? cross(ALp, e) + cross(Wp, cross(Wp, e)) + cross(AL, f) + cross(W, cross(W, f))
                            : cross(ALp, dx) + cross(Wp, cross(Wp, dx)) + (2.f * qd) * cross(Wp, zw);
            }
            // sweep 4: linear velocity-product accelerations
            for (int lvl = 1; lvl <= maxdepth; ++lvl) {
                const V3 pa = mk(shfl(Acc.x, psrc), shfl(Acc.y, psrc), shfl(Acc.z, psrc));
                if (bdepth == lvl) Acc = Acc + pa;
            }
        }
        if (act) {
            const V3 A = rootfree ? mk(0.f, 0.f, 0.f) : (hinge ? X - f : X - f - q * zw);
            float *k = kin(gl);
#pragma unroll
            for (int i = 0; i < 9; ++i) k[i] = R[i];
            st3(k + 9, X); st3(k + 12, W); st3(k + 15, V);
            if (bias) { st3(k + 18, AL); st3(k + 21, Acc); }
            st3(k + 24, zw); st3(k + 27, A);
            if (C::KINREG) {
#pragma unroll
                for (int i = 0; i < 9; ++i) kR[i] = R[i];
                kx = X; kw = W; kv = V; kal = AL; ka = Acc;
            }
        }
        __syncwarp();
    }

    __device__ void fk(bool bias) {
        if (C::PLANAR) fk_planar(bias);
        else fk_tree(bias);
    }

    // ---------------------------------------------------------------- contact generation
    // lane = candidate.  Ground: sphere centre / capsule end spheres against z = 0.  Pairs:
    // closest points of two segments.  A candidate is a contact when its distance is below the
    // owning link's manifold breaking threshold; at most MAXC deepest are kept.
    __device__ void collide(V3 xref, bool want_feet) {
        if (C::MAXC == 0) { nc = 0; return; }
        float *cd = sm + C::sCD;
        float *warm = st() + C::oW;
        constexpr int PASSES = C::NSLOT > 0 ? (C::NSLOT + C::LPE - 1) / C::LPE : 1;
        float dist[PASSES];
        bool act[PASSES];
        V3 pa[PASSES], pb[PASSES], nn[PASSES];
        int total = 0;
        unsigned feet = 0;
        // world end points of the geoms behind the pairs, once per pass (the constraint-row block is dead until build_rows)
        float *gw = sm + C::sY;
        if (C::NPAIR > 0) {
            static_assert(C::NPAIR == 0 || C::YSZ >= 6 * MGEOM, "geom end-point staging lives in the constraint-row block");
            for (int g = gl; g < m->ng; g += C::LPE) {
                const float *kg = kin(m->g_body[g]);
                const V3 xg = ld3(kg + 9);
                st3(gw + 6 * g, xg + mulR(kg, ld3(m->g_p0[g])));
                st3(gw + 6 * g + 3, xg + mulR(kg, ld3(m->g_p1[g])));
            }
            __syncwarp();
        }
#pragma unroll
        for (int p = 0; p < PASSES; ++p) {
            const int s = p * C::LPE + gl;
            act[p] = false; dist[p] = CUDART_INF_F;
            pa[p] = pb[p] = mk(0, 0, 0); nn[p] = mk(0, 0, 1);
            if (s < m->ncand) {
                const float *kb = kin(m->c_body[s]);
                const V3 c = ld3(kb + 9) + mulR(kb, ld3(m->c_p[s]));
                const float r = m->c_rad[s];
                dist[p] = c.z - r;
                act[p] = dist[p] < m->c_thr[s];
                pa[p] = mk(c.x, c.y, c.z - r); pb[p] = mk(c.x, c.y, 0.f);
            } else if (C::NPAIR > 0 && s >= C::NCAND && s - C::NCAND < m->npair) {
                const int pi = s - C::NCAND;
                const float *kb = kin(m->p_bb[pi]);
                const int ga = m->p_ga[pi];
                const V3 p1 = ld3(gw + 6 * ga), q1 = ld3(gw + 6 * ga + 3);
                if (C::HASX && m->p_box[pi]) {
                    // capsule / sphere against the cube: closest points of the segment and the box, in the box
                    // frame.  d/dt of the squared distance is monotone in t: 32 bisection steps (oracle.c closest_seg_box)
                    const V3 xb = ld3(kb + 9);
                    const V3 a0 = mulRt(kb, p1 - xb), dd = mulRt(kb, q1 - xb) - a0;
                    const float hh = m->p_half[pi];
                    // cheap exact reject: the box lies inside the sphere of radius sqrt(3) hh around its centre, so the
                    // segment-box distance is at least (segment-centre distance) - sqrt(3) hh.  While the cube is away from
                    // the robot (most of an episode) no lane of the warp enters the bisection.
                    const float dd2 = dot(dd, dd);
                    const V3 xc = a0 + (dd2 > 0.f ? __saturatef(-dot(a0, dd) / dd2) : 0.f) * dd;
                    const bool near = sqrtf(dot(xc, xc)) - 1.7320509f * hh - m->p_ra[pi] < m->p_thr[pi] + 1e-3f;
                    if (near) {
                    float lo_t = 0.f, hi_t = 1.f;
                    for (int it = 0; it < 32; ++it) {
                        const float t = 0.5f * (lo_t + hi_t);
                        const V3 xx = a0 + t * dd;
                        const V3 cc = mk(fminf(fmaxf(xx.x, -hh), hh), fminf(fmaxf(xx.y, -hh), hh), fminf(fmaxf(xx.z, -hh), hh));
                        const float gsl = dot(xx - cc, dd);
                        if (gsl < 0.f) lo_t = t; else hi_t = t;
                    }
                    const float t = 0.5f * (lo_t + hi_t);
                    const V3 xx = a0 + t * dd;
                    const V3 cc = mk(fminf(fmaxf(xx.x, -hh), hh), fminf(fmaxf(xx.y, -hh), hh), fminf(fmaxf(xx.z, -hh), hh));
                    const V3 d = mulR(kb, xx - cc);
                    const float len = sqrtf(dot(d, d)), ra = m->p_ra[pi];
                    dist[p] = len - ra;
                    // an axis that passes through the box: len = 0 up to the bisection's resolution (2e-8 in float32), no
                    // witness direction -> no contact; the cut-off sits above that resolution (same rule in oracle.c)
                    act[p] = dist[p] < m->p_thr[pi] && len > 1e-6f;
                    const V3 n = (1.f / fmaxf(len, 1e-20f)) * d;
                    nn[p] = n; pa[p] = xb + mulR(kb, xx) - ra * n; pb[p] = xb + mulR(kb, cc);
                    }
                } else {
                const int gb = m->p_gb[pi];
                const V3 p2 = ld3(gw + 6 * gb), q2 = ld3(gw + 6 * gb + 3);
                const V3 d1 = q1 - p1, d2 = q2 - p2, r = p1 - p2;
                const float aa = dot(d1, d1), ee = dot(d2, d2), f = dot(d2, r);
                float sp, tp;
                const float EPS = 1e-12f;
                if (aa <= EPS && ee <= EPS) { sp = tp = 0.f; }
                else if (aa <= EPS) { sp = 0.f; tp = __saturatef(f / ee); }
                else {
                    const float cc = dot(d1, r);
                    if (ee <= EPS) { tp = 0.f; sp = __saturatef(-cc / aa); }
                    else {
                        const float bb = dot(d1, d2), den = aa * ee - bb * bb;
                        sp = den > EPS ? __saturatef((bb * f - cc * ee) / den) : 0.f;
                        tp = (bb * sp + f) / ee;
                        if (tp < 0.f) { tp = 0.f; sp = __saturatef(-cc / aa); }
                        else if (tp > 1.f) { tp = 1.f; sp = __saturatef((bb - cc) / aa); }
                    }
                }
                const V3 ca = p1 + sp * d1, cb = p2 + tp * d2, d = ca - cb;
                const float len = sqrtf(dot(d, d)), ra = m->p_ra[pi], rb = m->p_rb[pi];
                dist[p] = len - ra - rb;
                act[p] = dist[p] < m->p_thr[pi] && len > 1e-9f;
                const V3 n = (1.f / fmaxf(len, 1e-20f)) * d;
                nn[p] = n; pa[p] = ca - ra * n; pb[p] = cb + rb * n;
                }
            }
            if (s < C::NSLOT) cd[s] = act[p] ? dist[p] : CUDART_INF_F;
            total += __popc(gballot(act[p]));
            // feet flags = what getContactPoints reports (rs/robot_bases.py:280-281): every foot candidate under its
            // breaking threshold, whether or not the solver's MAXC budget below keeps it
            if (act[p] && s < C::NCAND && m->c_foot[s] >= 0) feet |= 1u << m->c_foot[s];
        }
        __syncwarp();
        // solver budget: MAXC contacts, and no more than the row budget leaves after this sub-step's limit rows (nl)
        const int room = (C::MAXR - nl) / C::RPC;
        const int cap = room < C::MAXC ? room : C::MAXC;
        if (total > cap) ovf = 1;
        if (__any_sync(FULL, total > cap)) {
            // keep the MAXC smallest by (distance, slot)
#pragma unroll
            for (int p = 0; p < PASSES; ++p) {
                const int s = p * C::LPE + gl;
                int rank = 0;
                if (act[p]) {
                    for (int j = 0; j < C::NSLOT; ++j) {
                        const float dj = cd[j];
                        rank += (dj < dist[p] || (dj == dist[p] && j < s)) ? 1 : 0;
                    }
                    if (rank >= cap) act[p] = false;
                }
            }
        }
        int base = 0;
#pragma unroll
        for (int p = 0; p < PASSES; ++p) {
            const int s = p * C::LPE + gl;
            const unsigned bal = gballot(act[p]);
            const int idx = base + __popc(bal & ((1u << gl) - 1u));
            base += __popc(bal);
            if (act[p]) {
                float *ct = sm + C::sCT + idx * C::CTS;
                int ba, bb; float mu;
                if (s < C::NCAND) { ba = m->c_body[s]; bb = -1; mu = m->c_mu[s]; }
                else { ba = m->p_ba[s - C::NCAND]; bb = m->p_bb[s - C::NCAND]; mu = m->p_mu[s - C::NCAND]; }
                ct[0] = __int_as_float(ba); ct[1] = __int_as_float(bb); ct[2] = __int_as_float(s);
                st3(ct + 4, pa[p] - xref); st3(ct + 7, pb[p] - xref); st3(ct + 10, nn[p]);
                ct[13] = dist[p]; ct[14] = mu;
                if (C::TORS) { ct[3] = s < C::NCAND ? m->c_spin[s] : 0.f; ct[15] = s < C::NCAND ? m->c_roll[s] : 0.f; }
            } else if (s < C::NSLOT) {
                warm[s] = 0.f;
            }
        }
        nc = base;
        if (want_feet) {
            float *fo = sm + C::sMISC;
#pragma unroll
            for (int f = 0; f < C::NFEET; ++f) {
                const unsigned any = gballot((feet >> f) & 1u);
                if (gl == 0) fo[f] = any ? 1.f : 0.f;
            }
        }
        __syncwarp();
    }

    // x = L^-1 g with M = L^T L, lane k holds g_k / x_k (one shuffle per column).  Lt[i * LST + k] = L[k][i]: lane k reads
    // consecutive words; entries outside the tree's sparsity pattern are stored zeros.
    __device__ __forceinline__ float fwdsolve(float g, const float *Lt, const float *inv) const {
#pragma unroll
        for (int i = 0; i < C::ND; ++i) {
            const float xi = shfl(g, i) * inv[i];
            if (gl == i) g = xi;
            else if (gl > i) g -= Lt[i * C::LST + gl] * xi;
        }
        return g;
    }

    // ---------------------------------------------------------------- constraint rows
    // lane = row (slot s holds row s * LPE + lane): Jacobian J, Y = L^-1 J^T, right-hand side and bounds.
    // NS = number of register slots built; for NS = 2 the two rows' chains are interleaved statement by statement.
    template <int NS>
    __device__ __forceinline__ void build_rows(const V3 xref, const float h, float (&rhs)[2], float (&dinv)[2], float (&lo)[2],
                                               float (&hi)[2], float (&lmb)[2], float (&mu)[2], float (&Yr)[2][C::ND]) {
        const float *S = st();
        const float *u = uvec();
        const float *lim = sm + C::sLIM;
        const float *SH = sm + C::sSH;
        const float *Lm = sm + C::sL, *inv = sm + C::sINV;
        float *Ym = sm + C::sY, *lam = sm + C::sLAM;
        const float *warm = S + C::oW;
        const int nr = nl + C::RPC * nc;
        float J[NS][C::ND], pen[NS];
        int kind[NS], slot[NS];
        V3 d[NS], pA[NS], pB[NS], pAx[NS], pBx[NS];
        unsigned ma[NS], mb[NS];
        int ldof[NS]; float ldir[NS];
#pragma unroll
        for (int sl = 0; sl < 2; ++sl) {
            rhs[sl] = 0.f; dinv[sl] = 0.f; lo[sl] = 0.f; hi[sl] = 0.f; lmb[sl] = 0.f; mu[sl] = 0.f;
#pragma unroll
            for (int k = 0; k < C::ND; ++k) Yr[sl][k] = 0.f;
        }
#pragma unroll
        for (int sl = 0; sl < NS; ++sl) {
            const int i = sl * C::LPE + gl;
            pen[sl] = 0.f; kind[sl] = -1; slot[sl] = 0; ma[sl] = 0u; mb[sl] = 0u; ldof[sl] = -1; ldir[sl] = 0.f;
            d[sl] = pA[sl] = pB[sl] = pAx[sl] = pBx[sl] = mk(0.f, 0.f, 0.f);
            if (i < nl) {
                kind[sl] = 0;
                const int code = __float_as_int(lim[2 * i]);
                ldof[sl] = code & 0xff;
                ldir[sl] = (code >> 8) ? -1.f : 1.f;
                pen[sl] = lim[2 * i + 1];
            } else if (i < nr) {
                const int ci = i - nl;
                int c;
                const float *ct;
                float coef = 0.f;
                if (ci < nc) { kind[sl] = 1; c = ci; ct = sm + C::sCT + c * C::CTS; d[sl] = ld3(ct + 10); coef = ct[14]; }
                else {
                    // bounded rows: [torsional: spin about n, roll about t1, roll about t2 per contact] then [lateral t1, t2]
                    int bi = ci - nc, ax;
                    if (C::TORS && bi < 3 * nc) { kind[sl] = 3; c = bi / 3; ax = bi - 3 * c; }
                    else { kind[sl] = 2; bi -= 3 * C::TORS * nc; c = bi >> 1; ax = 1 + (bi & 1); }
                    ct = sm + C::sCT + c * C::CTS;
                    const V3 n = ld3(ct + 10);
                    V3 t1, t2;   // btPlaneSpace1
                    if (fabsf(n.z) > 0.70710678f) {
                        const float aa = n.y * n.y + n.z * n.z, k = rsqrtf(aa);
                        t1 = mk(0.f, -n.z * k, n.y * k); t2 = mk(aa * k, -n.x * t1.z, n.x * t1.y);
                    } else {
                        const float aa = n.x * n.x + n.y * n.y, k = rsqrtf(aa);
                        t1 = mk(-n.y * k, n.x * k, 0.f); t2 = mk(-n.z * t1.y, n.z * t1.x, aa * k);
                    }
                    d[sl] = ax == 0 ? n : (ax == 1 ? t1 : t2);
                    coef = kind[sl] == 2 ? ct[14] : (ax == 0 ? ct[3] : ct[15]);
                }
                const int ba = __float_as_int(ct[0]), bb = __float_as_int(ct[1]);
                slot[sl] = __float_as_int(ct[2]);
                pen[sl] = ct[13] + m->slop;
                mu[sl] = coef;
                pA[sl] = ld3(ct + 4); pB[sl] = ld3(ct + 7);
                ma[sl] = m->anc[ba]; mb[sl] = bb >= 0 ? m->anc[bb] : 0u;
                pAx[sl] = pA[sl]; pBx[sl] = pB[sl];     // the same points seen from the cube's centre
                if (C::HASX) { const V3 rc = ld3(kin(C::NB - 1) + 9) - xref; pAx[sl] = pA[sl] - rc; pBx[sl] = pB[sl] - rc; }
            }
        }
        // Jacobian entries, one S_k load for all slots
#pragma unroll
        for (int k = 0; k < C::ND; ++k) {
            const float4 s03 = *reinterpret_cast<const float4 *>(SH + k * 12);
            const float2 s45 = *reinterpret_cast<const float2 *>(SH + k * 12 + 4);
            const V3 so = mk(s03.x, s03.y, s03.z), sv = mk(s03.w, s45.x, s45.y);
            const bool xk = C::HASX && k >= C::XD0;
#pragma unroll
            for (int sl = 0; sl < NS; ++sl) {
                // lanes are rows of different bodies: the ancestor tests are hardly ever warp-uniform, so evaluate and select
                const float va = (C::TORS && kind[sl] == 3) ? dot(d[sl], so) : dot(d[sl], sv + cross(so, xk ? pAx[sl] : pA[sl]));
                float val = ((ma[sl] >> k) & 1u) ? va : 0.f;
                if (C::NPAIR > 0) {
                    const float vb = dot(d[sl], sv + cross(so, xk ? pBx[sl] : pB[sl]));
                    val -= ((mb[sl] >> k) & 1u) ? vb : 0.f;
                }
                val = kind[sl] == 0 ? ((k == ldof[sl]) ? ldir[sl] : 0.f) : val;
                J[sl][k] = val;
            }
        }
        float rel[NS];
#pragma unroll
        for (int sl = 0; sl < NS; ++sl) rel[sl] = 0.f;
#pragma unroll
        for (int k = 0; k < C::ND; ++k) {
            const float uk = u[k];
#pragma unroll
            for (int sl = 0; sl < NS; ++sl) rel[sl] += J[sl][k] * uk;
        }
        // Y = L^-T J^T by back substitution (M = L^T L), one L load for all slots; L[i][k] = Lt[k * LST + i] is structurally
        // zero unless k is an ancestor of i: those terms are skipped at compile time
#pragma unroll
        for (int k = C::ND - 1; k >= 0; --k) {
            float sacc[NS];
#pragma unroll
            for (int sl = 0; sl < NS; ++sl) sacc[sl] = J[sl][k];
#pragma unroll
            for (int i = C::ND - 1; i > k; --i) {
                if (C::coupled(i, k)) {
                    const float lik = Lm[k * C::LST + i];
#pragma unroll
                    for (int sl = 0; sl < NS; ++sl) sacc[sl] -= lik * J[sl][i];
                }
            }
            const float ik = inv[k];
#pragma unroll
            for (int sl = 0; sl < NS; ++sl) J[sl][k] = sacc[sl] * ik;
        }
        const float ih = 1.f / h;
#pragma unroll
        for (int sl = 0; sl < NS; ++sl) {
            const int i = sl * C::LPE + gl;
            float dd = 0.f;
#pragma unroll
            for (int k = 0; k < C::ND; ++k) { dd += J[sl][k] * J[sl][k]; Yr[sl][k] = J[sl][k]; }
            const float di = dd > 1e-12f ? 1.f / dd : 0.f;
            dinv[sl] = di;
            if (kind[sl] == 0) {
                float poserr = -pen[sl] * m->erp_limit * ih;
                if (m->limit_split && !(pen[sl] > m->split_thr)) poserr = 0.f;
                rhs[sl] = (poserr - rel[sl]) * di; lo[sl] = 0.f; hi[sl] = m->limit_max_imp;
            } else if (kind[sl] == 1) {
                float poserr = 0.f, velerr = -rel[sl];
                if (pen[sl] > 0.f) velerr -= pen[sl] * ih; else poserr = -pen[sl] * m->erp_contact * ih;
                rhs[sl] = (poserr + velerr) * di; lo[sl] = 0.f; hi[sl] = 1e10f;
                lmb[sl] = warm[slot[sl]] * m->warm;
            } else if (kind[sl] >= 2) {
                rhs[sl] = -rel[sl] * di;
            } else {
                dinv[sl] = 0.f;
            }
            if (i < C::MAXR) {
#pragma unroll
                for (int k = 0; k < C::ND; ++k) Ym[i * C::LST + k] = J[sl][k];
                lam[i] = lmb[sl];
            }
        }
    }

    // ---------------------------------------------------------------- projected Gauss-Seidel
    // Bullet's row order (btMultiBodyConstraintSolver::solveSingleIteration): joint-limit rows (direction alternating per
    // sweep), contact normals, friction rows.  Lane = row: every lane evaluates the update of its own row from registers each
    // step, the row whose turn it is publishes its impulse change with one shuffle, and all lanes fold it into their residual
    // r = A lambda.  A friction row is bounded by mu times its contact's normal impulse as it stands after this sweep's normal
    // rows (they all precede the friction rows), and skipped while that impulse is zero: the bound is fetched once per sweep.
    // TWO: the warp holds an env with more than LPE rows; rows LPE.. live in a second register slot and are always friction
    // rows (NLIM + MAXC <= LPE), all normals in the first.
    // contact index of the bi-th bounded (torsional / lateral friction) row of this env, see build_rows
    __device__ __forceinline__ int bounded_row_contact(int bi) const {
        if (C::TORS) { if (bi < 3 * nc) return bi / 3; bi -= 3 * nc; }
        return bi >> 1;
    }

    template <bool TWO>
    __device__ __forceinline__ void pgs(const int nrmax, const float (&rhs)[2], const float (&dinv)[2], const float (&lo)[2],
                                        const float (&hi)[2], float (&lmb)[2], const float (&mu)[2], float (&r)[2]) {
        static_assert(C::NLIM + C::MAXC <= C::LPE, "limit and normal rows must fit the first register slot");
        const int niter = m->niter;
        const int nr = nl + C::RPC * nc;
        const float *Ag = sm + C::sA + gl;                                   // this lane's column of A
        const bool fr0 = gl >= nl + nc && gl < nr;
        const int myn0 = fr0 ? nl + bounded_row_contact(gl - nl - nc) : 0;
        const bool fr1 = TWO && C::LPE + gl < nr;
        const int myn1 = fr1 ? nl + bounded_row_contact(C::LPE + gl - nl - nc) : 0;
        const float rhs0 = rhs[0], dinv0 = dinv[0], mu0 = mu[0], rhs1 = rhs[1], dinv1 = dinv[1], mu1 = mu[1];
        float lam0 = lmb[0], r0 = r[0], lam1 = lmb[1], r1 = r[1];
        float lo0 = lo[0], hi0 = hi[0];
        // every lane evaluates its row's candidate change each step: clamp(lambda + rhs - r / A_ii) - lambda with c = lambda + rhs
        // kept beside lambda (it changes only when the lane's own row is taken), one addition less in the serial chain
        float c0 = lam0 + rhs0;       // (the second slot keeps the plain form: one more live register costs it more, -0.8 %)
#define PBG_CD0 (fminf(fmaxf(fmaf(-r0, dinv0, c0), lo0), hi0) - lam0)
#define PBG_TAKE0(cond) if (cond) { lam0 += dl; c0 = lam0 + rhs0; }
        const int nlmax = wmax(nl), ncmax = wmax(nc);
        const int nf0 = (nr < C::LPE ? nr : C::LPE) - nl - nc;               // friction rows of the first slot
        const int nf0max = wmax(nf0);
        for (int it = 0; it < niter; ++it) {
            for (int t = 0; t < nlmax; ++t) {
                // A row index past this env's rows (its warp-mate has more of them) is neutralised OFF the dependent chain: the
                // impulse change that comes back from the shuffle is some finite number of this env's own lanes, no lane
                // takes it and its column of A is read as zeros -- bit-identical to zeroing the change itself
                const bool on = t < nl;
                const int i = on ? ((it & 1) ? t : nl - 1 - t) : 0;
                const float cd = PBG_CD0;
                const float dl = shfl(cd, i);        // unconditional: both env groups of the warp take part
                PBG_TAKE0(on && gl == i)
                const float *arow = Ag + i * C::MAXRP;
                r0 = fmaf(on ? arow[0] : 0.f, dl, r0);
                if (TWO) r1 = fmaf(on ? arow[C::LPE] : 0.f, dl, r1);
            }
            {
                const float *arow = Ag + nl * C::MAXRP;
#pragma unroll 1      // smaller code beats fewer loop branches: unroll 1 / 2 / 4 = 5.014e7 / 4.985e7 / 4.933e7 (Ant)
                for (int j = 0; j < ncmax; ++j) {
                    const bool on = j < nc;
                    const float cd = PBG_CD0;
                    const float dl = shfl(cd, nl + j);                       // (the source lane is taken modulo LPE)
                    PBG_TAKE0(on && gl == nl + j)
                    r0 = fmaf(on ? arow[0] : 0.f, dl, r0);                   // a row index past this env's rows may point at stale storage
                    if (TWO) r1 = fmaf(on ? arow[C::LPE] : 0.f, dl, r1);
                    arow += C::MAXRP;
                }
            }
            // friction bounds of the sweep; a row whose normal impulse is zero is skipped (C4.6): bounds that pin it to its
            // current impulse make its change come out as exactly 0 without a select in the dependent chain
            const float ln0 = shfl(lam0, myn0);
            if (fr0) { hi0 = mu0 * ln0; lo0 = -hi0; if (!(ln0 > 0.f)) { hi0 = lam0; lo0 = lam0; } }
            {
                const float *arow = Ag + (nl + nc) * C::MAXRP;
#pragma unroll 1
                for (int j = 0; j < nf0max; ++j) {
                    const bool on = j < nf0;
                    const float cd = PBG_CD0;
                    const float dl = shfl(cd, nl + nc + j);
                    PBG_TAKE0(on && gl == nl + nc + j)
                    r0 = fmaf(on ? arow[0] : 0.f, dl, r0);
                    if (TWO) r1 = fmaf(on ? arow[C::LPE] : 0.f, dl, r1);
                    arow += C::MAXRP;
                }
            }
            if (TWO) {
                const float ln1 = shfl(lam0, myn1);
                float hi1 = fr1 ? mu1 * ln1 : 0.f, lo1 = -hi1;
                if (!(ln1 > 0.f)) { hi1 = lam1; lo1 = lam1; }    // skipped row (a lane without a second row keeps lam1 = 0)
                const float *arow = Ag + C::LPE * C::MAXRP;
#pragma unroll 2
                for (int t = C::LPE; t < nrmax; ++t) {
                    const float cd = fminf(fmaxf(lam1 + fmaf(-r1, dinv1, rhs1), lo1), hi1) - lam1;
                    const float dl = shfl(cd, t - C::LPE);       // 0 from a lane without a second row
                    if (gl == t - C::LPE) lam1 += dl;
                    r0 = fmaf(arow[0], dl, r0);
                    r1 = fmaf(arow[C::LPE], dl, r1);
                    arow += C::MAXRP;
                }
            }
        }
        lmb[0] = lam0; r[0] = r0; lmb[1] = lam1; r[1] = r1;
#undef PBG_CD0
#undef PBG_TAKE0
    }

    // ---------------------------------------------------------------- dynamics of one substep
    __device__ void substep(bool last) {
        const float h = m->h;
        float *S = st();
        PBG_PHASE_BEGIN;
        // --- joint limit rows (lane = joint): a side is a row only while violated (C4.4)
        float *lim = sm + C::sLIM;
        {
            bool lact = false; float pen = 0.f; int side = 0;
            if (gl < C::NJ && m->jlimited[gl]) {
                const float q = S[C::oQ + gl];
                const float pl = q - m->jlo[gl], pu = m->jhi[gl] - q;
                if (pl <= 0.f) { lact = true; pen = pl; side = 0; }
                else if (pu <= 0.f) { lact = true; pen = pu; side = 1; }
            }
            const unsigned bal = gballot(lact);
            if (C::NLIM > 0 && lact) {
                const int idx = __popc(bal & ((1u << gl) - 1u));
                lim[2 * idx] = __int_as_float((gl + 6 * C::FLOATING) | (side << 8));
                lim[2 * idx + 1] = pen;
            }
            nl = __popc(bal);
        }
        __syncwarp();
        fk(true);
        PBG_PHASE(1);
        const V3 xref = ld3(kin(m->torso_body) + 9);
        collide(xref, last);
        PBG_PHASE(2);

        // --- body wrench + composite inertia entries (lane = body)
        float *acc = sm + C::sACC;
        if (gl < C::NB) {
            // this body's kinematics of the pass, from its kin() record
            const float *kb = kin(gl);
            float R[9];
#pragma unroll
            for (int i = 0; i < 9; ++i) R[i] = C::KINREG ? kR[i] : kb[i];
            const V3 x = C::KINREG ? kx : ld3(kb + 9), w = C::KINREG ? kw : ld3(kb + 12), v = C::KINREG ? kv : ld3(kb + 15);
            const V3 al = C::KINREG ? kal : ld3(kb + 18), a = C::KINREG ? ka : ld3(kb + 21);
            const float bmass = m->mass[gl];
            // world inertia I = R Ib R^T
            float T[9];   // R * Ib
            {
                const float *Ib = m->inertia[gl];
                const float Ixx = Ib[0], Iyy = Ib[1], Izz = Ib[2], Ixy = Ib[3], Ixz = Ib[4], Iyz = Ib[5];
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    T[3 * i + 0] = R[3 * i] * Ixx + R[3 * i + 1] * Ixy + R[3 * i + 2] * Ixz;
                    T[3 * i + 1] = R[3 * i] * Ixy + R[3 * i + 1] * Iyy + R[3 * i + 2] * Iyz;
                    T[3 * i + 2] = R[3 * i] * Ixz + R[3 * i + 1] * Iyz + R[3 * i + 2] * Izz;
                }
            }
            float Iw[6];   // xx yy zz xy xz yz
            Iw[0] = T[0] * R[0] + T[1] * R[1] + T[2] * R[2];
            Iw[1] = T[3] * R[3] + T[4] * R[4] + T[5] * R[5];
            Iw[2] = T[6] * R[6] + T[7] * R[7] + T[8] * R[8];
            Iw[3] = T[0] * R[3] + T[1] * R[4] + T[2] * R[5];
            Iw[4] = T[0] * R[6] + T[1] * R[7] + T[2] * R[8];
            Iw[5] = T[3] * R[6] + T[4] * R[7] + T[5] * R[8];
            auto Imul = [&](V3 q) {
                return mk(Iw[0] * q.x + Iw[3] * q.y + Iw[4] * q.z, Iw[3] * q.x + Iw[1] * q.y + Iw[5] * q.z,
                          Iw[4] * q.x + Iw[5] * q.y + Iw[2] * q.z);
            };
            V3 F = bmass * mk(a.x, a.y, a.z + m->gravity);
            V3 N = Imul(al) + cross(w, Imul(w));
            // Bullet's per-link damping, one entry per massive Bullet link of the body
            const float kd = m->kdamp;
            const V3 wl = mulRt(R, w);
            const float wn = sqrtf(dot(w, w));
            for (int s = m->ds_begin[gl]; s < m->ds_begin[gl + 1]; ++s) {
                const V3 ro = mulR(R, ld3(m->ds_off[s]));
                const V3 vs = v + cross(w, ro);
                const float vn = sqrtf(dot(vs, vs));
                const V3 Fd = (-m->ds_mass[s] * (kd + kd * vn)) * vs;
                const V3 tl = mk(m->ds_inertia[s][0] * wl.x, m->ds_inertia[s][1] * wl.y, m->ds_inertia[s][2] * wl.z);
                const V3 Nd = (-(kd + kd * wn)) * mulR(R, tl);
                F = F - Fd;
                N = N - Nd - cross(ro, Fd);
            }
            // the cube's dofs are referenced to its own centre (it may be far from the robot: a common reference
            // point would bury its 5e-4 kg m^2 inertia under m r^2 in fp32)
            const V3 r = (C::HASX && gl == C::NB - 1) ? mk(0.f, 0.f, 0.f) : x - xref;
            float *ab = acc + gl * 17;
            const float rr = dot(r, r);
            ab[0] = bmass;
            st3(ab + 1, bmass * r);
            ab[4] = Iw[0] + bmass * (rr - r.x * r.x); ab[5] = Iw[1] + bmass * (rr - r.y * r.y);
            ab[6] = Iw[2] + bmass * (rr - r.z * r.z);
            ab[7] = Iw[3] - bmass * r.x * r.y; ab[8] = Iw[4] - bmass * r.x * r.z; ab[9] = Iw[5] - bmass * r.y * r.z;
            st3(ab + 10, N + cross(r, F));
            st3(ab + 13, F);
        }
        __syncwarp();
        // --- subtree sums (lane = component): children precede... bodies are in DFS order, so a
        // reverse sweep folds every body into its parent after its own subtree is complete.
        if (gl < 16) {
            for (int b = C::NB - 1; b >= 1; --b) {
                const int p = m->parent[b];
                if (p >= 0) acc[p * 17 + gl] += acc[b * 17 + gl];
            }
        }
        __syncwarp();

        PBG_PHASE(3);
        // --- motion subspace S_k, H_k = Ic S_k, f_k = tau_k - S_k . W (lane = dof)
        float *SH = sm + C::sSH;
        float *fv = sm + C::sF;
        float Sk[6] = {0, 0, 0, 0, 0, 0}, Hk[6] = {0, 0, 0, 0, 0, 0};
        if (gl < C::ND) {
            int body;
            if (C::FLOATING && gl < 6) {
                body = 0;
                const V3 r0 = ld3(kin(0) + 9) - xref;
                const V3 e = mk(gl % 3 == 0, gl % 3 == 1, gl % 3 == 2);
                if (gl < 3) { Sk[0] = e.x; Sk[1] = e.y; Sk[2] = e.z; const V3 sv = cross(r0, e); Sk[3] = sv.x; Sk[4] = sv.y; Sk[5] = sv.z; }
                else { Sk[3] = e.x; Sk[4] = e.y; Sk[5] = e.z; }
            } else if (C::HASX && gl >= C::XD0) {
                body = C::NB - 1;
                const int c = gl - C::XD0;
                Sk[c] = 1.f;                         // (omega, v) of the cube about its own centre
            } else {
                const int j = gl - 6 * C::FLOATING;
                body = m->jbody[j];
                const float *kb = kin(body);
                const V3 z = ld3(kb + 24);
                if (m->jrev[j]) {
                    const V3 ar = ld3(kb + 27) - xref;
                    const V3 sv = cross(ar, z);
                    Sk[0] = z.x; Sk[1] = z.y; Sk[2] = z.z; Sk[3] = sv.x; Sk[4] = sv.y; Sk[5] = sv.z;
                } else { Sk[3] = z.x; Sk[4] = z.y; Sk[5] = z.z; }
            }
            const float *ab = acc + body * 17;
            const float mc = ab[0];
            const V3 hc = ld3(ab + 1);
            const V3 so = mk(Sk[0], Sk[1], Sk[2]), sv = mk(Sk[3], Sk[4], Sk[5]);
            const V3 Iso = mk(ab[4] * so.x + ab[7] * so.y + ab[8] * so.z, ab[7] * so.x + ab[5] * so.y + ab[9] * so.z,
                              ab[8] * so.x + ab[9] * so.y + ab[6] * so.z);
            const V3 L = Iso + cross(hc, sv);
            const V3 P = mc * sv + cross(so, hc);
            Hk[0] = L.x; Hk[1] = L.y; Hk[2] = L.z; Hk[3] = P.x; Hk[4] = P.y; Hk[5] = P.z;
            const float Ck = dot(so, ld3(ab + 10)) + dot(sv, ld3(ab + 13));
            fv[gl] = tau - Ck;
#pragma unroll
            for (int i = 0; i < 6; ++i) { SH[gl * 12 + i] = Sk[i]; SH[gl * 12 + 6 + i] = Hk[i]; }
        }
        __syncwarp();

        // --- joint-space inertia: lane k keeps M[l][k] for the dofs l it supports (its descendants; the part of its row /
        // column the leaves-first factorisation reads), in registers
        float Mr[C::ND];
        const unsigned down = m->down[gl < C::ND ? gl : 0];
#pragma unroll
        for (int l = 0; l < C::ND; ++l) {
            const float4 s1 = *reinterpret_cast<const float4 *>(SH + l * 12 + 4);
            const float4 s2 = *reinterpret_cast<const float4 *>(SH + l * 12 + 8);
            // H_l = (s1.zw, s2.xyzw)
            const float a_dn = Sk[0] * s1.z + Sk[1] * s1.w + Sk[2] * s2.x + Sk[3] * s2.y + Sk[4] * s2.z + Sk[5] * s2.w;
            Mr[l] = ((down >> l) & 1u) ? a_dn : 0.f;
        }
        PBG_PHASE(4);
        // --- M = L^T L, leaves first (pivots from the last dof down): lane i ends up with column i of L, L[k][i] in Mr[k].
        // At pivot k every lane scales its entry of row k and publishes it; the rank-1 update then touches only the columns
        // that are ancestors of k (compile-time mask), so no fill-in appears outside the tree's pattern.  The right-hand side
        // f rides along as one more column: yf = L^-T f falls out of the same sweep (lane k holds yf_k after pivot k).
        float *col = sm + C::sCOL;
        float *inv = sm + C::sINV;
        float Mf = gl < C::ND ? fv[gl] : 0.f;
        float yf = 0.f;
#pragma unroll
        for (int k = C::ND - 1; k >= 0; --k) {
            const float dkk = shfl(Mr[k], k);
            const float iv = rsqrtf(fmaxf(dkk, 1e-20f));
            const float lki = gl <= k ? Mr[k] * iv : 0.f; // L[k][lane]; nothing above the diagonal
            Mr[k] = lki;
            const float yk = shfl(Mf, k) * iv;            // yf_k
            if (gl == k) { inv[k] = iv; yf = yk; }
            if (C::low(k) != 0u) {
                float *cb = col + (k & 1) * (C::NDP + 4);
                if (gl < C::ND) cb[gl] = lki;
                __syncwarp();
                Mf -= lki * yk;
#pragma unroll
                for (int c = 0; c < k; ++c)
                    if (C::coupled(k, c)) Mr[c] -= lki * cb[c];
            }
        }
        float *Lm = sm + C::sL;          // Lt[i][k] = L[k][i]
        if (gl < C::ND) {
#pragma unroll
            for (int c = 0; c < C::ND; ++c) Lm[gl * C::LST + c] = c >= gl ? Mr[c] : 0.f;
        }
        __syncwarp();
        PBG_PHASE(5);
        float *u = uvec();
        {
            // free acceleration qdd = L^-1 yf, then u <- clamp(u + h qdd): btMultiBody::applyDeltaVeeMultiDof clamps every
            // generalized velocity to +-maxCoordinateVelocity right here, before the constraint rows are built
            const float g = fwdsolve(yf, Lm, inv);
            if (gl < C::ND) {
                const float mv = m->maxvel;
                { const float un = u[gl] + h * g; u[gl] = un > mv ? mv : (un < -mv ? -mv : un); }   // btClamp: a NaN stays a NaN
            }
        }
        __syncwarp();

        PBG_PHASE(6);
        // --- constraint rows (lane = row, two slots)
        const int nr = nl + C::RPC * nc;
        const int nrmax = wmax(nr);
#ifdef PBG_PHASE_CLOCKS
        dbg_nov = max(dbg_nov, nrmax - C::LPE);
#endif
        float *Ym = sm + C::sY;
        float *Am = sm + C::sA;
        float *lam = sm + C::sLAM;
        float *warm = S + C::oW;
        float rhs[2], dinv[2], lo[2], hi[2], lmb[2], mu[2], Yr[2][C::ND];
        // One register slot in the common case.  When a warp needs the second slot (an env with more than LPE
        // rows) both slots are built in ONE pass with their dependent chains interleaved and the S / L loads shared.
        if (nrmax <= C::LPE) build_rows<1>(xref, h, rhs, dinv, lo, hi, lmb, mu, Yr);
        else build_rows<2>(xref, h, rhs, dinv, lo, hi, lmb, mu, Yr);
        __syncwarp();

        PBG_PHASE(7);
        // --- Delassus matrix A = Y Y^T and warm-started residual r = A lambda0
        float r[2] = {0.f, 0.f};
        if (nrmax > 0) {
#pragma unroll 2      // two rows' dot products in flight (Humanoid +0.8 %, Ant +0.15 %)
            for (int i = 0; i < nrmax; ++i) {
                // row i of Y as float4 loads (rows are 16-byte aligned; entries ND.. of a row are padding and never used)
                float yi[C::NDP];
#pragma unroll
                for (int k4 = 0; k4 < C::NDP / 4; ++k4) {
                    const float4 v = *reinterpret_cast<const float4 *>(Ym + i * C::LST + 4 * k4);
                    yi[4 * k4] = v.x; yi[4 * k4 + 1] = v.y; yi[4 * k4 + 2] = v.z; yi[4 * k4 + 3] = v.w;
                }
                const float l0 = lam[i];
#pragma unroll
                for (int sl = 0; sl < 2; ++sl) {
                    if (sl * C::LPE >= nrmax) continue;
                    // second-slot entries against the first LPE rows are the transposes of first-slot entries computed when the
                    // loop reaches that second-slot row (same products, same order: bit-identical); they are copied below
                    if (sl == 1 && i < C::LPE) continue;
                    // two partial sums: halves the dependent FMA chain of the dot product
                    float a0 = 0.f, a1 = 0.f;
#pragma unroll
                    for (int k = 0; k + 1 < C::ND; k += 2) { a0 = fmaf(yi[k], Yr[sl][k], a0); a1 = fmaf(yi[k + 1], Yr[sl][k + 1], a1); }
                    if (C::ND & 1) a0 = fmaf(yi[C::ND - 1], Yr[sl][C::ND - 1], a0);
                    const float aij = a0 + a1;
                    const int jrow = sl * C::LPE + gl;
                    if (jrow < C::MAXR) Am[i * C::MAXRP + jrow] = aij;
                    r[sl] += aij * l0;
                }
            }
            if (nrmax > C::LPE) {
                __syncwarp();
                const int jrow = C::LPE + gl;
                if (jrow < C::MAXR) {
#pragma unroll 4
                    for (int i = 0; i < C::LPE; ++i) {
                        const float a = Am[jrow * C::MAXRP + i];
                        Am[i * C::MAXRP + jrow] = a;
                        r[1] = fmaf(a, lam[i], r[1]);
                    }
                }
            }
        }
        __syncwarp();

        PBG_PHASE(8);
        // a second block barrier per sub-step, right before the solver: PGS is the longest dependent chain of the sub-step and the
        // warps that enter it together fetch its code once (Ant +1.2 %, Humanoid +1.0 %; the short planar steps lose 1 % to it)
        if (C::WARPS > 2 && !C::PLANAR) __syncthreads();
        // --- projected Gauss-Seidel on lambda (see pgs() below)
        if (nrmax > C::LPE) pgs<true>(nrmax, rhs, dinv, lo, hi, lmb, mu, r);
        else pgs<false>(nrmax, rhs, dinv, lo, hi, lmb, mu, r);
        // --- publish lambda, store warm-start impulses
#pragma unroll
        for (int sl = 0; sl < 2; ++sl) {
            const int i = sl * C::LPE + gl;
            if (i < nr && i < C::MAXR) {
                lam[i] = lmb[sl];
                if (dbg) {
                    dbg[0] = (float)nl; dbg[1] = (float)nc;
                    dbg[2 + 4 * i] = rhs[sl]; dbg[3 + 4 * i] = dinv[sl]; dbg[4 + 4 * i] = lmb[sl]; dbg[5 + 4 * i] = r[sl];
                }
                if (i >= nl && i < nl + nc) warm[__float_as_int(sm[C::sCT + (i - nl) * C::CTS + 2])] = lmb[sl];
            }
        }
        __syncwarp();

        PBG_PHASE(9);
        // --- du = L^-T (Y^T lambda)   (lane = dof), second clamp (btMultiBody::processDeltaVeeMultiDof2)
        float g = 0.f;
        if (gl < C::ND) {
            for (int i = 0; i < nr; ++i) g += Ym[i * C::LST + gl] * lam[i];
        }
        if (wmax(nr) > 0) g = fwdsolve(g, Lm, inv);
        __syncwarp();
        if (gl < C::ND) {
            const float mv = m->maxvel;
            { const float un = u[gl] + g; u[gl] = un > mv ? mv : (un < -mv ? -mv : un); }
        }
        __syncwarp();
        // --- semi-implicit Euler (btMultiBody::stepPositionsMultiDof)
        if (gl < C::NJ) S[C::oQ + gl] += h * u[6 * C::FLOATING + gl];
        if ((C::FLOATING && gl == C::LPE - 1) || (C::HASX && gl == C::LPE - 2)) {
            const bool cube = C::HASX && gl == C::LPE - 2;
            float *Q = cube ? S + C::oX + 3 : S + 3;
            const V3 om = ld3(cube ? u + C::XD0 : u);
            float wn = sqrtf(dot(om, om));
            if (wn * h > 0.25f * CUDART_PI_F) wn = 0.5f * (0.5f * CUDART_PI_F) / h;
#ifndef PBG_QTRIG
            // sin(x) / wn = (h / 2) sinc(x) and cos(x) of the half angle x = wn h / 2 <= pi / 8 (the clamp above) as Taylor polynomials
            // in x^2: truncation below 2e-10 relative, i.e. as accurate as sinf / cosf in fp32, without their range reduction, the
            // division and btTransformUtil's small-angle branch (whose two-term series this reproduces for wn < 0.001)
            const float xh = 0.5f * wn * h, x2 = xh * xh;
            const float sc = 0.5f * h * fmaf(x2, fmaf(x2, fmaf(x2, fmaf(x2, 2.7557319e-6f, -1.9841270e-4f), 8.3333333e-3f), -0.16666667f), 1.f);
            const float aw = fmaf(x2, fmaf(x2, fmaf(x2, fmaf(x2, fmaf(x2, -2.7557319e-7f, 2.4801587e-5f), -1.3888889e-3f), 4.1666667e-2f), -0.5f), 1.f);
            const float ax = om.x * sc, ay = om.y * sc, az = om.z * sc;
#else
            float sc;
            if (wn < 0.001f) sc = 0.5f * h - h * h * h * 0.020833333333f * wn * wn;
            else sc = sinf(0.5f * wn * h) / wn;
            const float ax = om.x * sc, ay = om.y * sc, az = om.z * sc, aw = cosf(0.5f * wn * h);
#endif
            const float bx = Q[0], by = Q[1], bz = Q[2], bw = Q[3];
            float qx = aw * bx + ax * bw + ay * bz - az * by;
            float qy = aw * by - ax * bz + ay * bw + az * bx;
            float qz = aw * bz + ax * by - ay * bx + az * bw;
            float qw = aw * bw - ax * bx - ay * by - az * bz;
            const float nn = rsqrtf(qx * qx + qy * qy + qz * qz + qw * qw);
            Q[0] = qx * nn; Q[1] = qy * nn; Q[2] = qz * nn; Q[3] = qw * nn;
        }
        if (C::FLOATING && gl < 3) S[gl] += h * u[3 + gl];
        if (C::HASX && gl >= 3 && gl < 6) S[C::oX + gl - 3] += h * u[C::XD0 + gl];
        __syncwarp();
        PBG_PHASE(10);
    }

    // ---------------------------------------------------------------- task layer
    // WalkerBase.calc_state + WalkerBaseBulletEnv._step reward terms
    // (rs/robot_locomotors.py:31-79, rs/gym_locomotion_envs.py:54-114).  Returns done (group-uniform).
    // `reset_pass`: first observation of an episode (latches initial_z, sets potential, no reward).
    __device__ bool walker_task(const float *act, float *obs_out, float *rew_out, float *terms_out, bool reset_pass, bool pred) {
        float *S = st();
        float *T = S + C::oT;
        fk(false);
        // joint features (lane = action index)
        float jpos = 0.f, jvel = 0.f, aval = 0.f;
        bool atlim = false;
        if (gl < C::NACT) {
            const int j = m->act_joint[gl];
            float pos = S[C::oQ + j];
            const float vel = S[C::oU + 6 * C::FLOATING + j];
            if (m->jlimited[j]) {
                const float lo = m->jlo[j], hi = m->jhi[j];
                pos = 2.f * (pos - 0.5f * (lo + hi)) / (hi - lo);
            }
            jpos = pos; jvel = (m->jrev[j] ? 0.1f : 0.5f) * vel;
            atlim = fabsf(jpos) > 0.99f;
            aval = act ? act[gl] : 0.f;
        }
        const int nlim = __popc(gballot(atlim));
        const float se = gsum(gl < C::NACT ? fabsf(aval * jvel) : 0.f);
        const float ss = gsum(gl < C::NACT ? aval * aval : 0.f);
        // mean x / y of robot.parts (lane = body)
        float px = 0.f, py = 0.f, pn = 0.f;
        if (gl < C::NB) {
            const float *kb = kin(gl);
            const V3 o = mulR(kb, ld3(m->part_sum[gl]));
            pn = m->part_cnt[gl];
            px = pn * kb[9] + o.x; py = pn * kb[10] + o.y;
        }
        px = gsum(px); py = gsum(py); pn = gsum(pn);
        const float floorp = T[T_FLOOR];
        const float bx = px / (pn + floorp), by = py / (pn + floorp);
        // torso pose / speed
        const float *kt = kin(m->torso_body);
        const V3 toff = mulR(kt, ld3(m->torso_off));
        const float z = kt[11] + toff.z;
        const V3 tsp = ld3(kt + 15) + cross(ld3(kt + 12), toff);
        // getEulerFromQuaternion(torso orientation) (rs/robot_bases.py:216-217) taken straight from the
        // rotation matrix: roll = atan2(R21, R22), pitch = asin(-R20), yaw = atan2(R10, R00).  The pitch
        // uses atan2(-R20, sqrt(R00^2 + R10^2)): identical value, but well conditioned in fp32 near +-pi/2.
        float roll, pitch;
        if (C::PLANAR) {
            // R = Ry(theta): roll = atan2(+0, cos) and pitch = atan2(sin, |cos|) follow from the accumulated angle itself
            const float thw = kt[30] - 6.2831853071795865f * rintf(kt[30] * 0.15915494309189535f);      // (-pi, pi]
            const bool fwd = kt[8] >= 0.f;
            roll = fwd ? 0.f : 3.14159265358979324f;
            pitch = fwd ? thw : (thw >= 0.f ? 3.14159265358979324f - thw : -3.14159265358979324f - thw);
        } else {
            roll = atan2_shared(kt[7], kt[8]);
            pitch = atan2_shared(-kt[6], sqrtf(kt[0] * kt[0] + kt[3] * kt[3]));
        }
        // yaw = atan2(R10, R00) and the target bearing theta = atan2(dy, dx) enter the observation only through sin / cos of
        // (theta - yaw) and of -yaw: taken from the normalised vectors directly, no trigonometric calls (atan2(0, 0) = 0 kept)
        const float nyaw = sqrtf(kt[0] * kt[0] + kt[3] * kt[3]);
        const float cyaw = nyaw > 0.f ? kt[0] / nyaw : 1.f, syaw = nyaw > 0.f ? kt[3] / nyaw : 0.f;
        float initz = T[T_INITZ];
        if (reset_pass) {
            initz = m->initial_z >= 0.f ? m->initial_z : z;
        }
        float tx = T[T_TX], ty = T[T_TY];
        double ddy = (double)ty - (double)by, ddx = (double)tx - (double)bx;
        float thx = tx - bx, thy = ty - by;
        double dist = sqrt(ddy * ddy + ddx * ddx);
        const int kind = m->kind;
        // compile-time capabilities of the configuration prune the blocks a kind can never enter (they are large: the cube attack
        // alone holds five Philox evaluations)
        const bool flagrun = C::NB_ROBOT == 18 && C::OBS == 44 && (kind == 7 || kind == 8);
        float flag_timeout = 0.f;
        int flag_cnt = 0;
        // HumanoidFlagrunHarder.calc_potential (rs/robot_locomotors.py:280-302) has side effects (crawl
        // bookkeeping) and is evaluated once per flag move and once per step, in that order
        const bool harder = C::HASX && kind == 8;
        bool crawl_has = harder && T[T_CRAWL_HAS] != 0.f;
        double crawl_start = harder ? __hiloint2double(__float_as_int(T[T_CRAWL_START_HI]), __float_as_int(T[T_CRAWL_START_LO])) : 0.0;
        double crawl_ign = harder ? __hiloint2double(__float_as_int(T[T_CRAWL_IGN_HI]), __float_as_int(T[T_CRAWL_IGN_LO])) : 0.0;
        const float leak = fminf(fmaxf(z, 0.f), 0.8f) / 0.8f + 1.0f;          // potential_leak()
        auto harder_potential = [&](double d) {
            double frp = -d * m->inv_dt_scene;
            if (z < 0.8f) {
                if (!crawl_has) { crawl_start = frp - crawl_ign; crawl_has = true; }
                crawl_ign = frp - crawl_start;
                frp = crawl_start;
            } else {
                frp -= crawl_ign;
                crawl_has = false;
            }
            return frp + (double)leak * 100.0;
        };
        if (flagrun) {
            // HumanoidFlagrun.calc_state (rs/robot_locomotors.py:220-227): the flag moves when it is reached
            // (dist < 1) or after 600 / frame_skip steps; the observation is then taken against the new flag
            flag_timeout = T[T_FLAGTIMEOUT] - 1.f;
            flag_cnt = __float_as_int(T[T_FLAGCNT]);
            if (dist < 1.0 || flag_timeout <= 0.f) {
                const unsigned ep = (unsigned)__float_as_int(T[T_EPISODE]);
                tx = 0.5f * rng_uniform(rng_seed, rng_env, ep, 1u, 2u * flag_cnt, -m->halflen, m->halflen);
                ty = 0.5f * rng_uniform(rng_seed, rng_env, ep, 1u, 2u * flag_cnt + 1u, -m->halfwidth, m->halfwidth);
                flag_cnt += 1;
                flag_timeout = 600.f / (float)m->nsub;
                ddy = (double)ty - (double)by; ddx = (double)tx - (double)bx;
                thx = tx - bx; thy = ty - by;
                dist = sqrt(ddy * ddy + ddx * ddx);
                if (harder) (void)harder_potential(dist);     // robot.potential = calc_potential(): value unused (quirk Q5)
            }
        }
        const float sy = -syaw, cy = cyaw;                                      // sin / cos of -yaw
        const float vx = cy * tsp.x - sy * tsp.y, vy = sy * tsp.x + cy * tsp.y, vz = tsp.z;
        const float nth = sqrtf(thx * thx + thy * thy);
        const float cth = nth > 0.f ? thx / nth : 1.f, sth = nth > 0.f ? thy / nth : 0.f;
        const float sa = sth * cyaw - cth * syaw, ca = cth * cyaw + sth * syaw;  // sin / cos of (theta - yaw)
        // observation, clipped to +-5
        auto clip5 = [](float q) { return fminf(fmaxf(q, -5.f), 5.f); };
        const float o0 = clip5(z - initz);
        const bool mjf = C::OBS > 64 && (kind == 14 || kind == 15);
        if (mjf && obs_out && pred) {
            // MuJoCo-style Ant / Humanoid: [z, quat xyzw, q] ++ [base v, base omega, qdot] (unclipped), zeros after that
            if (gl == 0) { obs_out[0] = S[2]; obs_out[1] = S[3]; obs_out[2] = S[4]; obs_out[3] = S[5]; obs_out[4] = S[6]; }
            if (gl < 3) { obs_out[5 + C::NJ + gl] = S[C::oU + 3 + gl]; obs_out[8 + C::NJ + gl] = S[C::oU + gl]; }
            if (gl < C::NACT) {
                const int j = m->act_joint[gl];
                obs_out[5 + gl] = S[C::oQ + j]; obs_out[11 + C::NJ + gl] = S[C::oU + 6 * C::FLOATING + j];
            }
        }
        if (!mjf && obs_out && pred) {
            if (gl == 0) {
                obs_out[0] = o0; obs_out[1] = clip5(sa); obs_out[2] = clip5(ca); obs_out[3] = clip5(0.3f * vx);
                obs_out[4] = clip5(0.3f * vy); obs_out[5] = clip5(0.3f * vz); obs_out[6] = clip5(roll); obs_out[7] = clip5(pitch);
            }
            if (gl < C::NACT) { obs_out[8 + 2 * gl] = clip5(jpos); obs_out[9 + 2 * gl] = clip5(jvel); }
            if (gl < C::NFEET) obs_out[8 + 2 * C::NACT + gl] = clip5(S[C::oF + gl]);   // previous step's flags (quirk Q2)
        }
        bool done = false;
        int h_frame = 0, h_og = 0, h_att = 0;
        float h_alive = 0.f;
        if (harder) {
            h_frame = __float_as_int(T[T_FRAME]); h_og = __float_as_int(T[T_ONGROUND]); h_att = __float_as_int(T[T_ATTACKS]);
            if (!reset_pass) {
                // alive_bonus (rs/robot_locomotors.py:250-273); it runs before the step's calc_potential
                if (h_frame % 30 == 0 && h_frame > 100 && h_og == 0) {
                    // cube attack (rs/robot_locomotors.py:251-266): thrown from 4 m away, 1 m up, at where the robot
                    // will be when it arrives; the cube keeps its orientation, its spin is zeroed
                    if (C::HASX && gl == 0 && pred) {
                        const unsigned ep = (unsigned)__float_as_int(T[T_EPISODE]);
                        const float angle = rng_uniform(rng_seed, rng_env, ep, 2u, 5u * h_att, -3.14f, 3.14f);
                        const float speed = rng_uniform(rng_seed, rng_env, ep, 2u, 5u * h_att + 1u, 20.f, 30.f);
                        const float ttt = 4.0f / speed;
                        const V3 tgt = mk(bx + tsp.x * ttt, by + tsp.y * ttt, z + tsp.z * ttt);
                        const float2 sca_ = sincos_shared(angle);
                        const float sa_ = sca_.x, ca_ = sca_.y;
                        const V3 pos = mk(tgt.x + 4.0f * ca_, tgt.y + 4.0f * sa_, tgt.z + 1.0f);
                        const V3 dv = tgt - pos;
                        const float sc = speed * rsqrtf(dot(dv, dv));
                        float *X = S + C::oX, *XU = S + C::oU + C::XD0;
                        X[0] = pos.x; X[1] = pos.y; X[2] = pos.z;
                        XU[0] = 0.f; XU[1] = 0.f; XU[2] = 0.f;
                        XU[3] = dv.x * sc + rng_uniform(rng_seed, rng_env, ep, 2u, 5u * h_att + 2u, -1.f, 1.f);
                        XU[4] = dv.y * sc + rng_uniform(rng_seed, rng_env, ep, 2u, 5u * h_att + 3u, -1.f, 1.f);
                        XU[5] = dv.z * sc + rng_uniform(rng_seed, rng_env, ep, 2u, 5u * h_att + 4u, -1.f, 1.f);
                    }
                    h_att += 1;
                }
                const float zz8 = clip5(z - initz) + initz;
                if (zz8 < 0.8f) h_og += 1; else if (h_og > 0) h_og -= 1;
                h_frame += 1;
                h_alive = h_og < 170 ? leak : -1.f;
            }
        }
        const double pot_new = harder ? harder_potential(dist) : -dist * m->inv_dt_scene;
        {
            // alive uses state[0] + initial_z after the float32 round trip (rs/gym_locomotion_envs.py:61)
            // (MuJoCo-style variants pass state[0] = torso z itself: mujoco/gym_locomotion_envs.py:60)
            const float zz = mjf ? (S[2] + initz) : ((m->initial_z >= 0.f) ? (o0 + initz) : (float)((double)o0 + (double)initz));
            const float *fc = S + C::oF;
            float alive;
            if (kind == 2 || kind == 3) alive = (zz > 0.8f && fabsf(pitch) < 1.0f) ? 1.f : -1.f;
            else if (kind == 4) alive = (fabsf(pitch) < 1.0f && fc[1] == 0.f && fc[2] == 0.f && fc[4] == 0.f && fc[5] == 0.f) ? 1.f : -1.f;
            else if (kind == 5 || kind == 14) alive = zz > 0.26f ? 1.f : -1.f;
            else if (kind == 8) alive = h_alive;
            else alive = zz > 0.78f ? 2.f : -1.f;
            done = alive < 0.f;
            // non-finite observation ends the episode (rs/gym_locomotion_envs.py:63-65)
            bool bad = !(isfinite(o0) && isfinite(sa) && isfinite(ca) && isfinite(vx) && isfinite(vy) && isfinite(vz) &&
                         isfinite(roll) && isfinite(pitch));
            bad = bad || (gl < C::NACT && !(isfinite(jpos) && isfinite(jvel)));
            const bool anybad = gballot(bad) != 0u;
            done = (done || anybad) && !reset_pass;
            const double pot_old = __hiloint2double(__float_as_int(T[T_POT_HI]), __float_as_int(T[T_POT_LO]));
            const float progress = (float)(pot_new - pot_old);
            const float elec = m->elec_cost * (se / (float)C::NACT) + m->stall_cost * (ss / (float)C::NACT);
            const float limc = m->limit_cost * (float)nlim;
            // A non-finite state (a rare fp32 blow-up under saturated random torques: ~3 per 10^6 Humanoid env steps, or a
            // non-finite action) ends the episode like the reference's "~INF~" branch; its reward is reported as 0 rather
            // than NaN so that one bad env cannot poison a learner's batch statistics.
            if (gl == 0 && pred && !reset_pass && anybad) {
                if (rew_out) *rew_out = 0.f;
                if (terms_out) { terms_out[0] = terms_out[1] = terms_out[2] = terms_out[3] = terms_out[4] = 0.f; }
                T[T_HAVEZ] = 2.f;
            } else
            if (gl == 0 && pred && !reset_pass && mjf) {
                // WalkerBaseMuJoCoEnv._step: [alive, progress, joints_at_limit_cost, feet_collision_cost]
                if (rew_out) *rew_out = alive + progress + limc;
                if (terms_out) { terms_out[0] = alive; terms_out[1] = progress; terms_out[2] = limc; terms_out[3] = 0.f; terms_out[4] = 0.f; }
            }
            if (gl == 0 && pred && !reset_pass && !mjf && !anybad) {
                if (rew_out) *rew_out = alive + progress + elec + limc;
                if (terms_out) { terms_out[0] = alive; terms_out[1] = progress; terms_out[2] = elec; terms_out[3] = limc; terms_out[4] = 0.f; }
            }
        }
        __syncwarp();
        if (gl == 0 && pred && harder) {
            T[T_FRAME] = __int_as_float(h_frame); T[T_ONGROUND] = __int_as_float(h_og); T[T_ATTACKS] = __int_as_float(h_att);
            T[T_CRAWL_HAS] = crawl_has ? 1.f : 0.f;
            T[T_CRAWL_START_LO] = __int_as_float(__double2loint(crawl_start)); T[T_CRAWL_START_HI] = __int_as_float(__double2hiint(crawl_start));
            T[T_CRAWL_IGN_LO] = __int_as_float(__double2loint(crawl_ign)); T[T_CRAWL_IGN_HI] = __int_as_float(__double2hiint(crawl_ign));
        }
        if (gl == 0 && pred && flagrun) {
            T[T_TX] = tx; T[T_TY] = ty; T[T_FLAGTIMEOUT] = flag_timeout; T[T_FLAGCNT] = __int_as_float(flag_cnt);
        }
        if (gl == 0 && pred) {
            T[T_POT_LO] = __int_as_float(__double2loint(pot_new));
            T[T_POT_HI] = __int_as_float(__double2hiint(pot_new));
            if (reset_pass) { T[T_INITZ] = initz; T[T_FLOOR] = 1.f; }   // quirk Q1: the floor joins robot.parts after the reset's calc_state
        }
        __syncwarp();
        return done;
    }

    // InvertedPendulum (rs/robot_pendula.py:27-51, rs/gym_pendulum_envs.py:26-39)
    __device__ bool pendulum_task(float *obs_out, float *rew_out, float *terms_out, bool reset_pass, bool pred) {
        const float *S = st();
        const float xx = S[C::oQ], th = S[C::oQ + 1], vx = S[C::oU], thd = S[C::oU + 1];
        float s, c;
        sincosf(th, &s, &c);
        if (m->kind == 9 || m->kind == 11) {
            // InvertedDoublePendulum (rs/robot_pendula.py:57-87, rs/gym_pendulum_envs.py:50-83): pole2's link COM is
            // the middle of the second pole; reward = 10 - 0.01 x^2 - (y + 0.3 - 2)^2, done when y + 0.3 <= 1
            fk(false);
            const float *k2 = kin(C::NB - 1);
            const float px = k2[9], py = k2[11];
            const float ga = S[C::oQ + 2], gad = S[C::oU + 2];
            float s2, c2;
            sincosf(ga, &s2, &c2);
            const float dist = 0.01f * px * px + (py + 0.3f - 2.f) * (py + 0.3f - 2.f);
            if (m->kind == 11) {
                // MuJoCo-style variant (pybulletgym/envs/mujoco/robot_pendula.py:73-88, gym_pendulum_envs.py:56-69):
                // [x, sin, sin, cos, cos, clip(qvel, +-10), qfrc_constraint = 0], velocity penalty in the reward
                const float velp = 1e-3f * thd * thd + 5e-3f * gad * gad;
                if (gl == 0 && pred) {
                    if (obs_out) {
                        obs_out[0] = xx; obs_out[1] = s; obs_out[2] = s2; obs_out[3] = c; obs_out[4] = c2;
                        obs_out[5] = fminf(fmaxf(vx, -10.f), 10.f); obs_out[6] = fminf(fmaxf(thd, -10.f), 10.f);
                        obs_out[7] = fminf(fmaxf(gad, -10.f), 10.f); obs_out[8] = 0.f; obs_out[9] = 0.f; obs_out[10] = 0.f;
                    }
                    if (!reset_pass) {
                        if (rew_out) *rew_out = 10.f - dist - velp;
                        if (terms_out) { terms_out[0] = 10.f; terms_out[1] = -dist; terms_out[2] = -velp; terms_out[3] = 0.f; terms_out[4] = 0.f; }
                    }
                }
                return !reset_pass && py + 0.3f <= 1.f;
            }
            if (gl == 0 && pred) {
                if (obs_out) {
                    obs_out[0] = xx; obs_out[1] = vx; obs_out[2] = px; obs_out[3] = c; obs_out[4] = s; obs_out[5] = thd;
                    obs_out[6] = c2; obs_out[7] = s2; obs_out[8] = gad;
                }
                if (!reset_pass) {
                    if (rew_out) *rew_out = 10.f - dist;
                    if (terms_out) { terms_out[0] = 10.f; terms_out[1] = -dist; terms_out[2] = 0.f; terms_out[3] = 0.f; terms_out[4] = 0.f; }
                }
            }
            return !reset_pass && py + 0.3f <= 1.f;
        }
        if (gl == 0 && pred) {
            if (obs_out) { obs_out[0] = xx; obs_out[1] = vx; obs_out[2] = c; obs_out[3] = s; obs_out[4] = thd; }
            if (!reset_pass) {
                const float rw = (m->kind == 1) ? c : 1.f;
                if (rew_out) *rew_out = rw;
                if (terms_out) { terms_out[0] = rw; terms_out[1] = 0.f; terms_out[2] = 0.f; terms_out[3] = 0.f; terms_out[4] = 0.f; }
            }
        }
        return !reset_pass && m->kind == 0 && fabsf(th) > 0.2f;
    }

    // MuJoCo-style Hopper / Walker2D (pybulletgym/envs/mujoco/robot_locomotors.py:86-165, gym_locomotion_envs.py:121-206):
    // obs = qpos[1:] ++ clip(qvel, +-10) over all dofs (root joints included); reward = [dx / dt, 1, -1e-3 |a|^2].
    // HalfCheetah (kind 16, mujoco/robot_locomotors.py:169-211, gym_locomotion_envs.py:211-244): qvel unclipped,
    // reward = [dx / dt, -0.1 |a|^2], never done.
    __device__ bool mjwalker_task(const float *act, float *obs_out, float *rew_out, float *terms_out, bool reset_pass, bool pred) {
        float *S = st();
        float *T = S + C::oT;
        fk(false);
        const float *kt = kin(m->torso_body);
        const float x = kt[9] + mulR(kt, ld3(m->torso_off)).x;       // robot_body.get_pose()[0]: torso link COM
        const float q = gl < C::NJ ? S[C::oQ + gl] : 0.f;
        const bool cheetah = m->kind == 16;
        const float qd = gl < C::NJ ? (cheetah ? S[C::oU + gl] : fminf(fmaxf(S[C::oU + gl], -10.f), 10.f)) : 0.f;
        const float aval = (act && gl < C::NACT) ? act[gl] : 0.f;
        const float ss = gsum(aval * aval);
        // done = not (finite and |state[2:]| < 100 and height / angle window)
        bool bad = gl < C::NJ && (!(isfinite(q) && isfinite(qd)) || (!cheetah && gl >= 3 && !(fabsf(q) < 100.f)));   // qvel is clipped to 10
        const bool anybad = gballot(bad) != 0u;
        const float height = S[C::oQ + 1], ang = S[C::oQ + 2];
        bool ok = !anybad;
        if (m->kind == 12) ok = ok && height > -0.3f && fabsf(ang) < 0.2f;
        else if (!cheetah) ok = ok && 1.0f > height && height > -0.2f && -1.0f < ang && ang < 1.0f;
        if (pred) {
            if (obs_out && gl < C::NJ) {
                if (gl >= 1) obs_out[gl - 1] = q;
                obs_out[C::NJ - 1 + gl] = qd;
            }
            if (gl == 0) {
                if (!reset_pass) {
                    const float pot = anybad ? 0.f : (float)(((double)x - (double)T[T_POT_LO]) / m->dt_scene);
                    const float pc = anybad ? 0.f : (cheetah ? -0.1f : -1e-3f) * ss;
                    const float al = (anybad || cheetah) ? 0.f : 1.0f;        // a non-finite state ends the episode with reward 0
                    if (rew_out) *rew_out = pot + al + pc;
                    if (terms_out) {
                        terms_out[0] = pot; terms_out[1] = cheetah ? pc : al; terms_out[2] = cheetah ? 0.f : pc; terms_out[3] = 0.f; terms_out[4] = 0.f;
                    }
                    if (anybad) T[T_HAVEZ] = 2.f;
                }
                T[T_POT_LO] = x;
            }
        }
        __syncwarp();
        return !reset_pass && !ok;
    }

    // Reacher (rs/robot_manipulators.py:28-50, rs/gym_manipulator_envs.py:15-32): dofs joint0, joint1, target_x, target_y.
    // Never done; reward = potential change + electricity + stuck-joint cost.
    __device__ bool reacher_task(const float *act, float *obs_out, float *rew_out, float *terms_out, bool reset_pass, bool pred) {
        float *S = st();
        float *T = S + C::oT;
        fk(false);
        const float *kf = kin(m->aux_body[0]), *kt = kin(m->aux_body[1]);
        const V3 d = (ld3(kf + 9) + mulR(kf, ld3(m->aux_off[0]))) - (ld3(kt + 9) + mulR(kt, ld3(m->aux_off[1])));
        const float theta = S[C::oQ], theta_dot = 0.1f * S[C::oU];
        const float lo = m->jlo[1], hi = m->jhi[1];
        const float gamma = 2.f * (S[C::oQ + 1] - 0.5f * (lo + hi)) / (hi - lo), gamma_dot = 0.1f * S[C::oU + 1];
        float st_, ct_;
        sincosf(theta, &st_, &ct_);
        const double pot_new = -100.0 * sqrt((double)d.x * d.x + (double)d.y * d.y + (double)d.z * d.z);
        if (gl == 0 && pred) {
            if (obs_out) {
                obs_out[0] = S[C::oQ + 2]; obs_out[1] = S[C::oQ + 3]; obs_out[2] = d.x; obs_out[3] = d.y;
                obs_out[4] = ct_; obs_out[5] = st_; obs_out[6] = theta_dot; obs_out[7] = gamma; obs_out[8] = gamma_dot;
            }
            if (!reset_pass) {
                const double pot_old = __hiloint2double(__float_as_int(T[T_POT_HI]), __float_as_int(T[T_POT_LO]));
                const float a0 = act ? act[0] : 0.f, a1 = act ? act[1] : 0.f;
                const float progress = (float)(pot_new - pot_old);
                const float elec = -0.10f * (fabsf(a0 * theta_dot) + fabsf(a1 * gamma_dot)) - 0.01f * (fabsf(a0) + fabsf(a1));
                const float stuck = fabsf(fabsf(gamma) - 1.f) < 0.01f ? -0.1f : 0.f;
                if (rew_out) *rew_out = progress + elec + stuck;
                if (terms_out) { terms_out[0] = progress; terms_out[1] = elec; terms_out[2] = stuck; terms_out[3] = 0.f; terms_out[4] = 0.f; }
            }
            T[T_POT_LO] = __int_as_float(__double2loint(pot_new));
            T[T_POT_HI] = __int_as_float(__double2hiint(pot_new));
        }
        __syncwarp();
        return false;
    }

    // Fused policy inference: the observation staged in shared memory goes through the MLP, lane = output
    // unit, weights streamed from global memory (every env group of the grid reads the same addresses: L1 / L2 hits).
    __device__ void policy_actions(const PolicyDev &P, const float *obs_s, float *hid, float *act_out) {
        float *h1v = hid, *h2v = hid + P.h1;
        for (int o = gl; o < P.h1; o += C::LPE) {
            float acc = __ldg(P.b1 + o);
            for (int d = 0; d < C::OBS; ++d) acc = fmaf(d < C::OBSNZ ? obs_s[d] : 0.f, __ldg(P.w1 + d * P.h1 + o), acc);
            h1v[o] = fmaxf(acc, 0.f);
        }
        __syncwarp();
        for (int o = gl; o < P.h2; o += C::LPE) {
            float acc = __ldg(P.b2 + o);
            for (int d = 0; d < P.h1; ++d) acc = fmaf(h1v[d], __ldg(P.w2 + d * P.h2 + o), acc);
            h2v[o] = fmaxf(acc, 0.f);
        }
        __syncwarp();
        for (int o = gl; o < C::NACT; o += C::LPE) {
            float acc = __ldg(P.b3 + o);
            for (int d = 0; d < P.h2; ++d) acc = fmaf(h2v[d], __ldg(P.w3 + d * C::NACT + o), acc);
            act_out[o] = acc;            // raw policy output: the torque clips it, the electricity cost does not (quirk Q3)
        }
        __syncwarp();
    }

    // `pred` guards every persistent write, so that a whole warp can run the task / reset code
    // while only some of its env groups need it (no divergent __syncwarp / shuffles).
    __device__ bool task(const float *act, float *obs_out, float *rew_out, float *terms_out, bool reset_pass,
                         bool pred = true) {
        if (m->kind <= 1 || m->kind == 9 || m->kind == 11) return pendulum_task(obs_out, rew_out, terms_out, reset_pass, pred);
        if (m->kind == 10) return reacher_task(act, obs_out, rew_out, terms_out, reset_pass, pred);
        if (m->kind == 12 || m->kind == 13 || m->kind == 16) return mjwalker_task(act, obs_out, rew_out, terms_out, reset_pass, pred);
        return walker_task(act, obs_out, rew_out, terms_out, reset_pass, pred);
    }

    // Episode reset: MJCF pose, joint noise (rs/robot_locomotors.py:16-24), zero velocities,
    // cleared feet flags and warm-start cache.  noise == nullptr draws from the counter RNG.
    __device__ void reset_state(const LaunchArgs &la, unsigned long long env, const float *noise, int floor_in_parts,
                                bool pred = true) {
        float *S = st();
        float *T = S + C::oT;
        const unsigned ep = (unsigned)__float_as_int(T[T_EPISODE]) + 1u;
        __syncwarp();
        if (pred) {
            for (int i = gl; i < C::oT; i += C::LPE) S[i] = 0.f;
            for (int i = gl; i < 2 * C::NFEET; i += C::LPE) S[C::oF + i] = 0.f;
        }
        __syncwarp();
        if (pred) {
            if (C::FLOATING) {
                if (gl < 3) S[gl] = m->base_pos0[gl];
                if (gl < 4) S[3 + gl] = m->base_quat0[gl];
            }
            if (C::HASX) {
                // restoreState + resetBasePositionAndOrientation(cube, [-1.5, 0, 0.05], identity) (rs/robot_locomotors.py:240-243)
                if (gl < 3) S[C::oX + gl] = m->cube_pos0[gl];
                if (gl == 3) S[C::oX + 6] = 1.f;
            }
            if (gl < C::NNOISE) {
                // draw ranges: U(-0.1, 0.1) per joint; Reacher: target_x/y U(+-0.27), joint0/1 U(+-3.14) (rs/robot_manipulators.py:12-21)
                const float rr = m->kind == 10 ? (gl < 2 ? 0.27f : 3.14f) : 0.1f;
                const float nz = noise ? noise[gl] : rng_uniform(la.seed, env, ep, 0u, (unsigned)gl, -rr, rr);
                if (m->kind <= 1) { if (gl == 0) S[C::oQ + 1] = nz + (m->kind == 1 ? 3.1415f : 0.f); }
                else if (m->kind == 9 || m->kind == 11) S[C::oQ + 1 + gl] = nz;   // hinge, hinge2 (rs/robot_pendula.py:66-68)
                else if (m->kind == 10) S[C::oQ + (gl ^ 2)] = nz;        // draws: target_x, target_y, joint0, joint1 -> dofs 2, 3, 0, 1
                else if (m->kind == 12 || m->kind == 13 || m->kind == 16) S[C::oQ + gl] = nz;   // every ordered joint incl. the root joints
                else S[C::oQ + m->act_joint[gl]] = nz;
            }
            if (gl == 0) {
                T[T_EPISODE] = __int_as_float((int)ep);
                T[T_STEPS] = __int_as_float(0);
                T[T_RETURN] = 0.f;
                T[T_FLOOR] = floor_in_parts ? 1.f : 0.f;
                T[T_TX] = m->walk_tx; T[T_TY] = m->walk_ty;
                T[T_HAVEZ] = 0.f;
                T[T_FLAGCNT] = __int_as_float(0); T[T_FLAGTIMEOUT] = 0.f;
                if (m->kind == 7 || m->kind == 8) {
                    // robot_specific_reset -> flag_reposition() (rs/robot_locomotors.py:200-218)
                    T[T_TX] = 0.5f * rng_uniform(la.seed, env, ep, 1u, 0u, -m->halflen, m->halflen);
                    T[T_TY] = 0.5f * rng_uniform(la.seed, env, ep, 1u, 1u, -m->halfwidth, m->halfwidth);
                    T[T_FLAGCNT] = __int_as_float(1);
                    T[T_FLAGTIMEOUT] = 600.f / (float)m->nsub;
                }
                T[T_FRAME] = __int_as_float(0); T[T_ONGROUND] = __int_as_float(0); T[T_CRAWL_HAS] = 0.f;
                T[T_CRAWL_START_LO] = T[T_CRAWL_START_HI] = T[T_CRAWL_IGN_LO] = T[T_CRAWL_IGN_HI] = 0.f;
                T[T_ATTACKS] = __int_as_float(0);
            }
        }
        __syncwarp();
    }
};

// ---------------------------------------------------------------------------------------------
// Fused policy on the tensor cores (pbg_set_policy_tensor_cores).  The scalar MLP in Env::policy_actions evaluates one env per
// lane group and streams the whole weight set through every warp: 15 % (Ant, 28-128-64-8) to 20 % (Humanoid, 44-256-128-17) of a
// fused step.  Here the CTA's envs form the rows of one small GEMM per layer, Y[EPB x N] = act(X[EPB x K] W[K x N] + b):
// work items are 16 x 8 output tiles (mma.sync.m16n8k8, TF32 inputs rounded to nearest, FP32 accumulation) dealt round-robin to
// the warps, A fragments come from the env blocks in shared memory (row = env slot, stride ENV_FLOATS), B fragments straight
// from the row-major weights in global memory (each weight is read once per CTA instead of once per warp).  TF32 keeps 10
// mantissa bits of the inputs: actions differ from the FP32 path by ~1e-3, which is why this is opt-in and the FP32 path stays
// the one that makes K fused steps bit-identical to K single steps with a torch FP32 policy.
__device__ __forceinline__ unsigned f2tf32(float x) {
    unsigned r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void mma_tf32_16x8x8(float (&c)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
// one layer for all env slots of the CTA; xin / yout point at slot 0's vectors, slot r's are r * ENV_FLOATS further
template <class C>
__device__ __forceinline__ void mlp_layer_tc(const float *__restrict__ W, const float *__restrict__ bias, const int K, const int N,
                                             const bool relu, const float *xin, float *yout) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    constexpr int MT = (C::EPB + 15) / 16;
    const int NT = (N + 7) >> 3;
    for (int item = warp; item < MT * NT; item += C::WARPS) {
        const int mt = item % MT, nt = item / MT;
        const int r0 = mt * 16 + g, r1 = r0 + 8;
        const bool v0 = r0 < C::EPB, v1 = r1 < C::EPB;
        const float *x0 = xin + (v0 ? r0 : 0) * C::ENV_FLOATS, *x1 = xin + (v1 ? r1 : 0) * C::ENV_FLOATS;
        const int n = nt * 8 + g;
        const bool nv = n < N;
        const float *wn = W + (nv ? n : 0);
        float c[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
        for (int k0 = 0; k0 < K; k0 += 8) {
            const int ka = k0 + t, kb = ka + 4;
            const bool ia = ka < K, ib = kb < K;
            unsigned a[4], b[2];
            a[0] = f2tf32(v0 && ia ? x0[ka] : 0.f); a[1] = f2tf32(v1 && ia ? x1[ka] : 0.f);
            a[2] = f2tf32(v0 && ib ? x0[kb] : 0.f); a[3] = f2tf32(v1 && ib ? x1[kb] : 0.f);
            b[0] = f2tf32(nv && ia ? __ldg(wn + ka * N) : 0.f); b[1] = f2tf32(nv && ib ? __ldg(wn + kb * N) : 0.f);
            mma_tf32_16x8x8(c, a, b);
        }
        float *y0 = yout + (v0 ? r0 : 0) * C::ENV_FLOATS, *y1 = yout + (v1 ? r1 : 0) * C::ENV_FLOATS;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int col = nt * 8 + 2 * t + j;
            if (col < N) {
                const float bj = __ldg(bias + col);
                float u0 = c[j] + bj, u1 = c[2 + j] + bj;
                if (relu) { u0 = fmaxf(u0, 0.f); u1 = fmaxf(u1, 0.f); }
                if (v0) y0[col] = u0;
                if (v1) y1[col] = u1;
            }
        }
    }
}
// observation (staged in every slot's sOUT) -> actions (every slot's sACTN); every thread of the CTA takes part
template <class C>
__device__ void policy_actions_tc(const PolicyDev &P, float *smem) {
    __syncthreads();                                     // every warp's observation is staged
    mlp_layer_tc<C>(P.w1, P.b1, C::OBSNZ, P.h1, true, smem + C::sOUT, smem + C::sU);
    __syncthreads();
    mlp_layer_tc<C>(P.w2, P.b2, P.h1, P.h2, true, smem + C::sU, smem + C::sU + P.h1);
    __syncthreads();
    mlp_layer_tc<C>(P.w3, P.b3, P.h2, C::NACT, false, smem + C::sU + P.h1, smem + C::sACTN);
    __syncthreads();
}

// POLICY = true is the multi-step fused-policy instantiation (MODE_POLICY only); keeping it a separate kernel leaves the
// single-step kernel's register allocation untouched.
template <class C, bool POLICY = false>
__global__ void __launch_bounds__(C::THREADS, C::MIN_BLOCKS) env_kernel(const DevModel *__restrict__ model, StepBuffers B, LaunchArgs la) {
    extern __shared__ __align__(16) float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    Env<C> e;
    PBG_PHASE_BEGIN;
#ifdef PBG_PHASE_CLOCKS
    const long long _kstart = clock64();
#endif
    e.gl = lane & (C::LPE - 1);
    e.grp = lane / C::LPE;
    const int slot = warp * C::EPW + e.grp;
    e.sm = smem + slot * C::ENV_FLOATS;
    const long long env_raw = (long long)blockIdx.x * C::EPB + slot;
    const bool valid = env_raw < la.E;
    const long long env = valid ? env_raw : (la.E - 1);
    const unsigned long long genv = la.env_offset + (unsigned long long)env;
    const int gl = e.gl;
    // The env's state and actions are requested first (coalesced, vectorised; the actions may live in mapped pinned host memory --
    // pbg_step_host's zero-copy path -- and are read exactly once), so that their latency runs under the model-table copy below.
    float *gs = B.state + env * C::SSTRIDE;
    constexpr int NLD = (C::SSTRIDE / 4 + C::LPE - 1) / C::LPE;
    float4 sreg[NLD];
#pragma unroll
    for (int r = 0; r < NLD; ++r) {
        const int i = r * C::LPE + gl;
        sreg[r] = i < C::SSTRIDE / 4 ? reinterpret_cast<const float4 *>(gs)[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const float *gact = B.actions ? B.actions + env * C::NACT : nullptr;
    const float areg = (gl < C::NACT && gact) ? gact[gl] : 0.f;
#ifndef PBG_NO_SMEM_MODEL
    if (C::MODEL_IN_SMEM) {
        static_assert(sizeof(DevModel) % 16 == 0 && C::ENV_BYTES % 16 == 0, "model copy is done in 16-byte words");
        float4 *dst = reinterpret_cast<float4 *>(reinterpret_cast<char *>(smem) + C::ENV_BYTES);
        const float4 *src = reinterpret_cast<const float4 *>(model);
        for (int i = threadIdx.x; i < int(sizeof(DevModel) / 16); i += C::THREADS) dst[i] = src[i];
        __syncthreads();
        model = reinterpret_cast<const DevModel *>(dst);
    }
#endif
    e.m = model;
    e.load_lane_constants();
    e.rng_seed = la.seed; e.rng_env = genv;
    if (B.debug && env_raw == la.debug_env) e.dbg = B.debug;

    float *S = e.st();
#pragma unroll
    for (int r = 0; r < NLD; ++r) {
        const int i = r * C::LPE + gl;
        if (i < C::SSTRIDE / 4) reinterpret_cast<float4 *>(S)[i] = sreg[r];
    }
    if (gl < C::NACT) e.sm[C::sACTN + gl] = areg;
    __syncwarp();
    float *T = S + C::oT;
    const int mode = la.mode;
    PBG_PHASE(12);

    if (mode == MODE_GET) {
        float *cs = B.canon + env * C::CANON;
        if (valid) {
            if (C::FLOATING) {
                if (gl < 7) cs[gl] = S[gl];
                if (gl < 6) cs[7 + gl] = S[C::oU + gl];
            }
            for (int j = gl; j < C::NJ; j += C::LPE) {
                cs[13 * C::FLOATING + j] = S[C::oQ + j];
                cs[13 * C::FLOATING + C::NJ + j] = S[C::oU + 6 * C::FLOATING + j];
            }
            if (C::HASX) {
                float *cx = cs + 13 * C::FLOATING + 2 * C::NJ;
                if (gl < 7) cx[gl] = S[C::oX + gl];
                if (gl < 6) cx[7 + gl] = S[C::oU + C::XD0 + gl];
            }
        }
        return;
    }
    if (mode == MODE_SET) {
        const float *cs = B.canon + env * C::CANON;
        if (C::FLOATING) {
            if (gl < 7) S[gl] = cs[gl];
            if (gl < 6) S[C::oU + gl] = cs[7 + gl];
        }
        for (int j = gl; j < C::NJ; j += C::LPE) {
            S[C::oQ + j] = cs[13 * C::FLOATING + j];
            S[C::oU + 6 * C::FLOATING + j] = cs[13 * C::FLOATING + C::NJ + j];
        }
        if (C::HASX) {
            const float *cx = cs + 13 * C::FLOATING + 2 * C::NJ;
            if (gl < 7) S[C::oX + gl] = cx[gl];
            if (gl < 6) S[C::oU + C::XD0 + gl] = cx[7 + gl];
        }
        for (int i = gl; i < C::NSLOT; i += C::LPE) S[C::oW + i] = 0.f;
        __syncwarp();
        if (valid) for (int i = gl; i < C::SSTRIDE / 4; i += C::LPE)
            reinterpret_cast<float4 *>(gs)[i] = reinterpret_cast<const float4 *>(S)[i];
        return;
    }

    const float *act = gact ? e.sm + C::sACTN : nullptr;
    float *obs = B.obs ? B.obs + env * C::OBS : nullptr;
    // outputs are staged in shared memory, then written by consecutive lanes
    float *so = e.sm + C::sOUT;
    float *so_obs = so, *so_rew = so + 64, *so_terms = so + 65;

    // MODE_POLICY: `nsteps` env steps in this one launch, actions from the fused MLP on the staged observation;
    // the state stays in shared memory in between.  Everything else is exactly MODE_STEP.
    const bool policy_mode = POLICY;
    const int mode_eff = policy_mode ? (int)MODE_STEP : mode;
    const int nsteps = policy_mode ? la.nsteps : 1;
    float ret_acc = 0.f;
    bool any_done = false;
    bool store_state = mode_eff == MODE_STEP || mode == MODE_OBSERVE;
    if (policy_mode) {
        for (int i = gl; i < C::OBSNZ; i += C::LPE) so_obs[i] = obs[i];      // the observation the caller holds for this state
        act = e.sm + C::sACTN;
        __syncwarp();
    }
  for (int step = 0; step < nsteps; ++step) {
    const bool last_step = step == nsteps - 1;
    // hidden activations go to the kinematics / Delassus scratch, which is dead between the task phase and the next sub-step
    if (policy_mode) {
        if (B.policy.tc) policy_actions_tc<C>(B.policy, smem);
        else e.policy_actions(B.policy, so_obs, e.sm + C::sU, e.sm + C::sACTN);
    }
    if (mode_eff == MODE_STEP || mode == MODE_PHYSICS) {
        // apply_action: tau = power * power_coef * clip(a, -1, 1) (rs/robot_locomotors.py:26-29),
        // plus the joint damping torque, both held for all substeps
        float tq = 0.f;
        if (gl >= 6 * C::FLOATING && gl < C::XD0) {
            const int j = gl - 6 * C::FLOATING;
            tq = -model->jdamp[j] * S[C::oU + gl];
            const int ai = model->jact[j];
            // a non-finite action (the reference asserts on it, rs/robot_locomotors.py:27) must not be laundered into a
            // finite torque by fminf / fmaxf: it poisons this env's state, which ends its episode as a non-finite observation
            if (ai >= 0) { const float av = act[ai]; tq += model->jtorque[j] * (isfinite(av) ? fminf(fmaxf(av, -1.f), 1.f) : CUDART_NAN_F); }
        }
        e.tau = tq;
        e.ovf = 0; e.dbg_nov = -100;
        const int nsub = model->nsub;
        for (int s = 0; s < nsub; ++s) {
#ifndef PBG_SYNC_MODE
#define PBG_SYNC_MODE 1
#endif
            // re-align the CTA's warps (instruction-cache locality); mode 1: every substep, 2: every other, 0: never
            { PBG_PHASE_BEGIN; if (C::WARPS > 2 && (PBG_SYNC_MODE == 1 || (PBG_SYNC_MODE == 2 && (s & 1) == 0))) __syncthreads(); PBG_PHASE(0); }
            e.substep(s == nsub - 1);
        }
    }
    // optional: distance of every contact candidate slot in the step's last collision pass (+inf: not in contact);
    // the shell's BodyPart.contact_list() (rs/robot_bases.py:280-281) is built from it
    if (C::NSLOT > 0 && valid && B.cand_out && (mode_eff == MODE_STEP || mode == MODE_PHYSICS))
        for (int i = gl; i < C::NSLOT; i += C::LPE) B.cand_out[env * C::NSLOT + i] = e.sm[C::sCD + i];
    if (mode == MODE_PHYSICS) {
        if (C::MAXC > 0 && gl < C::NFEET) S[C::oP + gl] = e.sm[C::sMISC + gl];
        __syncwarp();
        if (valid) {
            for (int i = gl; i < C::SSTRIDE / 4; i += C::LPE)
                reinterpret_cast<float4 *>(gs)[i] = reinterpret_cast<const float4 *>(S)[i];
            if (B.ncontact_out && gl == 0) B.ncontact_out[env] = e.nc;
        }
        return;
    }

    // ---- task layer, two passes through ONE task() call site (keeps the code small):
    //   pass 0  calc_state / reward / termination of the stepped state        (STEP, OBSERVE)
    //   pass 1  episode reset + first observation, predicated per env group   (RESET; STEP when finished)
    // calc_state must still see the previous step's feet flags (quirk Q2: the reference updates
    // robot.feet_contact after calc_state / alive_bonus); this step's flags were staged by the last
    // substep's collide() and replace them after pass 0.
    PBG_PHASE_RESET;
    if (mode == MODE_OBSERVE) e.nc = 0;
    const bool reset_mode = mode == MODE_RESET;
    const bool mask_on = reset_mode && (!B.mask || B.mask[env]);
    bool want_reset = mask_on;
    for (int pass = reset_mode ? 1 : 0; pass < 2; ++pass) {
        bool rp = false, pred = true;
        if (pass == 1) {
            if (!reset_mode && !__any_sync(0xffffffffu, want_reset)) break;
            __syncwarp();
            e.reset_state(la, genv, (reset_mode && B.noise) ? B.noise + env * C::NNOISE : nullptr,
                          reset_mode ? la.floor_in_parts : 1, want_reset);
            rp = want_reset;                 // envs left out of a masked reset only get their observation refreshed
            pred = reset_mode ? true : want_reset;
            if (reset_mode) store_state = want_reset;
        }
        const bool done = e.task(act, so_obs, so_rew, so_terms, rp, pred);
        if (pass == 1) break;
        if (mode_eff == MODE_STEP && C::MAXC > 0 && gl < C::NFEET) S[C::oF + gl] = e.sm[C::sMISC + gl];
        if (mode == MODE_OBSERVE && C::MAXC > 0 && gl < C::NFEET) S[C::oF + gl] = S[C::oP + gl];
        __syncwarp();
        bool trunc = false;
        if (mode_eff == MODE_STEP) {
            if (gl == 0) {
                T[T_STEPS] = __int_as_float(__float_as_int(T[T_STEPS]) + 1);
                T[T_RETURN] += isfinite(so_rew[0]) ? so_rew[0] : 0.f;     // keep the episode statistics finite
            }
            __syncwarp();
            trunc = !done && __float_as_int(T[T_STEPS]) >= model->max_steps;
        }
        const bool finished = done || trunc;
        if (mode_eff == MODE_STEP && e.ovf && valid && gl == 0 && B.stats) atomicAdd(&B.stats[6], 1ull);
        if (policy_mode) { ret_acc += so_rew[0]; any_done = any_done || finished; }
        if (valid && gl == 0 && !policy_mode) {
            if (B.reward) B.reward[env] = so_rew[0];
            if (B.done) B.done[env] = (done || (trunc && la.auto_reset)) ? 1 : 0;
            if (B.truncated) B.truncated[env] = trunc ? 1 : 0;
            if (B.ncontact_out) B.ncontact_out[env] = e.nc;
        }
        if (valid && B.terms) { if (gl < 5) B.terms[env * 5 + gl] = so_terms[gl]; }
        if (valid && B.feet_out) { if (gl < C::NFEET) B.feet_out[env * C::NFEET + gl] = S[C::oF + gl]; }
        if (mode_eff == MODE_STEP && finished) {
            if (valid && gl == 0 && B.stats) {
                atomicAdd(&B.stats[0], 1ull);
                atomicAdd(&B.stats[1], (unsigned long long)__float_as_int(T[T_STEPS]));
                atomicAdd(reinterpret_cast<double *>(&B.stats[2]), (double)T[T_RETURN]);
                if (trunc) atomicAdd(&B.stats[3], 1ull);
                if (T[T_HAVEZ] == 2.f) atomicAdd(&B.stats[4], 1ull);
            }
            if (la.auto_reset && valid && B.final_obs)
                for (int i = gl; i < C::OBS; i += C::LPE) B.final_obs[env * C::OBS + i] = i < C::OBSNZ ? so_obs[i] : 0.f;
        }
        want_reset = mode_eff == MODE_STEP && finished && la.auto_reset;
    }
    __syncwarp();
    (void)last_step;
    PBG_PHASE(11);
  }
    if (threadIdx.x == 0 && B.stats && mode_eff == MODE_STEP) {
        // env steps taken, counted on the device so that CUDA-graph replays of pbg_step count too
        const long long nv = (long long)la.E - (long long)blockIdx.x * C::EPB;
        if (nv > 0) atomicAdd(&B.stats[5], (unsigned long long)(nv < C::EPB ? nv : C::EPB) * (unsigned long long)nsteps);
    }
    if (policy_mode && valid && gl == 0) {
        if (B.reward) B.reward[env] = ret_acc;            // sum of the rewards of the nsteps steps
        if (B.done) B.done[env] = any_done ? 1 : 0;       // an episode ended (and restarted) during the rollout
    }
    PBG_PHASE_RESET;
    if (valid) {
        if (obs) for (int i = gl; i < C::OBS; i += C::LPE) obs[i] = i < C::OBSNZ ? so_obs[i] : 0.f;
        if (store_state)
            for (int i = gl; i < C::SSTRIDE / 4; i += C::LPE)
                reinterpret_cast<float4 *>(gs)[i] = reinterpret_cast<const float4 *>(S)[i];
    }
    PBG_PHASE(13);
#ifdef PBG_PHASE_CLOCKS
    {
        const int wid = blockIdx.x * C::WARPS + warp;
        unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        if ((threadIdx.x & 31) == 0 && wid < 4096) { g_warpclk[3 * wid] = (unsigned)(clock64() - _kstart); g_warpclk[3 * wid + 1] = smid; g_warpclk[3 * wid + 2] = (unsigned)(e.dbg_nov + 100); }
    }
#endif
}

}  // namespace pbg
