// Device-side articulation tables and per-env state layout of libpbg_b200.
//
// The tables are the flattened ReducedModel produced by pybullet_gym_b200/mjcf/compiler.py; they
// replace what the reference reads back from pybullet after loadMJCF
// (/root/reference/pybulletgym/envs/roboschool/robot_bases.py:54-89).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pbg {

constexpr int MB = 20;      // max bodies of the reduced tree (humanoid: 18)
constexpr int MJ = 24;      // max joint dofs
constexpr int MSUB = 32;    // max Bullet links folded into the bodies (humanoid: 30)
constexpr int MCAND = 40;   // max ground contact candidates (humanoid: 4 spheres + 13 capsules * 2, + 8 cube corners)
constexpr int MPAIR = 84;   // max geom pairs (humanoid: 66 self-collision pairs + 17 geoms against the cube)
constexpr int MGEOM = 24;   // max collision geoms that take part in geom pairs (humanoid: 17)
constexpr int MFEET = 8;
constexpr int TASK_FLOATS = 24;

struct alignas(16) DevModel {
    int nb, nj, nd, floating, ncand, npair, nact, nfeet, obs_dim, kind, maxdepth, torso_body, nlim;
    int parent[MB], jtype[MB], depth[MB], dof[MB];
    unsigned up[MJ + 6];        // per dof: dofs that move the dof's body (ancestors or self)
    unsigned down[MJ + 6];      // per dof: dofs whose body this dof moves (descendants or self)
    unsigned anc[MB];           // per body: dofs that move the body
    float q0m[MB][9], anchor_p[MB][3], com_off[MB][3], axis[MB][3], mass[MB], inertia[MB][6];
    float axis_len[MB];         // axis[] holds the unit direction (rotation), axis_len the length of the MJCF vector (motion subspace)
    float part_cnt[MB], part_sum[MB][3];      // robot.parts entries folded into the body: count, sum of offsets
    int ds_begin[MB + 1];                     // per-link damping entries, grouped by body
    float ds_off[MSUB][3], ds_mass[MSUB], ds_inertia[MSUB][3];
    int jbody[MJ], jrev[MJ], jlimited[MJ], jact[MJ];
    float jlo[MJ], jhi[MJ], jdamp[MJ], jtorque[MJ];
    int act_joint[MJ];
    int c_body[MCAND], c_foot[MCAND];
    float c_p[MCAND][3], c_rad[MCAND], c_thr[MCAND], c_mu[MCAND];
    float c_spin[MCAND], c_roll[MCAND];      // combined spinning / rolling friction of a ground candidate (torsional rows)
    int p_ba[MPAIR], p_bb[MPAIR];
    float p_ra[MPAIR], p_rb[MPAIR], p_thr[MPAIR], p_mu[MPAIR];
    float p_half[MPAIR];                      // pairs against the cube: its half extent
    int p_box[MPAIR];                         // 1: body B is the cube
    // the geoms behind the pairs: their end points are taken to world coordinates once per collision pass (every humanoid geom
    // is in ~8 pairs) and the pairs read them from shared memory
    int ng, g_body[MGEOM], p_ga[MPAIR], p_gb[MPAIR];      // p_gb = -1: body B is the cube
    float g_p0[MGEOM][3], g_p1[MGEOM][3];
    float cube_pos0[3];
    int aux_body[2];                          // Reacher: fingertip / target bodies and their link-COM offsets
    float aux_off[2][3];
    float torso_off[3];
    float base_pos0[3], base_quat0[4];
    // scene
    float gravity, h, erp_contact, erp_limit, slop, warm, kdamp, maxvel, limit_max_imp, split_thr;
    int nsub, niter, limit_split, max_contacts;
    // task
    float initial_z, elec_cost, stall_cost, limit_cost, walk_tx, walk_ty, halflen, halfwidth;
    double dt_scene, inv_dt_scene;
    int max_steps;
};

// task block of the per-env state (float slots)
enum {
    T_POT_LO = 0, T_POT_HI = 1, T_INITZ = 2, T_STEPS = 3, T_EPISODE = 4, T_RETURN = 5, T_FLOOR = 6, T_TX = 7, T_TY = 8,
    T_FLAGTIMEOUT = 9, T_HAVEZ = 10, T_FLAGCNT = 11,
    // HumanoidFlagrunHarder (rs/robot_locomotors.py:230-302)
    T_FRAME = 12, T_ONGROUND = 13, T_CRAWL_HAS = 14, T_CRAWL_START_LO = 15, T_CRAWL_START_HI = 16, T_CRAWL_IGN_LO = 17,
    T_CRAWL_IGN_HI = 18, T_ATTACKS = 19
};

// Two-hidden-layer ReLU policy (the shape of the reference's pretrained agents,
// /root/reference/pybulletgym/examples/roboschool-weights/enjoy_TF_*.py): a = W3^T relu(W2^T relu(W1^T obs + b1) + b2) + b3,
// weights row-major [in, out] in device memory
struct PolicyDev {
    const float *w1, *b1, *w2, *b2, *w3, *b3;
    int h1, h2;
    int tc;     // 1: the CTA's envs go through the tensor cores together (mma.sync, TF32 inputs, FP32 accumulation)
};

struct StepBuffers {
    float *state;             // [E, SSTRIDE]
    const float *actions;     // [E, nact]
    const float *noise;       // [E, nact] optional injected reset noise
    const uint8_t *mask;      // [E] optional reset mask
    float *obs;               // [E, obs_dim]
    float *reward;            // [E]
    uint8_t *done;            // [E]
    float *terms;             // [E,5] optional
    float *final_obs;         // [E, obs_dim] optional
    uint8_t *truncated;       // [E] optional
    float *feet_out;          // [E, nfeet] optional
    int *ncontact_out;        // [E] optional
    float *cand_out;          // [E, NSLOT] optional: per contact-candidate slot, distance in the last collision pass (+inf: inactive)
    unsigned long long *stats;  // [8] device episode statistics
    float *canon;             // [E, state_dim] for get/set state
    float *debug;             // development: constraint-row dump of env `debug_env`
    PolicyDev policy;         // MODE_POLICY
};

enum { MODE_STEP = 0, MODE_PHYSICS = 1, MODE_OBSERVE = 2, MODE_RESET = 3, MODE_GET = 4, MODE_SET = 5, MODE_POLICY = 6 };

struct LaunchArgs {
    int E;
    int mode;
    int auto_reset;
    int floor_in_parts;
    unsigned long long seed, env_offset;
    int debug_env;
    int nsteps;               // MODE_POLICY: env steps per launch
};

}  // namespace pbg
