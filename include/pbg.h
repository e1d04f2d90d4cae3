/*
 * pbg.h -- C ABI of libpbg_b200.so, the B200-native batched physics backend.
 *
 * This is the drop-in boundary of the hot path.  In the reference the path sits behind ~30
 * pybullet client calls made from Python once per joint / link / foot and env step
 * (/root/reference/pybulletgym/envs/roboschool/, "rs/" below).  Here one call
 * steps a whole batch of environments on one GPU.  Plain pointers and sizes only; no torch types.
 *
 * Conventions
 *   - return 0 on success, a negative pbg_status otherwise; pbg_last_error() gives the message;
 *     nothing throws across the ABI.
 *   - all *_dev pointers are device pointers on the handle's device, row-major [num_envs, dim],
 *     fp32 (done: uint8).  The caller owns every I/O buffer; the library owns its internal state.
 *   - `stream` is a cudaStream_t passed as void* (0 = default stream).  Calls are asynchronous and
 *     ordered by the stream, except the *_host entry points, which synchronise before returning.
 *   - a handle is bound to one device and is not thread-safe.  Multi-GPU = one handle per device.
 *     Every call makes the handle's device the calling thread's current CUDA device (cudaSetDevice) and leaves it so.
 *   - a handle remembers the stream of its last stream-ordered call; a call on a different stream (and the *_host / pbg_stats
 *     entry points, which use a private stream / block) first waits for the work enqueued on that previous stream, so e.g.
 *     pbg_reset on one stream followed by pbg_step_host needs no synchronisation by the caller.  I/O buffers the CALLER writes
 *     on yet another stream are the caller's to order.
 *   - stepping (pbg_step*, pbg_physics_step*, pbg_observe, pbg_rollout_policy) before the first pbg_reset* / pbg_set_state /
 *     pbg_restore returns PBG_ERR_INVALID.
 *   - the stream-ordered entry points (pbg_reset*, pbg_step, pbg_physics_step*, pbg_observe, pbg_rollout_policy,
 *     pbg_get_state / pbg_set_state, pbg_get_feet_contact) enqueue kernels only -- no allocation, no synchronisation, all
 *     counters on the device -- and may be recorded into a CUDA graph (cudaStreamBeginCapture) and replayed.
 */
#ifndef PBG_H
#define PBG_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PBG_VERSION 103

typedef enum {
    PBG_OK = 0,
    PBG_ERR_INVALID = -1,      /* bad argument / model does not fit the compiled kernel configuration */
    PBG_ERR_CUDA = -2,         /* a CUDA runtime call failed */
    PBG_ERR_UNSUPPORTED = -3   /* env kind without a kernel */
} pbg_status;

/* env kinds: one per reference robot/env class pair (rs/gym_locomotion_envs.py:122-173,
 * rs/gym_pendulum_envs.py:7-49) */
enum {
    PBG_KIND_PENDULUM = 0, PBG_KIND_PENDULUM_SWINGUP = 1, PBG_KIND_HOPPER = 2, PBG_KIND_WALKER2D = 3,
    PBG_KIND_HALFCHEETAH = 4, PBG_KIND_ANT = 5, PBG_KIND_HUMANOID = 6, PBG_KIND_FLAGRUN = 7,
    PBG_KIND_FLAGRUN_HARDER = 8, PBG_KIND_DOUBLE_PENDULUM = 9, PBG_KIND_REACHER = 10,
    PBG_KIND_DOUBLE_PENDULUM_MJ = 11,  /* pybulletgym/envs/mujoco/gym_pendulum_envs.py:40-75 */
    PBG_KIND_HOPPER_MJ = 12, PBG_KIND_WALKER2D_MJ = 13,  /* pybulletgym/envs/mujoco/gym_locomotion_envs.py:121-206 */
    PBG_KIND_ANT_MJ = 14, PBG_KIND_HUMANOID_MJ = 15,     /* pybulletgym/envs/mujoco/gym_locomotion_envs.py:246-260 */
    PBG_KIND_HALFCHEETAH_MJ = 16                         /* pybulletgym/envs/mujoco/gym_locomotion_envs.py:211-244 */
};

enum { PBG_JT_FIXED = 0, PBG_JT_REVOLUTE = 1, PBG_JT_PRISMATIC = 2, PBG_JT_FREE = 3 };
enum { PBG_G_SPHERE = 0, PBG_G_CAPSULE = 1, PBG_G_BOX = 2 };

/*
 * Flat articulation tables produced by the host-side MJCF compiler
 * (pybullet_gym_b200/mjcf/compiler.py: ReducedModel).  Replaces what the reference obtains from
 * loadMJCF + getNumJoints/getJointInfo/getBodyInfo (rs/robot_bases.py:54-89,116) and the scene
 * parameters it sets through setGravity / setDefaultContactERP / setPhysicsEngineParameter /
 * changeDynamics (rs/scene_bases.py:70-73, rs/scene_stadium.py:33).
 * All arrays are copied by pbg_create; the caller may free them afterwards.
 */
typedef struct pbg_model {
    /* reduced dynamics tree, bodies in depth-first order */
    int32_t nb, nj, floating;
    const int32_t *parent;        /* [nb] */
    const int32_t *jtype;         /* [nb] PBG_JT_* of the joint connecting the body to its parent */
    const double *q0;             /* [nb*4] parent frame -> body frame at q = 0, (x,y,z,w) */
    const double *anchor_p;       /* [nb*3] joint anchor in the parent frame, from the parent COM (root: world pose) */
    const double *com_off;        /* [nb*3] body COM relative to the anchor, body frame */
    const double *axis;           /* [nb*3] joint axis, body frame */
    const double *mass;           /* [nb] */
    const double *inertia;        /* [nb*6] xx yy zz xy xz yz about the COM, body frame */
    /* joints (one per non-root body, same order) */
    const double *jnt_lower, *jnt_upper;     /* [nj]; lower > upper: no limit */
    const double *jnt_damping;               /* [nj] */
    const int32_t *jnt_act;                  /* [nj] action index driving the joint, -1: none */
    const double *jnt_torque;                /* [nj] power * power_coef (rs/robot_locomotors.py:29) */
    /* Bullet links folded into the bodies: robot.parts membership and per-link damping */
    int32_t ns;
    const int32_t *sub_body;      /* [ns] */
    const double *sub_off;        /* [ns*3] link COM relative to the body COM, body frame */
    const double *sub_mass;       /* [ns] */
    const double *sub_inertia;    /* [ns*3] */
    const int32_t *sub_in_parts;  /* [ns] */
    int32_t torso_sub;            /* robot_body: index into sub_* */
    /* collision geometry */
    int32_t ng;
    const int32_t *geom_body, *geom_type, *geom_ground, *geom_foot;   /* [ng]; geom_foot: index into foot_list or -1 */
    const double *geom_radius, *geom_p0, *geom_p1, *geom_friction, *geom_threshold;
    int32_t npair;
    const int32_t *pair_a, *pair_b;
    /* scene */
    double gravity, timestep;
    int32_t frame_skip, num_solver_iterations;
    double contact_erp, erp, linear_slop, warmstarting_factor, link_damping, max_coordinate_velocity;
    double ground_friction, limit_max_impulse, split_impulse_threshold;
    int32_t limit_split_impulse, max_contacts;
    /* task (rs/robot_locomotors.py, rs/gym_locomotion_envs.py:48-52) */
    int32_t kind, action_dim, obs_dim, nfeet, max_episode_steps;
    double initial_z;             /* < 0: latch torso z at the first calc_state after reset */
    double electricity_cost, stall_torque_cost, joints_at_limit_cost;
    double walk_target_x, walk_target_y;
    double stadium_halflen, stadium_halfwidth;
    /* HumanoidFlagrunHarder's aggressive cube (rs/robot_locomotors.py:236-266, gym_utils.py:9-16,
     * assets/things/cube_small.urdf): a free box thrown at the robot every 30 frames.  cube = 0: none.
     * Replaces loadURDF(cube_small.urdf) + changeDynamics(mass=1.2) + resetBasePositionAndOrientation /
     * resetBaseVelocity on the cube body. */
    int32_t cube;
    double cube_half, cube_mass, cube_inertia, cube_friction, cube_threshold;
    double cube_pos0[3];
    /* links (indices into sub_*) whose COM the task layer reads besides torso_sub: Reacher's fingertip and
     * target (rs/robot_manipulators.py:17-18,33); -1: unused */
    int32_t aux_sub[2];
    /* Torsional friction (changeDynamics(spinningFriction=, rollingFriction=), mujoco/robot_locomotors.py:207-210): when
     * torsional_friction != 0 every ground contact gets, between its normal row and its two lateral friction rows, one row
     * about the contact normal bounded by spin * lambda_n and two rows about the tangents bounded by roll * lambda_n, with
     * Bullet's combination rule spin = geom_spin * ground_friction + ground_spinning_friction * geom_friction (same for roll). */
    int32_t torsional_friction;
    const double *geom_spin, *geom_roll;      /* [ng] */
    double ground_spinning_friction, ground_rolling_friction;
} pbg_model;

typedef struct pbg_handle pbg_handle;

/* device-side episode statistics, accumulated since creation / the last pbg_stats(reset=1) */
typedef struct pbg_episode_stats {
    double return_sum;            /* sum of finished-episode returns */
    double length_sum;            /* sum of finished-episode lengths */
    int64_t episodes;             /* finished episodes (terminated or truncated) */
    int64_t truncated;            /* of which ended by max_episode_steps */
    int64_t nonfinite;            /* episodes ended by a non-finite observation (rs/gym_locomotion_envs.py:63-65) */
    int64_t steps;                /* env steps taken */
    int64_t contact_overflow;     /* env steps in which a sub-step had more contact candidates under their breaking threshold
                                     than the kernel's solver budget (pbg_max_contacts); the deepest were kept.  The feet flags
                                     are not affected by the budget. */
} pbg_episode_stats;

int pbg_version(void);

/* Creates `num_envs` independent worlds of one env kind on `device`.  Replaces BulletClient() +
 * scene/robot construction in BaseBulletEnv._reset (rs/env_bases.py:46-71).  env i of this handle
 * uses RNG stream (seed, env_offset + i), so shards of one job stay decorrelated across GPUs. */
int pbg_create(const pbg_model *model, int32_t num_envs, int32_t device, uint64_t seed, uint64_t env_offset,
               pbg_handle **out);
int pbg_destroy(pbg_handle *h);
const char *pbg_last_error(const pbg_handle *h);   /* h may be NULL: error of the last failed pbg_create */

/* sizes */
int pbg_num_envs(const pbg_handle *h);
int pbg_obs_dim(const pbg_handle *h);
int pbg_action_dim(const pbg_handle *h);
int pbg_noise_dim(const pbg_handle *h);     /* reset draws per env that pbg_reset_with injects */
int pbg_state_dim(const pbg_handle *h);     /* canonical state: [pos3 quat4(xyzw) omega3 vel3] (floating) + q[nj] + qd[nj]
                                               + [cube pos3 quat4 omega3 vel3] (worlds with the cube) */

/* Episode reset (rs/gym_locomotion_envs.py:22-39, rs/robot_locomotors.py:16-24): restores the
 * MJCF pose, draws U(-0.1,0.1) joint noise from the counter RNG, returns the first observation.
 * mask_dev: uint8[num_envs] or NULL (= all).  floor_in_parts: reference quirk Q1 -- 0 reproduces the
 * very first reset of an env's life (floor not yet in robot.parts), 1 every later reset. */
int pbg_reset(pbg_handle *h, const uint8_t *mask_dev, int32_t floor_in_parts, float *obs_dev, void *stream);
/* Same, with the reset draws given by the caller (float[num_envs, pbg_noise_dim]: one per actuated joint for
 * the walkers, the hinge for the pendulum, hinge + hinge2 for the double pendulum), for parity tests. */
int pbg_reset_with(pbg_handle *h, const float *joint_noise_dev, int32_t floor_in_parts, float *obs_dev, void *stream);

/* One env step for every env: apply_action + stepSimulation + calc_state + reward/termination
 * (rs/gym_locomotion_envs.py:54-114, rs/gym_pendulum_envs.py:26-39) fused in one kernel launch.
 * Envs that finish (done or max_episode_steps) are reset inside the kernel when auto-reset is on;
 * obs then holds the first observation of the new episode and final_obs_dev (optional) the last
 * observation of the finished one.  reward_terms_dev: optional float[num_envs,5] =
 * [alive, progress, electricity, joints_at_limit, feet_collision] (the reference's self.rewards).
 * truncated_dev: optional uint8[num_envs], 1 where the episode hit max_episode_steps. */
int pbg_step(pbg_handle *h, const float *actions_dev, float *obs_dev, float *reward_dev, uint8_t *done_dev,
             float *reward_terms_dev, float *final_obs_dev, uint8_t *truncated_dev, void *stream);

/* Fused policy rollouts (the reference's counterpart is the numpy MLP loop of
 * pybulletgym/examples/roboschool-weights/enjoy_TF_*.py: obs -> relu(W1) -> relu(W2) -> W3 -> env.step).
 * pbg_set_policy copies a two-hidden-layer ReLU policy (host pointers, row-major [in, out]: w1[obs_dim, h1], w2[h1, h2],
 * w3[h2, action_dim]) to the device.  pbg_rollout_policy then advances every env by `nsteps` env steps in ONE kernel launch:
 * observation -> MLP -> action -> step, state resident in shared memory, auto-reset as in pbg_step.  obs_dev is in/out: the
 * observation of the current state on entry (what pbg_reset / pbg_step returned), the last one on return;
 * reward_sum_dev[num_envs] (optional) receives the sum of the nsteps rewards, done_any_dev[num_envs] (optional) whether an
 * episode ended during the rollout. */
int pbg_set_policy(pbg_handle *h, int32_t h1, int32_t h2, const float *w1, const float *b1, const float *w2, const float *b2,
                   const float *w3, const float *b3);
int pbg_rollout_policy(pbg_handle *h, int32_t nsteps, float *obs_dev, float *reward_sum_dev, uint8_t *done_any_dev, void *stream);
/* Evaluates the fused policy on the tensor cores: the envs of a CTA form the rows of one small GEMM per layer (mma.sync
 * m16n8k8, TF32 inputs rounded to nearest, FP32 accumulation), every weight is read once per CTA instead of once per warp.
 * Opt-in because TF32 keeps 10 mantissa bits of the inputs: actions differ from the default FP32 evaluation by ~1e-3
 * (absolute, for O(1) activations), so a rollout is no longer bit-identical to single steps driven by an FP32 policy outside.
 * Takes effect at the next pbg_rollout_policy; 0 (default) = scalar FP32. */
int pbg_set_policy_tensor_cores(pbg_handle *h, int32_t enabled);

/* Host-buffer variant (the reference-facing call: numpy in, numpy out).  Copies actions H2D,
 * steps, copies obs/reward/done D2H and synchronises.  Buffers should be pinned for full speed. */
int pbg_step_host(pbg_handle *h, const float *actions_host, float *obs_host, float *reward_host, uint8_t *done_host);

/* pbg_step_host transport.  By default buffers that are pinned (cudaHostAlloc / cudaHostRegister; torch
 * pin_memory()) are read and written by the step kernel directly through their UVA mapping ("zero-copy": no
 * cudaMemcpy, one launch + one synchronise); pageable buffers, or enabled = 0, use staged H2D / D2H copies.
 * pbg_last_host_path: 1 = zero-copy, 2 = staged copies, 0 = pbg_step_host not called yet. */
int pbg_set_zero_copy(pbg_handle *h, int32_t enabled);
int pbg_last_host_path(const pbg_handle *h);

int pbg_set_auto_reset(pbg_handle *h, int32_t enabled);

/* Canonical state access for parity tests (replaces getJointState / getBasePositionAndOrientation /
 * getBaseVelocity / resetJointState / resetBasePositionAndOrientation / resetBaseVelocity,
 * rs/robot_bases.py:233-275,323-356).  float[num_envs, pbg_state_dim]. */
int pbg_get_state(pbg_handle *h, float *state_dev, void *stream);
int pbg_set_state(pbg_handle *h, const float *state_dev, void *stream);
/* Whole-handle snapshot: every env's physics state, warm-start impulses, task bookkeeping (potential, episode step and
 * episode counters that drive the reset RNG, targets, cube) and the episode statistics -- what pybullet's saveState /
 * restoreState (rs/gym_pendulum_envs.py:20-27) plus a pickle of the env object would hold.  pbg_restore into the same
 * handle, or another handle created with the same model, num_envs, seed and env_offset, resumes bit-identically.
 * `buf` may be device or (pinned or pageable) host memory of pbg_snapshot_bytes(h) bytes; the copy is ordered on `stream`
 * (pageable host memory makes it synchronous).  The blob is opaque and only valid for this library version. */
int64_t pbg_snapshot_bytes(const pbg_handle *h);
int pbg_snapshot(pbg_handle *h, void *buf, void *stream);
int pbg_restore(pbg_handle *h, const void *buf, void *stream);

/* Physics only (apply_action + stepSimulation), no task bookkeeping; for single-step parity tests. */
int pbg_physics_step(pbg_handle *h, const float *actions_dev, void *stream);
/* The task half of an env step on the current state, without physics: calc_state + reward + termination with the given
 * actions in the electricity terms; outputs as pbg_step.  pbg_physics_step + pbg_observe == pbg_step without the episode
 * counters / auto-reset.  NOT a pure read: like the reference's _step it latches this step's feet flags (quirk Q2), updates
 * the stored potential, and for the Flagrun kinds advances flag_timeout / frame / on_ground counters and may move the flag or
 * throw the cube.  For the 1e-5 observation/reward parity tier. */
int pbg_observe(pbg_handle *h, const float *actions_dev, float *obs_dev, float *reward_dev, uint8_t *done_dev,
                float *reward_terms_dev, void *stream);
/* feet_contact flags as the reference's robot.feet_contact, float[num_envs, nfeet] */
int pbg_get_feet_contact(pbg_handle *h, float *out_dev, void *stream);
/* pbg_physics_step that also reports the contact points active in the last substep, int32[num_envs]
 * (what getContactPoints would list after stepSimulation, rs/robot_bases.py:281) */
int pbg_physics_step_counts(pbg_handle *h, const float *actions_dev, int32_t *ncontact_dev, void *stream);
/* contact points the kernel of `kind` keeps per env (deepest first); the model passed to pbg_create
 * must carry the same max_contacts */
int pbg_max_contacts(int kind);
/* Constraint-row budget of the kernel for this env kind: with nl violated joint limits in a sub-step at most
 * min(pbg_max_contacts, (pbg_max_rows - nl) / 3) contacts (the deepest) get solver rows.  Equals nlim + 3 * max_contacts (never
 * binding) except for the humanoid kinds, where it is 36; an oracle compared with the library applies the
 * same rule. */
int pbg_max_rows(int kind);

int pbg_stats(pbg_handle *h, pbg_episode_stats *out_host, int32_t reset);

/* Re-keys the counter RNG that draws the reset noise (when not injected), the Flagrun flag positions and the cube attacks;
 * takes effect at the next reset.  The reference's env.seed(s) (rs/env_bases.py:41-44) re-seeds the np_random these come from. */
int pbg_set_seed(pbg_handle *h, uint64_t seed);

/* Task bookkeeping per env, double[num_envs, PBG_TASK_VIEW_DIM] on the device:
 *   [0] potential (env.potential, rs/gym_locomotion_envs.py:67-68)   [1] walk_target_x   [2] walk_target_y
 *   [3] flag_timeout (rs/robot_locomotors.py:204-218)   [4] frame   [5] on_ground_frame_counter (:250-273)
 *   [6] steps of the running episode   [7] return of the running episode   [8] initial_z   [9] episode index
 *   [10] cube attacks so far   [11] flag moves so far
 * Lets a host shell mirror robot.walk_target_x/y, flag_timeout, frame and env.potential after each step. */
#define PBG_TASK_VIEW_DIM 12
int pbg_get_task_view(pbg_handle *h, double *out_dev, void *stream);

/* Contact export (what getContactPoints(bodyA, -1, linkA, -1) lists, rs/robot_bases.py:280-281).  After
 * pbg_enable_contact_export(h, 1) every pbg_step / pbg_physics_step* records, per env and contact-candidate slot, the
 * distance of that candidate in the step's last collision pass, or +inf when it is not within its breaking threshold.
 * Slots, in order: for every geom with geom_ground != 0, in geom order, one slot per sphere / two per capsule (end spheres)
 * against the floor; 8 cube corners against the floor (worlds with the cube); the npair self-collision geom pairs; every
 * geom against the cube (worlds with the cube).  pbg_get_contact_candidates copies float[num_envs, pbg_num_contact_slots]. */
int pbg_num_contact_slots(const pbg_handle *h);
int pbg_enable_contact_export(pbg_handle *h, int32_t enabled);
int pbg_get_contact_candidates(pbg_handle *h, float *out_dev, void *stream);

/* Measures the FP32 CUDA-core peak of `device` with an FFMA microbenchmark (TFLOP/s); the roofline
 * denominator MEASURED_PEAKS.json does not carry. */
int pbg_measure_fp32_peak(int32_t device, double *tflops_out);

/* number of kernels this handle has launched so far (bench.py's gpu_launches) */
int64_t pbg_launch_count(const pbg_handle *h);

#ifdef __cplusplus
}
#endif
#endif
