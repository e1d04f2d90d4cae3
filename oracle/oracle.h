/*
 * oracle.h -- CPU restatement of the reference hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library; the product path (pybullet_gym_b200 + libpbg_b200.so) never does.
 *
 * PARITY STATUS.  The observation / reward / termination layer is restated from reference source
 * and pinned by golden vectors generated from the reference's own Python (tests/golden/,
 * tools/gen_golden_task.py).  The physics step (what `stepSimulation()` does inside the un-vendored
 * pybullet wheel, /root/reference/setup.py:34 `pybullet>=1.7.8`, no pinned version) is restated
 * from the published btMultiBody algorithm and is **parity unpinned**:
 * pybullet cannot be imported or installed in this image, and the reference ships no golden
 * vectors for it.
 *
 * The oracle simulates the *Bullet-shaped* link list (one massless link per <joint>, fixed
 * `jointfix` links kept as 0-dof links) with Featherstone's articulated-body algorithm, builds
 * constraint rows the way btMultiBodyConstraintSolver does (unit-impulse responses through the
 * articulated inertias, delta-velocity PGS, limit rows -> contact normals -> friction rows), and
 * integrates semi-implicitly.  It shares no code with the CUDA library.
 */
#ifndef PBG_ORACLE_H
#define PBG_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_JT_FIXED = 0, ORC_JT_REVOLUTE = 1, ORC_JT_PRISMATIC = 2, ORC_JT_FREE = 3 };
enum { ORC_G_SPHERE = 0, ORC_G_CAPSULE = 1, ORC_G_BOX = 2 };
enum { ORC_KIND_PENDULUM = 0, ORC_KIND_PENDULUM_SWINGUP = 1, ORC_KIND_HOPPER = 2, ORC_KIND_WALKER2D = 3,
       ORC_KIND_HALFCHEETAH = 4, ORC_KIND_ANT = 5, ORC_KIND_HUMANOID = 6, ORC_KIND_FLAGRUN = 7,
       ORC_KIND_FLAGRUN_HARDER = 8, ORC_KIND_DOUBLE_PENDULUM = 9, ORC_KIND_REACHER = 10, ORC_KIND_DOUBLE_PENDULUM_MJ = 11,
       ORC_KIND_HOPPER_MJ = 12, ORC_KIND_WALKER2D_MJ = 13, ORC_KIND_ANT_MJ = 14, ORC_KIND_HUMANOID_MJ = 15,
       ORC_KIND_HALFCHEETAH_MJ = 16 };

/* Bullet-shaped model + scene + task constants.  All arrays are owned by the caller. */
typedef struct {
    /* links (index 0 = base) */
    int32_t nl, floating;
    const int32_t *parent, *jtype;            /* [nl] */
    const double *axis, *pos, *quat, *com;    /* [nl*3], [nl*3], [nl*4 xyzw], [nl*3] */
    const double *mass, *inertia;             /* [nl], [nl*3] */
    const double *lower, *upper, *damping;    /* [nl] joint limits / damping of the link's joint */
    const int32_t *in_parts;                  /* [nl] link is an entry of robot.parts */
    /* geoms */
    int32_t ng;
    const int32_t *g_link, *g_type, *g_ground;
    const double *g_radius, *g_p0, *g_p1, *g_friction, *g_threshold;
    int32_t npair;
    const int32_t *pair_a, *pair_b;
    /* scene (rs/scene_bases.py:60-73, rs/scene_stadium.py:33) */
    double gravity, dt_sub;
    int32_t nsub, niter;
    double erp_contact, erp_limit, linear_slop, warmstart, link_damping, max_coord_vel;
    double ground_friction, limit_max_impulse, split_impulse_threshold;
    int32_t limit_split_impulse, max_contacts;
    /* task */
    int32_t kind, nact, nfeet, torso_link, obs_dim;
    const int32_t *act_link;                  /* [nact] link whose joint action n drives */
    const double *act_torque;                 /* [nact] power * power_coef */
    const int32_t *foot_link;                 /* [nfeet] */
    double initial_z;                         /* < 0: latch torso z at reset */
    double elec_cost, stall_cost, limit_cost, dt_scene;
    double walk_target_x, walk_target_y;
    int32_t max_episode_steps;
    double stadium_halflen, stadium_halfwidth;   /* rs/scene_stadium.py:13-14, Flagrun target range */
    /* torsional friction rows (spinning about the normal, rolling about the two tangents), SURVEY C1.11 / C6-9 */
    int32_t torsional;
    const double *g_spin, *g_roll;            /* [ng] 2nd / 3rd MJCF friction numbers */
    double ground_spin, ground_roll;          /* the floor's spinning / rolling friction (changeDynamics on the floor body) */
    /* HumanoidFlagrunHarder's aggressive cube: a second free body (rs/robot_locomotors.py:236-266,
     * gym_utils.py:9-16, assets/things/cube_small.urdf).  cube = 0: absent */
    int32_t cube;
    double cube_half, cube_mass, cube_inertia, cube_friction, cube_threshold;
    double cube_pos0[3];
    /* friction_cone = 1: the two friction rows of a contact are solved as a pair and projected onto the circle
     * mu * lambda_n, also when lambda_n = 0 (btMultiBodyConstraintSolver::resolveConeFrictionConstraintRows, Bullet >= 2.88);
     * 0: independent rows with box bounds, skipped while lambda_n = 0 (the older path; the CUDA kernel's).  Oracle-only probe. */
    int32_t friction_cone;
    /* links whose COM the task layer reads besides torso_link (Reacher: fingertip, target); -1: unused */
    int32_t aux_link[2];
    /* ground_manifold = 1: geom-vs-floor contacts come from a persistent manifold (one new point per collision pass, up to four
     * cached points per geom, 0.02-style breaking threshold) instead of the instantaneous end-sphere candidates
     * the CUDA kernel tests.  Oracle-only probe (tools/policy_probe.py). */
    int32_t ground_manifold;
} orc_model;

typedef struct orc_env orc_env;

/* canonical state: [base pos(3) quat(4 xyzw) omega(3) vel(3)] (floating only) + q[nd] + qd[nd]
 * + [cube pos(3) quat(4) omega(3) vel(3)] (models with a cube only) */
int orc_state_size(const orc_model *m);
int orc_num_dofs(const orc_model *m);

orc_env *orc_create(const orc_model *m, uint64_t seed, uint64_t env_index);
void orc_destroy(orc_env *e);
/* reset: MJCF pose, every actuated joint <- U(-0.1,0.1) from the counter RNG; floor_in_parts: quirk Q1 */
void orc_reset(orc_env *e, int floor_in_parts, double *obs);
/* reset with injected joint noise (ordered-joint order; double pendulum: hinge, hinge2); used to replay
 * reference-style resets */
void orc_reset_with(orc_env *e, const double *joint_noise, int floor_in_parts, double *obs);
/* one env step: returns done; terms = [alive, progress, electricity, joints_at_limit, feet_collision] */
int orc_step(orc_env *e, const double *action, double *obs, double *reward, double *terms);
/* physics only: nsub substeps with the given action, no task bookkeeping */
void orc_physics_step(orc_env *e, const double *action);
void orc_get_state(const orc_env *e, double *state);
void orc_set_state(orc_env *e, const double *state);
/* observation / task bookkeeping of the *current* state (no physics); same outputs as orc_step */
int orc_observe(orc_env *e, const double *action, double *obs, double *reward, double *terms);
/* hooks for the fake-pybullet backend (tools/fake_pybullet.py) */
void orc_physics_step_torque(orc_env *e, const double *tau /* [nd] */);
void orc_link_state(orc_env *e, double *out /* [nl*10]: com3 quat4 vel3 */);
int orc_get_contacts(const orc_env *e, int32_t *la, int32_t *lb, double *dist);
void orc_set_joint(orc_env *e, int dof, double q, double qd);
void orc_get_joint(const orc_env *e, int dof, double *q, double *qd);
/* cube pose / velocity access (resetBasePositionAndOrientation / resetBaseVelocity / getBasePositionAndOrientation
 * on the cube body); NULL arguments are left unchanged / not written */
void orc_set_cube(orc_env *e, const double *pos, const double *quat, const double *omega, const double *vel);
void orc_get_cube(const orc_env *e, double *pos, double *quat, double *omega, double *vel);
/* replay tape: when set, every random draw of the task layer (flag positions, cube attack) is read from it
 * in order instead of the counter RNG -- used to replay the reference's np_random draws */
void orc_set_tape(orc_env *e, const double *tape, int n);
/* diagnostics */
int orc_get_rows(const orc_env *e, double *out);
int orc_num_contacts(const orc_env *e);
void orc_feet_contact(const orc_env *e, double *out);
void orc_link_com(orc_env *e, double *xyz_out /* [nl*3] */);
double orc_energy(orc_env *e);              /* kinetic + potential, for invariants */
/* mass matrix M (nd_total x nd_total, row-major) via unit-impulse responses, for cross-checks */
void orc_mass_matrix_inv(orc_env *e, double *Minv);
/* throughput helper for the CPU baseline: run `steps` env steps with U(-1,1) actions, auto-reset */
long orc_rollout(orc_env *e, long steps, uint64_t action_seed, double *ret_sum, long *episodes);
/* `n` whole random-policy episodes (reset, U(-1,1) actions until done or `cap` steps): per-episode return / length */
long orc_episodes(orc_env *e, int n, int cap, uint64_t action_seed, double *returns, int32_t *lengths);
/* collide() passes in which max_contacts dropped candidates since creation */
int orc_cap_overflows(const orc_env *e);
/* constraint-row budget (0 = none): with nl violated joint limits in a sub-step, at most (max_rows - nl) / 3 contacts are kept
 * (the deepest), on top of max_contacts; mirrors pbg_max_rows() of the CUDA library */
void orc_set_max_rows(orc_env *e, int max_rows);
/* per foot: how close its nearest floor candidate was to flipping the feet flag in the last collision pass
 * (min |distance - breaking threshold|); lets a test excuse exactly the flag disagreements that are round-off */
void orc_feet_margin(const orc_env *e, double *out);
/* how far the last orc_observe's termination test was from flipping: min |value - threshold| over the comparisons it made
 * (z, pitch, pole angle ...; not the integer / flag inputs) */
double orc_done_margin(const orc_env *e);

#ifdef __cplusplus
}
#endif
#endif
