/*
 * oracle.c -- CPU restatement of the reference hot path (see oracle.h).  TEST INFRASTRUCTURE ONLY.
 *
 * Reference lines each block follows:
 *   apply_action            rs/robot_locomotors.py:26-29,185-189 ; rs/robot_pendula.py:20-25
 *   stepSimulation          rs/scene_bases.py:60-76  ([EXT] SURVEY.md Appendix C2-C5, parity unpinned)
 *   calc_state              rs/robot_locomotors.py:31-64 ; rs/robot_bases.py:306-321
 *   reward / termination    rs/gym_locomotion_envs.py:54-114 ; rs/gym_pendulum_envs.py:26-39
 *   reset                   rs/robot_locomotors.py:16-24 ; rs/gym_locomotion_envs.py:22-39
 * (rs/ = /root/reference/pybulletgym/envs/roboschool/)
 */
#include "oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define MAXL 48
#define MAXD 32
#define MAXU (MAXD + 12)        /* base 6 + joints + cube 6 */
#define MAXG 32
#define MAXCAND 320
#define MAXROWS 256
#define MAXFEET 8

typedef double v3[3];
typedef double m3[3][3];
typedef double sv[6];          /* spatial vector: angular(3), linear(3), world coordinates at the world origin */
typedef double sm[6][6];

/* ------------------------------------------------------------------ small math */
static void v3set(v3 a, double x, double y, double z) { a[0] = x; a[1] = y; a[2] = z; }
static void v3cpy(v3 a, const v3 b) { a[0] = b[0]; a[1] = b[1]; a[2] = b[2]; }
static void v3add(v3 o, const v3 a, const v3 b) { o[0] = a[0] + b[0]; o[1] = a[1] + b[1]; o[2] = a[2] + b[2]; }
static void v3sub(v3 o, const v3 a, const v3 b) { o[0] = a[0] - b[0]; o[1] = a[1] - b[1]; o[2] = a[2] - b[2]; }
static void v3axpy(v3 o, double s, const v3 a) { o[0] += s * a[0]; o[1] += s * a[1]; o[2] += s * a[2]; }
static double v3dot(const v3 a, const v3 b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static void v3cross(v3 o, const v3 a, const v3 b) {
    double x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
    o[0] = x; o[1] = y; o[2] = z;
}
static double v3norm(const v3 a) { return sqrt(v3dot(a, a)); }
static void m3mulv(v3 o, m3 m, const v3 a) {
    double x = m[0][0] * a[0] + m[0][1] * a[1] + m[0][2] * a[2];
    double y = m[1][0] * a[0] + m[1][1] * a[1] + m[1][2] * a[2];
    double z = m[2][0] * a[0] + m[2][1] * a[1] + m[2][2] * a[2];
    o[0] = x; o[1] = y; o[2] = z;
}
static void m3mul(m3 o, m3 a, m3 b) {
    m3 t;
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) t[i][j] = a[i][0] * b[0][j] + a[i][1] * b[1][j] + a[i][2] * b[2][j];
    memcpy(o, t, sizeof(m3));
}
static void q2m(m3 m, const double q[4]) {   /* (x,y,z,w) */
    double x = q[0], y = q[1], z = q[2], w = q[3];
    m[0][0] = 1 - 2 * (y * y + z * z); m[0][1] = 2 * (x * y - w * z); m[0][2] = 2 * (x * z + w * y);
    m[1][0] = 2 * (x * y + w * z); m[1][1] = 1 - 2 * (x * x + z * z); m[1][2] = 2 * (y * z - w * x);
    m[2][0] = 2 * (x * z - w * y); m[2][1] = 2 * (y * z + w * x); m[2][2] = 1 - 2 * (x * x + y * y);
}
static void qmul(double o[4], const double a[4], const double b[4]) {
    double x = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
    double y = a[3] * b[1] - a[0] * b[2] + a[1] * b[3] + a[2] * b[0];
    double z = a[3] * b[2] + a[0] * b[1] - a[1] * b[0] + a[2] * b[3];
    double w = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
    o[0] = x; o[1] = y; o[2] = z; o[3] = w;
}
static void axis_angle_m(m3 m, const v3 ax, double ang) {
    /* btQuaternion(axis, angle) divides by the axis length: a non-unit joint axis rotates by `ang` all the
     * same, while the motion subspace S keeps the raw axis */
    double q[4], s = sin(0.5 * ang) / sqrt(ax[0] * ax[0] + ax[1] * ax[1] + ax[2] * ax[2]);
    q[0] = ax[0] * s; q[1] = ax[1] * s; q[2] = ax[2] * s; q[3] = cos(0.5 * ang);
    q2m(m, q);
}
static void sv_zero(sv a) { for (int i = 0; i < 6; i++) a[i] = 0; }
static double sv_dot(const sv a, const sv b) { double s = 0; for (int i = 0; i < 6; i++) s += a[i] * b[i]; return s; }
static void sm_mulv(sv o, sm m, const sv a) {
    sv t;
    for (int i = 0; i < 6; i++) { double s = 0; for (int j = 0; j < 6; j++) s += m[i][j] * a[j]; t[i] = s; }
    memcpy(o, t, sizeof(sv));
}
/* motion x motion */
static void sv_crm(sv o, const sv v, const sv s) {
    v3 a, b, c;
    v3cross(a, v, s); v3cross(b, v, s + 3); v3cross(c, v + 3, s);
    o[0] = a[0]; o[1] = a[1]; o[2] = a[2]; o[3] = b[0] + c[0]; o[4] = b[1] + c[1]; o[5] = b[2] + c[2];
}
/* motion x* force */
static void sv_crf(sv o, const sv v, const sv f) {
    v3 a, b, c;
    v3cross(a, v, f); v3cross(b, v + 3, f + 3); v3cross(c, v, f + 3);
    o[0] = a[0] + b[0]; o[1] = a[1] + b[1]; o[2] = a[2] + b[2]; o[3] = c[0]; o[4] = c[1]; o[5] = c[2];
}
/* solve 6x6 SPD-ish system by Gaussian elimination with partial pivoting */
static void solve6(sm A, const sv b, sv x) {
    double a[6][7];
    for (int i = 0; i < 6; i++) { for (int j = 0; j < 6; j++) a[i][j] = A[i][j]; a[i][6] = b[i]; }
    for (int c = 0; c < 6; c++) {
        int p = c;
        for (int r = c + 1; r < 6; r++) if (fabs(a[r][c]) > fabs(a[p][c])) p = r;
        if (p != c) for (int j = 0; j < 7; j++) { double t = a[c][j]; a[c][j] = a[p][j]; a[p][j] = t; }
        double d = a[c][c];
        for (int r = c + 1; r < 6; r++) { double f = a[r][c] / d; for (int j = c; j < 7; j++) a[r][j] -= f * a[c][j]; }
    }
    for (int i = 5; i >= 0; i--) { double s = a[i][6]; for (int j = i + 1; j < 6; j++) s -= a[i][j] * x[j]; x[i] = s / a[i][i]; }
}

/* ------------------------------------------------------------------ counter RNG (Philox4x32-10) */
static void philox(uint32_t c[4], const uint32_t k0[2]) {
    uint32_t k[2] = {k0[0], k0[1]};
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k[0], n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k[1], n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k[0] += 0x9E3779B9u; k[1] += 0xBB67AE85u;
    }
}
/* draw #n of stream `stream` in episode `ep` of env `env`: uniform in [lo,hi), float32 arithmetic
 * (bit-identical to the device RNG in pybullet_gym_b200/csrc) */
static double rng_uniform_s(uint64_t seed, uint64_t env, uint32_t ep, uint32_t stream, uint32_t n, float lo, float hi) {
    uint32_t c[4] = {(uint32_t)env, (uint32_t)(env >> 32), ep, (stream << 24) | (n >> 2)};
    uint32_t k[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    philox(c, k);
    float u = (float)(c[n & 3] >> 8) * (1.0f / 16777216.0f);
    return (double)fmaf(hi - lo, u, lo);
}
static double rng_uniform(uint64_t seed, uint64_t env, uint32_t ep, uint32_t n, float lo, float hi) {
    return rng_uniform_s(seed, env, ep, 0, n, lo, hi);
}

/* ------------------------------------------------------------------ env */
struct orc_env;
static double draw(struct orc_env *e, uint32_t stream, uint32_t n, float lo, float hi);

#define LINK_FLOOR (-1)
#define LINK_CUBE (-2)
typedef struct {
    int ga, gb;         /* geom (gb = -1: floor or cube) */
    int la, lb;         /* links; LINK_FLOOR, LINK_CUBE for the other bodies */
    int slot;           /* warm-start slot */
    v3 pa, pb, n;       /* witness points (world), normal from B to A */
    double dist, mu, mu_spin, mu_roll;
} contact;

typedef struct {
    double J[MAXU], U[MAXU];   /* jacobian, unit-impulse response */
    double rhs, lo, hi, dinv, lambda;
    int fric_of;               /* index of the normal row for friction rows, else -1 */
    double mu;
} row;

struct orc_env {
    orc_model m;
    uint64_t seed, env_index;
    uint32_t episode;
    int nd, nu;                 /* joint dofs, generalized velocities of the robot */
    int nut;                    /* nu + 6 when the world holds the cube */
    double xp[3], xq[4], xw[3], xv[3];   /* cube: COM position, orientation (xyzw), angular / linear velocity */
    int dof_of_link[MAXL], link_of_dof[MAXD];
    /* state */
    double bpos[3], bquat[4], bomega[3], bvel[3];
    double q[MAXD], qd[MAXD];
    double tau[MAXD];
    double warm[MAXCAND];
    /* kinematics cache */
    m3 R[MAXL]; v3 p[MAXL], c[MAXL];   /* link frame rotation, origin, COM (world) */
    sv S[MAXL], V[MAXL];
    sm IA[MAXL]; sv UU[MAXL]; double DD[MAXL];
    /* contacts */
    contact ct[MAXCAND]; int nct;
    int cand_active[MAXCAND];
    double feet_touch[MAXFEET];  /* foot-vs-floor candidates under the breaking threshold in the last collide(), BEFORE the solver cap */
    int max_rows;                /* > 0: constraint-row budget of the solver (limit rows first, contacts get the rest) */
    int cap_overflow;            /* collide() calls in which max_contacts dropped candidates */
    double done_margin;          /* distance of the last orc_observe's termination test from flipping (min over its comparisons) */
    double feet_margin[MAXFEET]; /* min over the foot's floor candidates of |distance - breaking threshold| in the last collide() */
    /* task state */
    double feet_contact[MAXFEET];
    double initial_z, potential;
    int have_initial_z, floor_in_parts, steps;
    double body_xyz[3], body_rpy[3], joint_speeds[MAXD];
    int joints_at_limit;
    double walk_target_x, walk_target_y, walk_target_dist;
    double torso_speed[3];
    row rows[MAXROWS];
    int last_nlim, last_nr;
    /* Flagrun / FlagrunHarder (rs/robot_locomotors.py:195-302) */
    double flag_timeout; int flag_count, frame, on_ground, crawl_has, attacks; double crawl_start, crawl_ign;
    double tape[4096]; int tape_n, tape_pos;
    /* ground_manifold = 1 (oracle-only probe): btPersistentManifold of each geom-vs-floor pair -- up to four cached
     * points, each the geom's point in its link frame and the floor point in world coordinates */
    struct { int on; v3 la, pb; } mf[MAXG][4];
    double last_dv[MAXU], last_ufree[MAXU], last_u0[MAXU];
};

static double draw(struct orc_env *e, uint32_t stream, uint32_t n, float lo, float hi) {
    if (e->tape_pos < e->tape_n) return e->tape[e->tape_pos++];
    return rng_uniform_s(e->seed, e->env_index, e->episode, stream, n, lo, hi);
}
void orc_set_tape(orc_env *e, const double *tape, int n) {
    if (n > 4096) n = 4096;
    for (int i = 0; i < n; i++) e->tape[i] = tape[i];
    e->tape_n = n; e->tape_pos = 0;
}

int orc_num_dofs(const orc_model *m) {
    int n = 0;
    for (int i = 0; i < m->nl; i++) if (m->jtype[i] == ORC_JT_REVOLUTE || m->jtype[i] == ORC_JT_PRISMATIC) n++;
    return n;
}
int orc_state_size(const orc_model *m) { return (m->floating ? 13 : 0) + 2 * orc_num_dofs(m) + (m->cube ? 13 : 0); }

orc_env *orc_create(const orc_model *m, uint64_t seed, uint64_t env_index) {
    orc_env *e = (orc_env *)calloc(1, sizeof(orc_env));
    e->m = *m;
    e->seed = seed; e->env_index = env_index;
    int n = 0;
    for (int i = 0; i < m->nl; i++) {
        e->dof_of_link[i] = -1;
        if (m->jtype[i] == ORC_JT_REVOLUTE || m->jtype[i] == ORC_JT_PRISMATIC) { e->dof_of_link[i] = n; e->link_of_dof[n] = i; n++; }
    }
    e->nd = n; e->nu = n + (m->floating ? 6 : 0);
    e->nut = e->nu + (m->cube ? 6 : 0);
    e->bquat[3] = 1.0; e->xq[3] = 1.0;
    for (int k = 0; k < 3; k++) e->xp[k] = m->cube_pos0[k];
    e->walk_target_x = m->walk_target_x; e->walk_target_y = m->walk_target_y;
    return e;
}
void orc_destroy(orc_env *e) { free(e); }

void orc_get_state(const orc_env *e, double *s) {
    int o = 0;
    if (e->m.floating) {
        for (int i = 0; i < 3; i++) s[o++] = e->bpos[i];
        for (int i = 0; i < 4; i++) s[o++] = e->bquat[i];
        for (int i = 0; i < 3; i++) s[o++] = e->bomega[i];
        for (int i = 0; i < 3; i++) s[o++] = e->bvel[i];
    }
    for (int i = 0; i < e->nd; i++) s[o++] = e->q[i];
    for (int i = 0; i < e->nd; i++) s[o++] = e->qd[i];
    if (e->m.cube) {
        for (int i = 0; i < 3; i++) s[o++] = e->xp[i];
        for (int i = 0; i < 4; i++) s[o++] = e->xq[i];
        for (int i = 0; i < 3; i++) s[o++] = e->xw[i];
        for (int i = 0; i < 3; i++) s[o++] = e->xv[i];
    }
}
void orc_set_state(orc_env *e, const double *s) {
    int o = 0;
    if (e->m.floating) {
        for (int i = 0; i < 3; i++) e->bpos[i] = s[o++];
        for (int i = 0; i < 4; i++) e->bquat[i] = s[o++];
        for (int i = 0; i < 3; i++) e->bomega[i] = s[o++];
        for (int i = 0; i < 3; i++) e->bvel[i] = s[o++];
    }
    for (int i = 0; i < e->nd; i++) e->q[i] = s[o++];
    for (int i = 0; i < e->nd; i++) e->qd[i] = s[o++];
    if (e->m.cube) {
        for (int i = 0; i < 3; i++) e->xp[i] = s[o++];
        for (int i = 0; i < 4; i++) e->xq[i] = s[o++];
        for (int i = 0; i < 3; i++) e->xw[i] = s[o++];
        for (int i = 0; i < 3; i++) e->xv[i] = s[o++];
    }
    memset(e->warm, 0, sizeof(e->warm));
}

/* ------------------------------------------------------------------ kinematics */
static void fk(orc_env *e) {
    const orc_model *m = &e->m;
    for (int i = 0; i < m->nl; i++) {
        m3 Rl; v3 pl;
        if (m->parent[i] < 0) {
            if (m->jtype[i] == ORC_JT_FREE) {
                /* state holds the pose of the base inertial frame (pybullet getBasePositionAndOrientation) */
                q2m(Rl, e->bquat);
                v3 t; m3mulv(t, Rl, &m->com[3 * i]);
                v3sub(pl, e->bpos, t);
            } else {
                q2m(Rl, &m->quat[4 * i]);
                v3cpy(pl, &m->pos[3 * i]);
            }
        } else {
            int pa = m->parent[i];
            m3 Rq; q2m(Rq, &m->quat[4 * i]);
            m3mul(Rl, e->R[pa], Rq);
            v3 t; m3mulv(t, e->R[pa], &m->pos[3 * i]);
            v3add(pl, e->p[pa], t);
            int d = e->dof_of_link[i];
            if (m->jtype[i] == ORC_JT_REVOLUTE) {
                m3 Rj; axis_angle_m(Rj, &m->axis[3 * i], e->q[d]);
                m3mul(Rl, Rl, Rj);
            } else if (m->jtype[i] == ORC_JT_PRISMATIC) {
                v3 a; m3mulv(a, Rl, &m->axis[3 * i]);
                v3axpy(pl, e->q[d], a);
            }
        }
        memcpy(e->R[i], Rl, sizeof(m3));
        v3cpy(e->p[i], pl);
        v3 t; m3mulv(t, Rl, &m->com[3 * i]);
        v3add(e->c[i], pl, t);
        sv_zero(e->S[i]);
        if (m->jtype[i] == ORC_JT_REVOLUTE) {
            v3 z; m3mulv(z, Rl, &m->axis[3 * i]);
            v3 az; v3cross(az, pl, z);          /* anchor = link-frame origin */
            e->S[i][0] = z[0]; e->S[i][1] = z[1]; e->S[i][2] = z[2];
            e->S[i][3] = az[0]; e->S[i][4] = az[1]; e->S[i][5] = az[2];
        } else if (m->jtype[i] == ORC_JT_PRISMATIC) {
            v3 z; m3mulv(z, Rl, &m->axis[3 * i]);
            e->S[i][3] = z[0]; e->S[i][4] = z[1]; e->S[i][5] = z[2];
        }
    }
}

static void velocities(orc_env *e) {
    const orc_model *m = &e->m;
    for (int i = 0; i < m->nl; i++) {
        if (m->parent[i] < 0) {
            sv_zero(e->V[i]);
            if (m->jtype[i] == ORC_JT_FREE) {
                v3 t; v3cross(t, e->bomega, e->c[i]);
                for (int k = 0; k < 3; k++) { e->V[i][k] = e->bomega[k]; e->V[i][3 + k] = e->bvel[k] - t[k]; }
            }
        } else {
            int d = e->dof_of_link[i];
            double qd = d >= 0 ? e->qd[d] : 0.0;
            for (int k = 0; k < 6; k++) e->V[i][k] = e->V[m->parent[i]][k] + e->S[i][k] * qd;
        }
    }
}

static void link_com_vel(const orc_env *e, int i, v3 out) {   /* classical COM velocity */
    v3 t; v3cross(t, e->V[i], e->c[i]);
    for (int k = 0; k < 3; k++) out[k] = e->V[i][3 + k] + t[k];
}
static void point_vel(const orc_env *e, int i, const v3 pt, v3 out) {
    v3 t; v3cross(t, e->V[i], pt);
    for (int k = 0; k < 3; k++) out[k] = e->V[i][3 + k] + t[k];
}

static void spatial_inertia(const orc_env *e, int i, sm I) {
    const orc_model *m = &e->m;
    m3 Ic, D = {{m->inertia[3 * i], 0, 0}, {0, m->inertia[3 * i + 1], 0}, {0, 0, m->inertia[3 * i + 2]}};
    m3 Rt;
    for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) Rt[a][b] = e->R[i][b][a];
    m3mul(Ic, (double(*)[3])e->R[i], D); m3mul(Ic, Ic, Rt);
    double ms = m->mass[i];
    const double *c = e->c[i];
    m3 cx = {{0, -c[2], c[1]}, {c[2], 0, -c[0]}, {-c[1], c[0], 0}};
    m3 cxcx; m3mul(cxcx, cx, cx);
    for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) {
        I[a][b] = Ic[a][b] - ms * cxcx[a][b];
        I[a][3 + b] = ms * cx[a][b];
        I[3 + a][b] = -ms * cx[a][b];
        I[3 + a][3 + b] = (a == b) ? ms : 0.0;
    }
}

/* articulated inertias (configuration only): IA, U = IA S, D = S^T U */
static void articulated_inertias(orc_env *e) {
    const orc_model *m = &e->m;
    for (int i = 0; i < m->nl; i++) spatial_inertia(e, i, e->IA[i]);
    for (int i = m->nl - 1; i >= 1; i--) {
        int pa = m->parent[i];
        if (pa < 0) continue;
        if (e->dof_of_link[i] >= 0) {
            sm_mulv(e->UU[i], e->IA[i], e->S[i]);
            e->DD[i] = sv_dot(e->S[i], e->UU[i]);
            for (int a = 0; a < 6; a++) for (int b = 0; b < 6; b++)
                e->IA[pa][a][b] += e->IA[i][a][b] - e->UU[i][a] * e->UU[i][b] / e->DD[i];
        } else {
            for (int a = 0; a < 6; a++) for (int b = 0; b < 6; b++) e->IA[pa][a][b] += e->IA[i][a][b];
        }
    }
}

/* Generalized acceleration response of the tree.
 *   pl[i]   : bias force of link i (spatial, world origin); may be NULL (=0)
 *   cl[i]   : velocity-product acceleration of link i; may be NULL
 *   tauj[d] : joint forces
 * out: [base omega_dot(3), base spatial linear acc(3)] (floating) + qdd.  Base part is *spatial*;
 * the caller converts to the COM-referenced generalized coordinates. */
static void aba_solve(orc_env *e, sv *pl, sv *cl, const double *tauj, sv a0_out, double *qdd) {
    const orc_model *m = &e->m;
    static __thread sv pA[MAXL]; static __thread double uu[MAXL]; static __thread sv acc[MAXL];
    for (int i = 0; i < m->nl; i++) { if (pl) memcpy(pA[i], pl[i], sizeof(sv)); else sv_zero(pA[i]); }
    for (int i = m->nl - 1; i >= 1; i--) {
        int pa = m->parent[i];
        if (pa < 0) continue;
        int d = e->dof_of_link[i];
        sv pa_;
        if (d >= 0) {
            uu[i] = tauj[d] - sv_dot(e->S[i], pA[i]);
            /* Ia c = IA c - U (U^T c)/D */
            sv Iac; sv_zero(Iac);
            if (cl) {
                sm_mulv(Iac, e->IA[i], cl[i]);
                /* IA here already is the articulated inertia of link i itself (children folded in) */
                double uc = sv_dot(e->UU[i], cl[i]) / e->DD[i];
                for (int k = 0; k < 6; k++) Iac[k] -= e->UU[i][k] * uc;
            }
            for (int k = 0; k < 6; k++) pa_[k] = pA[i][k] + Iac[k] + e->UU[i][k] * uu[i] / e->DD[i];
        } else {
            sv Iac; sv_zero(Iac);
            if (cl) sm_mulv(Iac, e->IA[i], cl[i]);
            for (int k = 0; k < 6; k++) pa_[k] = pA[i][k] + Iac[k];
        }
        for (int k = 0; k < 6; k++) pA[pa][k] += pa_[k];
    }
    for (int i = 0; i < m->nl; i++) {
        int pa = m->parent[i];
        if (pa < 0) {
            if (m->jtype[i] == ORC_JT_FREE) {
                sv neg; for (int k = 0; k < 6; k++) neg[k] = -pA[i][k];
                solve6(e->IA[i], neg, acc[i]);
                memcpy(a0_out, acc[i], sizeof(sv));
            } else { sv_zero(acc[i]); if (a0_out) sv_zero(a0_out); }
            continue;
        }
        sv ap; memcpy(ap, acc[pa], sizeof(sv));
        if (cl) for (int k = 0; k < 6; k++) ap[k] += cl[i][k];
        int d = e->dof_of_link[i];
        if (d >= 0) {
            qdd[d] = (uu[i] - sv_dot(e->UU[i], ap)) / e->DD[i];
            for (int k = 0; k < 6; k++) acc[i][k] = ap[k] + e->S[i][k] * qdd[d];
        } else memcpy(acc[i], ap, sizeof(sv));
    }
}

/* delta generalized velocity for a generalized impulse f (layout of u): M^-1 f
 * (btMultiBody::calculateAccelerationDeltasMultiDof) */
static void impulse_response(orc_env *e, const double *f, double *du) {
    const orc_model *m = &e->m;
    static __thread sv pl[MAXL];
    for (int i = 0; i < m->nl; i++) sv_zero(pl[i]);
    const double *fj = f;
    if (m->floating) {
        /* base wrench given at the base COM: torque f[0:3], force f[3:6] */
        v3 t; v3cross(t, e->c[0], f + 3);
        for (int k = 0; k < 3; k++) { pl[0][k] = -(f[k] + t[k]); pl[0][3 + k] = -f[3 + k]; }
        fj = f + 6;
    }
    sv a0; double qdd[MAXD];
    aba_solve(e, pl, NULL, fj, a0, qdd);
    int o = 0;
    if (m->floating) {
        v3 t; v3cross(t, a0, e->c[0]);
        for (int k = 0; k < 3; k++) { du[k] = a0[k]; du[3 + k] = a0[3 + k] + t[k]; }
        o = 6;
    }
    for (int d = 0; d < e->nd; d++) du[o + d] = qdd[d];
    if (m->cube) {
        /* the cube is a btMultiBody without links: M^-1 = diag(1/I, 1/m) (isotropic box inertia) */
        for (int k = 0; k < 3; k++) { du[e->nu + k] = f[e->nu + k] / m->cube_inertia; du[e->nu + 3 + k] = f[e->nu + 3 + k] / m->cube_mass; }
    }
}

/* jacobian row: d . (velocity of world point pt rigidly attached to link i) w.r.t. u */
static void point_jacobian(const orc_env *e, int i, const v3 pt, const v3 d, double sign, double *J) {
    const orc_model *m = &e->m;
    int o = m->floating ? 6 : 0;
    int l = i;
    if (i == LINK_CUBE) {
        v3 r, rd; v3sub(r, pt, e->xp); v3cross(rd, r, d);
        for (int k = 0; k < 3; k++) { J[e->nu + k] += sign * rd[k]; J[e->nu + 3 + k] += sign * d[k]; }
        return;
    }
    while (l >= 0) {
        int k = e->dof_of_link[l];
        if (k >= 0) {
            /* S = (z, A x z): point velocity = z x pt + A x z  (revolute) or z (prismatic) */
            v3 t; v3cross(t, e->S[l], pt);
            v3 v = {t[0] + e->S[l][3], t[1] + e->S[l][4], t[2] + e->S[l][5]};
            J[o + k] += sign * v3dot(d, v);
        }
        if (m->parent[l] < 0 && m->jtype[l] == ORC_JT_FREE) {
            v3 r; v3sub(r, pt, e->c[l]);
            v3 rd; v3cross(rd, r, d);
            for (int k2 = 0; k2 < 3; k2++) { J[k2] += sign * rd[k2]; J[3 + k2] += sign * d[k2]; }
        }
        l = m->parent[l];
    }
}

/* jacobian row: axis . (angular velocity of link i) w.r.t. u (torsional friction rows) */
static void angular_jacobian(const orc_env *e, int i, const v3 axis, double sign, double *J) {
    const orc_model *m = &e->m;
    int o = m->floating ? 6 : 0, l = i;
    if (i == LINK_CUBE) { for (int k = 0; k < 3; k++) J[e->nu + k] += sign * axis[k]; return; }
    while (l >= 0) {
        int k = e->dof_of_link[l];
        if (k >= 0) J[o + k] += sign * v3dot(axis, e->S[l]);      /* S angular part: z (revolute) / 0 (prismatic) */
        if (m->parent[l] < 0 && m->jtype[l] == ORC_JT_FREE) for (int k2 = 0; k2 < 3; k2++) J[k2] += sign * axis[k2];
        l = m->parent[l];
    }
}

/* ------------------------------------------------------------------ collision */
static void geom_world(const orc_env *e, int g, v3 a, v3 b) {
    const orc_model *m = &e->m;
    int l = m->g_link[g];
    v3 t;
    m3mulv(t, (double(*)[3])e->R[l], &m->g_p0[3 * g]); v3add(a, e->p[l], t);
    m3mulv(t, (double(*)[3])e->R[l], &m->g_p1[3 * g]); v3add(b, e->p[l], t);
}

static void closest_seg_seg(const v3 p1, const v3 q1, const v3 p2, const v3 q2, v3 c1, v3 c2) {
    v3 d1, d2, r;
    v3sub(d1, q1, p1); v3sub(d2, q2, p2); v3sub(r, p1, p2);
    double a = v3dot(d1, d1), ee = v3dot(d2, d2), f = v3dot(d2, r), s, t;
    const double EPS = 1e-12;
    if (a <= EPS && ee <= EPS) { s = t = 0; }
    else if (a <= EPS) { s = 0; t = f / ee; t = t < 0 ? 0 : (t > 1 ? 1 : t); }
    else {
        double c = v3dot(d1, r);
        if (ee <= EPS) { t = 0; s = -c / a; s = s < 0 ? 0 : (s > 1 ? 1 : s); }
        else {
            double b = v3dot(d1, d2), den = a * ee - b * b;
            if (den > EPS) { s = (b * f - c * ee) / den; s = s < 0 ? 0 : (s > 1 ? 1 : s); } else s = 0;
            t = (b * s + f) / ee;
            if (t < 0) { t = 0; s = -c / a; s = s < 0 ? 0 : (s > 1 ? 1 : s); }
            else if (t > 1) { t = 1; s = (b - c) / a; s = s < 0 ? 0 : (s > 1 ? 1 : s); }
        }
    }
    for (int k = 0; k < 3; k++) { c1[k] = p1[k] + d1[k] * s; c2[k] = p2[k] + d2[k] * t; }
}

/* Closest points of the segment p0-p1 and an origin-centred box with half extents h, everything in
 * the box frame.  The squared distance of the segment point P(t) to the box is convex and piecewise
 * quadratic in t; its derivative is monotone, so 32 bisection steps on t in [0,1] pin the minimiser
 * (the CUDA kernel runs the same iteration in fp32).  Bullet uses GJK/EPA for box-vs-multisphere; this
 * is the exact closest-point pair GJK converges to for separated shapes. */
static void closest_seg_box(const v3 p0, const v3 p1, double h, v3 cs, v3 cb) {
    v3 d; v3sub(d, p1, p0);
    double lo = 0.0, hi = 1.0;
    for (int it = 0; it < 32; it++) {
        double t = 0.5 * (lo + hi), g = 0.0;
        for (int k = 0; k < 3; k++) {
            double x = p0[k] + t * d[k], c = x < -h ? -h : (x > h ? h : x);
            g += (x - c) * d[k];
        }
        if (g < 0.0) lo = t; else hi = t;
    }
    double t = 0.5 * (lo + hi);
    for (int k = 0; k < 3; k++) {
        double x = p0[k] + t * d[k];
        cs[k] = x; cb[k] = x < -h ? -h : (x > h ? h : x);
    }
}

static void collide(orc_env *e) {
    const orc_model *m = &e->m;
    static __thread contact all[MAXCAND];
    int n = 0, slot = 0;
    for (int f = 0; f < m->nfeet; f++) e->feet_margin[f] = 1e30;
    for (int g = 0; g < m->ng; g++) {
        int npt = m->g_type[g] == ORC_G_CAPSULE ? 2 : 1;
        v3 a, b; geom_world(e, g, a, b);
        if (m->ground_manifold) {
            /* btConvexPlaneCollisionAlgorithm::processCollision for a non-polyhedral convex: ONE new point per pass (the support
             * vertex towards the plane, kept if closer than the breaking threshold; btPersistentManifold::getCacheEntry replaces
             * the nearest cached point within that threshold, else the point is added), then refreshContactPoints: cached
             * points move with their bodies and are dropped once they separate or drift sideways by more than the threshold */
            const int l = m->g_link[g];
            const double thr = m->g_threshold[g], rad = m->g_radius[g];
            if (m->g_ground[g]) {
                const double *c = (npt == 2 && b[2] < a[2]) ? b : a;
                if (c[2] - rad < thr) {
                    v3 pw = {c[0], c[1], c[2] - rad}, t, la;
                    v3sub(t, pw, e->p[l]);
                    for (int i = 0; i < 3; i++) la[i] = e->R[l][0][i] * t[0] + e->R[l][1][i] * t[1] + e->R[l][2][i] * t[2];
                    int best = -1; double bd = thr * thr;
                    for (int k = 0; k < 4; k++) if (e->mf[g][k].on) {
                        v3 d; v3sub(d, e->mf[g][k].la, la);
                        if (v3dot(d, d) < bd) { bd = v3dot(d, d); best = k; }
                    }
                    if (best < 0) {
                        for (int k = 0; k < 4 && best < 0; k++) if (!e->mf[g][k].on) best = k;
                        if (best < 0) {   /* full: the shallowest cached point makes room (sortCachedPoints keeps the deepest) */
                            double wz = -1e30;
                            for (int k = 0; k < 4; k++) {
                                v3 pa; m3mulv(pa, (double(*)[3])e->R[l], e->mf[g][k].la); v3add(pa, pa, e->p[l]);
                                if (pa[2] > wz) { wz = pa[2]; best = k; }
                            }
                        }
                        e->warm[slot + best] = 0.0;
                    }
                    e->mf[g][best].on = 1; v3cpy(e->mf[g][best].la, la); v3set(e->mf[g][best].pb, c[0], c[1], 0.0);
                }
            }
            for (int k = 0; k < 4; k++, slot++) {
                e->cand_active[slot] = 0;
                if (!e->mf[g][k].on) continue;
                v3 pa; m3mulv(pa, (double(*)[3])e->R[l], e->mf[g][k].la); v3add(pa, pa, e->p[l]);
                const double dx = e->mf[g][k].pb[0] - pa[0], dy = e->mf[g][k].pb[1] - pa[1];
                if (!m->g_ground[g] || pa[2] > thr || dx * dx + dy * dy > thr * thr) { e->mf[g][k].on = 0; continue; }
                contact *ct = &all[n++];
                ct->ga = g; ct->gb = -1; ct->la = l; ct->lb = -1; ct->slot = slot;
                v3set(ct->n, 0, 0, 1); v3cpy(ct->pa, pa); v3cpy(ct->pb, e->mf[g][k].pb);
                ct->dist = pa[2]; ct->mu = m->g_friction[g] * m->ground_friction;
                ct->mu_spin = m->torsional ? m->g_spin[g] * m->ground_friction + m->ground_spin * m->g_friction[g] : 0;
                ct->mu_roll = m->torsional ? m->g_roll[g] * m->ground_friction + m->ground_roll * m->g_friction[g] : 0;
            }
            continue;
        }
        for (int k = 0; k < npt; k++, slot++) {
            e->cand_active[slot] = 0;
            if (!m->g_ground[g]) continue;
            const double *c = k ? b : a;
            double dist = c[2] - m->g_radius[g];
            for (int f = 0; f < m->nfeet; f++) if (m->foot_link[f] == m->g_link[g]) {
                double mg = fabs(dist - m->g_threshold[g]);
                if (mg < e->feet_margin[f]) e->feet_margin[f] = mg;
            }
            if (dist < m->g_threshold[g]) {
                contact *ct = &all[n++];
                ct->ga = g; ct->gb = -1; ct->la = m->g_link[g]; ct->lb = -1; ct->slot = slot;
                v3set(ct->n, 0, 0, 1);
                v3set(ct->pa, c[0], c[1], c[2] - m->g_radius[g]);
                v3set(ct->pb, c[0], c[1], 0.0);
                ct->dist = dist; ct->mu = m->g_friction[g] * m->ground_friction;
                /* btManifoldResult::calculateCombinedRolling/SpinningFriction: rollA*fricB + rollB*fricA, floor has 0 */
                ct->mu_spin = m->torsional ? m->g_spin[g] * m->ground_friction + m->ground_spin * m->g_friction[g] : 0;
                ct->mu_roll = m->torsional ? m->g_roll[g] * m->ground_friction + m->ground_roll * m->g_friction[g] : 0;
            }
        }
    }
    /* cube corners against the floor (box-vs-plane manifold: the corners within the breaking threshold) */
    m3 Rx; q2m(Rx, e->xq);
    if (m->cube) for (int k = 0; k < 8; k++, slot++) {
        e->cand_active[slot] = 0;
        v3 cl = {(k & 1) ? m->cube_half : -m->cube_half, (k & 2) ? m->cube_half : -m->cube_half, (k & 4) ? m->cube_half : -m->cube_half};
        v3 c; m3mulv(c, Rx, cl); v3add(c, c, e->xp);
        if (c[2] < m->cube_threshold) {
            contact *ct = &all[n++];
            ct->ga = -1; ct->gb = -1; ct->la = LINK_CUBE; ct->lb = LINK_FLOOR; ct->slot = slot;
            v3set(ct->n, 0, 0, 1); v3cpy(ct->pa, c); v3set(ct->pb, c[0], c[1], 0.0);
            ct->dist = c[2]; ct->mu = m->cube_friction * m->ground_friction; ct->mu_spin = 0; ct->mu_roll = 0;
        }
    }
    int nground = n;
    for (int pi = 0; pi < m->npair; pi++, slot++) {
        e->cand_active[slot] = 0;
        int ga = m->pair_a[pi], gb = m->pair_b[pi];
        v3 a0, a1, b0, b1, ca, cb, d;
        geom_world(e, ga, a0, a1); geom_world(e, gb, b0, b1);
        closest_seg_seg(a0, a1, b0, b1, ca, cb);
        v3sub(d, ca, cb);
        double len = v3norm(d), dist = len - m->g_radius[ga] - m->g_radius[gb];
        double thr = m->g_threshold[ga] < m->g_threshold[gb] ? m->g_threshold[ga] : m->g_threshold[gb];
        if (dist < thr && len > 1e-9) {
            contact *ct = &all[n++];
            ct->ga = ga; ct->gb = gb; ct->la = m->g_link[ga]; ct->lb = m->g_link[gb]; ct->slot = slot;
            for (int k = 0; k < 3; k++) ct->n[k] = d[k] / len;
            for (int k = 0; k < 3; k++) { ct->pa[k] = ca[k] - ct->n[k] * m->g_radius[ga]; ct->pb[k] = cb[k] + ct->n[k] * m->g_radius[gb]; }
            ct->dist = dist; ct->mu = m->g_friction[ga] * m->g_friction[gb];
            ct->mu_spin = m->torsional ? m->g_spin[ga] * m->g_friction[gb] + m->g_spin[gb] * m->g_friction[ga] : 0;
            ct->mu_roll = m->torsional ? m->g_roll[ga] * m->g_friction[gb] + m->g_roll[gb] * m->g_friction[ga] : 0;
        }
    }
    /* robot geoms against the cube (default URDF filter: group 1, mask all -> every contype-1 geom collides) */
    if (m->cube) for (int g = 0; g < m->ng; g++, slot++) {
        e->cand_active[slot] = 0;
        v3 a, b, al, bl, t, cs, cb; geom_world(e, g, a, b);
        m3 Rxt; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) Rxt[i][j] = Rx[j][i];
        v3sub(t, a, e->xp); m3mulv(al, Rxt, t); v3sub(t, b, e->xp); m3mulv(bl, Rxt, t);
        closest_seg_box(al, bl, m->cube_half, cs, cb);
        v3 dl, d; v3sub(dl, cs, cb); m3mulv(d, Rx, dl);
        double len = v3norm(d), dist = len - m->g_radius[g];
        double thr = m->g_threshold[g] < m->cube_threshold ? m->g_threshold[g] : m->cube_threshold;
        /* a capsule axis that passes through the box has len = 0 up to the bisection's resolution (1e-10 here, 2e-8 in the
         * kernel's float32): no witness direction, no contact.  The cut-off sits above both resolutions. */
        if (dist < thr && len > 1e-6) {
            contact *ct = &all[n++];
            ct->ga = g; ct->gb = -1; ct->la = m->g_link[g]; ct->lb = LINK_CUBE; ct->slot = slot;
            v3 csw, cbw; m3mulv(csw, Rx, cs); v3add(csw, csw, e->xp); m3mulv(cbw, Rx, cb); v3add(cbw, cbw, e->xp);
            for (int k = 0; k < 3; k++) ct->n[k] = d[k] / len;
            for (int k = 0; k < 3; k++) { ct->pa[k] = csw[k] - ct->n[k] * m->g_radius[g]; ct->pb[k] = cbw[k]; }
            ct->dist = dist; ct->mu = m->g_friction[g] * m->cube_friction; ct->mu_spin = 0; ct->mu_roll = 0;
        }
    }
    (void)nground;
    /* what getContactPoints would report for the feet (rs/robot_bases.py:280-281, rs/gym_locomotion_envs.py:73) follows the
     * threshold test, not the solver's row budget: taken before the cap */
    for (int f = 0; f < m->nfeet; f++) {
        e->feet_touch[f] = 0.0;
        for (int i = 0; i < n; i++) if (all[i].lb == LINK_FLOOR && all[i].la == m->foot_link[f]) e->feet_touch[f] = 1.0;
    }
    /* cap: keep the max_contacts deepest, preserving candidate order */
    int cap = m->max_contacts > 0 ? m->max_contacts : MAXCAND;
    if (e->max_rows > 0) {
        /* row budget (the CUDA kernels of the humanoid kinds hold at most max_rows constraint rows): the violated joint limits
         * of this sub-step come first, every contact needs three rows */
        int nlim = 0;
        for (int d = 0; d < e->nd; d++) {
            int l = e->link_of_dof[d];
            if (!(m->lower[l] <= m->upper[l])) continue;
            if (e->q[d] - m->lower[l] <= 0) nlim++;
            if (m->upper[l] - e->q[d] <= 0) nlim++;
        }
        int room = (e->max_rows - nlim) / 3;
        if (room < 0) room = 0;
        if (room < cap) cap = room;
    }
    if (n > cap) e->cap_overflow++;
    while (n > cap) {
        int w = 0;
        for (int i = 1; i < n; i++) if (all[i].dist > all[w].dist) w = i;   /* first shallowest... */
        /* ties: drop the later candidate */
        for (int i = 0; i < n; i++) if (all[i].dist == all[w].dist) w = i;
        for (int i = w; i < n - 1; i++) all[i] = all[i + 1];
        n--;
    }
    for (int i = 0; i < n; i++) { e->ct[i] = all[i]; e->cand_active[all[i].slot] = 1; }
    e->nct = n;
    int nslot = slot;
    for (int s = 0; s < nslot; s++) if (!e->cand_active[s]) e->warm[s] = 0.0;
}

/* ------------------------------------------------------------------ one substep */
static void clamp_u(orc_env *e) {
    double mv = e->m.max_coord_vel;
    if (e->m.floating) for (int k = 0; k < 3; k++) {
        if (e->bomega[k] > mv) e->bomega[k] = mv; if (e->bomega[k] < -mv) e->bomega[k] = -mv;
        if (e->bvel[k] > mv) e->bvel[k] = mv; if (e->bvel[k] < -mv) e->bvel[k] = -mv;
    }
    for (int d = 0; d < e->nd; d++) { if (e->qd[d] > mv) e->qd[d] = mv; if (e->qd[d] < -mv) e->qd[d] = -mv; }
    if (e->m.cube) for (int k = 0; k < 3; k++) {
        if (e->xw[k] > mv) e->xw[k] = mv; if (e->xw[k] < -mv) e->xw[k] = -mv;
        if (e->xv[k] > mv) e->xv[k] = mv; if (e->xv[k] < -mv) e->xv[k] = -mv;
    }
}
static void get_u(const orc_env *e, double *u) {
    int o = 0;
    if (e->m.floating) { for (int k = 0; k < 3; k++) { u[k] = e->bomega[k]; u[3 + k] = e->bvel[k]; } o = 6; }
    for (int d = 0; d < e->nd; d++) u[o + d] = e->qd[d];
    if (e->m.cube) for (int k = 0; k < 3; k++) { u[e->nu + k] = e->xw[k]; u[e->nu + 3 + k] = e->xv[k]; }
}
static void add_u(orc_env *e, const double *du, double s) {
    int o = 0;
    if (e->m.floating) { for (int k = 0; k < 3; k++) { e->bomega[k] += s * du[k]; e->bvel[k] += s * du[3 + k]; } o = 6; }
    for (int d = 0; d < e->nd; d++) e->qd[d] += s * du[o + d];
    if (e->m.cube) for (int k = 0; k < 3; k++) { e->xw[k] += s * du[e->nu + k]; e->xv[k] += s * du[e->nu + 3 + k]; }
    clamp_u(e);
}

static void forward_dynamics(orc_env *e, double h) {
    const orc_model *m = &e->m;
    static __thread sv pl[MAXL], cl[MAXL];
    double kd = m->link_damping;
    for (int i = 0; i < m->nl; i++) {
        /* velocity-product acceleration and bias force */
        sv Sq; int d = e->dof_of_link[i];
        double qd = d >= 0 ? e->qd[d] : 0.0;
        for (int k = 0; k < 6; k++) Sq[k] = e->S[i][k] * qd;
        sv_crm(cl[i], e->V[i], Sq);
        sm I; spatial_inertia(e, i, I);
        sv IV; sm_mulv(IV, I, e->V[i]);
        sv_crf(pl[i], e->V[i], IV);
        /* external: gravity + Bullet's per-link linear/angular damping (C3.2), applied at the link COM */
        double ms = m->mass[i];
        if (ms > 0 || m->inertia[3 * i] > 0) {
            v3 vc; link_com_vel(e, i, vc);
            double vn = v3norm(vc);
            v3 f = {0, 0, -ms * m->gravity};
            v3axpy(f, -ms * (kd + kd * vn), vc);
            /* angular damping torque: -I_local w_local (k + k |w|) */
            v3 wl, tl, tw;
            m3 Rt; for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) Rt[a][b] = e->R[i][b][a];
            m3mulv(wl, Rt, e->V[i]);
            double wn = v3norm(wl);
            for (int k = 0; k < 3; k++) tl[k] = -m->inertia[3 * i + k] * wl[k] * (kd + kd * wn);
            m3mulv(tw, (double(*)[3])e->R[i], tl);
            v3 cf; v3cross(cf, e->c[i], f);
            for (int k = 0; k < 3; k++) { pl[i][k] -= tw[k] + cf[k]; pl[i][3 + k] -= f[k]; }
        }
    }
    sv a0; double qdd[MAXD];
    aba_solve(e, pl, cl, e->tau, a0, qdd);
    double du[MAXU]; int o = 0;
    if (m->floating) {
        /* classical COM acceleration of the base: a_O + wdot x c + w x v_com */
        v3 t1, t2, vc; v3cross(t1, a0, e->c[0]); link_com_vel(e, 0, vc); v3cross(t2, e->bomega, vc);
        for (int k = 0; k < 3; k++) { du[k] = a0[k]; du[3 + k] = a0[3 + k] + t1[k] + t2[k]; }
        o = 6;
    }
    for (int d = 0; d < e->nd; d++) du[o + d] = qdd[d];
    if (m->cube) {
        /* free body with isotropic inertia: gravity + Bullet's base damping; no gyroscopic term */
        double vn = v3norm(e->xv), wn = v3norm(e->xw);
        for (int k = 0; k < 3; k++) {
            du[e->nu + k] = -e->xw[k] * (kd + kd * wn);
            du[e->nu + 3 + k] = -e->xv[k] * (kd + kd * vn) - (k == 2 ? m->gravity : 0.0);
        }
    }
    add_u(e, du, h);
}

static int build_rows(orc_env *e, double h) {
    const orc_model *m = &e->m;
    int nr = 0, nu = e->nut, o = m->floating ? 6 : 0;
    double u[MAXU]; get_u(e, u);
    /* 1. joint limits (btMultiBodyJointLimitConstraint): only violated sides produce a row */
    for (int d = 0; d < e->nd; d++) {
        int l = e->link_of_dof[d];
        if (!(m->lower[l] <= m->upper[l])) continue;
        for (int side = 0; side < 2; side++) {
            double pen = side ? (m->upper[l] - e->q[d]) : (e->q[d] - m->lower[l]);
            if (pen > 0) continue;
            row *r = &e->rows[nr++];
            memset(r, 0, sizeof(row));
            double dir = side ? -1.0 : 1.0;
            r->J[o + d] = dir;
            impulse_response(e, r->J, r->U);
            double den = 0, rel = 0;
            for (int k = 0; k < nu; k++) { den += r->J[k] * r->U[k]; rel += r->J[k] * u[k]; }
            r->dinv = den > 1e-12 ? 1.0 / den : 0.0;
            double poserr = -pen * m->erp_limit / h;
            if (m->limit_split_impulse && !(pen > m->split_impulse_threshold)) poserr = 0.0;
            r->rhs = (poserr - rel) * r->dinv;
            r->lo = 0; r->hi = m->limit_max_impulse; r->fric_of = -1;
        }
    }
    int nlim = nr;
    /* 2. contact normals, 3. friction rows */
    int nrm0 = nr;
    for (int c = 0; c < e->nct; c++) {
        contact *ct = &e->ct[c];
        row *r = &e->rows[nr++];
        memset(r, 0, sizeof(row));
        point_jacobian(e, ct->la, ct->pa, ct->n, 1.0, r->J);
        if (ct->lb != LINK_FLOOR) point_jacobian(e, ct->lb, ct->pb, ct->n, -1.0, r->J);
        impulse_response(e, r->J, r->U);
        double den = 0, rel = 0;
        for (int k = 0; k < nu; k++) { den += r->J[k] * r->U[k]; rel += r->J[k] * u[k]; }
        r->dinv = den > 1e-12 ? 1.0 / den : 0.0;
        double pen = ct->dist + m->linear_slop, poserr = 0, velerr = -rel;   /* restitution 0.5*0 = 0 */
        if (pen > 0) velerr -= pen / h; else poserr = -pen * m->erp_contact / h;
        r->rhs = (poserr + velerr) * r->dinv;
        r->lo = 0; r->hi = 1e10; r->fric_of = -1;
        r->lambda = e->warm[ct->slot] * m->warmstart;
    }
    int nnrm = nr - nrm0;
    /* 2b. torsional friction: spinning about the normal, rolling about the two tangents (angular-only rows,
     * bounded by coefficient * normal impulse like the lateral rows) */
    for (int c = 0; c < e->nct && m->torsional; c++) {
        contact *ct = &e->ct[c];
        v3 t1, t2; const double *n = ct->n;
        if (fabs(n[2]) > 0.7071067811865475244) {
            double a = n[1] * n[1] + n[2] * n[2], k = 1.0 / sqrt(a);
            v3set(t1, 0, -n[2] * k, n[1] * k); v3set(t2, a * k, -n[0] * t1[2], n[0] * t1[1]);
        } else {
            double a = n[0] * n[0] + n[1] * n[1], k = 1.0 / sqrt(a);
            v3set(t1, -n[1] * k, n[0] * k, 0); v3set(t2, -n[2] * t1[1], n[2] * t1[0], a * k);
        }
        for (int ax = 0; ax < 3; ax++) {
            double coef = ax == 0 ? ct->mu_spin : ct->mu_roll;
            if (!(coef > 0)) continue;
            const double *axis = ax == 0 ? n : (ax == 1 ? t1 : t2);
            row *r = &e->rows[nr++];
            memset(r, 0, sizeof(row));
            angular_jacobian(e, ct->la, axis, 1.0, r->J);
            if (ct->lb != LINK_FLOOR) angular_jacobian(e, ct->lb, axis, -1.0, r->J);
            impulse_response(e, r->J, r->U);
            double den = 0, rel = 0;
            for (int k = 0; k < nu; k++) { den += r->J[k] * r->U[k]; rel += r->J[k] * u[k]; }
            r->dinv = den > 1e-12 ? 1.0 / den : 0.0;
            r->rhs = -rel * r->dinv;
            r->mu = coef; r->lo = -coef; r->hi = coef; r->fric_of = nrm0 + c;
        }
    }
    for (int c = 0; c < e->nct; c++) {
        contact *ct = &e->ct[c];
        /* btPlaneSpace1 */
        v3 t1, t2; const double *n = ct->n;
        if (fabs(n[2]) > 0.7071067811865475244) {
            double a = n[1] * n[1] + n[2] * n[2], k = 1.0 / sqrt(a);
            v3set(t1, 0, -n[2] * k, n[1] * k); v3set(t2, a * k, -n[0] * t1[2], n[0] * t1[1]);
        } else {
            double a = n[0] * n[0] + n[1] * n[1], k = 1.0 / sqrt(a);
            v3set(t1, -n[1] * k, n[0] * k, 0); v3set(t2, -n[2] * t1[1], n[2] * t1[0], a * k);
        }
        for (int fd = 0; fd < 2; fd++) {
            const double *t = fd ? t2 : t1;
            row *r = &e->rows[nr++];
            memset(r, 0, sizeof(row));
            point_jacobian(e, ct->la, ct->pa, t, 1.0, r->J);
            if (ct->lb != LINK_FLOOR) point_jacobian(e, ct->lb, ct->pb, t, -1.0, r->J);
            impulse_response(e, r->J, r->U);
            double den = 0, rel = 0;
            for (int k = 0; k < nu; k++) { den += r->J[k] * r->U[k]; rel += r->J[k] * u[k]; }
            r->dinv = den > 1e-12 ? 1.0 / den : 0.0;
            r->rhs = -rel * r->dinv;
            r->mu = ct->mu; r->lo = -ct->mu; r->hi = ct->mu; r->fric_of = nrm0 + c;
        }
    }
    (void)nnrm;
    return (nlim << 16) | nr;
}

static void solve_rows(orc_env *e, int packed) {
    int nlim = packed >> 16, nr = packed & 0xffff, nu = e->nut;
    e->last_nlim = nlim; e->last_nr = nr;
    int nnrm = e->nct, nrm0 = nlim, fr0 = nlim + nnrm;
    double dv[MAXU];
    for (int k = 0; k < nu; k++) dv[k] = 0;
    /* warm start */
    for (int i = 0; i < nr; i++) if (e->rows[i].lambda != 0.0) for (int k = 0; k < nu; k++) dv[k] += e->rows[i].U[k] * e->rows[i].lambda;
#define RESOLVE(r) do { \
        double jd = 0; for (int k = 0; k < nu; k++) jd += (r)->J[k] * dv[k]; \
        double dl = (r)->rhs - jd * (r)->dinv, sum = (r)->lambda + dl; \
        if (sum < (r)->lo) { dl = (r)->lo - (r)->lambda; (r)->lambda = (r)->lo; } \
        else if (sum > (r)->hi) { dl = (r)->hi - (r)->lambda; (r)->lambda = (r)->hi; } \
        else (r)->lambda = sum; \
        for (int k = 0; k < nu; k++) dv[k] += (r)->U[k] * dl; } while (0)
    for (int it = 0; it < e->m.niter; it++) {
        for (int j = 0; j < nlim; j++) { int idx = (it & 1) ? j : nlim - 1 - j; row *r = &e->rows[idx]; RESOLVE(r); }
        for (int j = 0; j < nnrm; j++) { row *r = &e->rows[nrm0 + j]; RESOLVE(r); }
        if (e->m.friction_cone && !e->m.torsional) {
            for (int j = fr0; j + 1 < nr; j += 2) {
                row *ra = &e->rows[j], *rb = &e->rows[j + 1];
                double lim = ra->mu * e->rows[ra->fric_of].lambda;
                double ja = 0, jb = 0;
                for (int k = 0; k < nu; k++) { ja += ra->J[k] * dv[k]; jb += rb->J[k] * dv[k]; }
                double sa = ra->lambda + (ra->rhs - ja * ra->dinv), sb = rb->lambda + (rb->rhs - jb * rb->dinv);
                if (sa * sa + sb * sb >= lim * lim) {
                    double nn = sqrt(sa * sa + sb * sb), sc = nn > 0 ? (lim > 0 ? lim : 0) / nn : 0;
                    sa *= sc; sb *= sc;
                }
                double da = sa - ra->lambda, db = sb - rb->lambda;
                ra->lambda = sa; rb->lambda = sb;
                for (int k = 0; k < nu; k++) dv[k] += ra->U[k] * da + rb->U[k] * db;
            }
        } else
        for (int j = fr0; j < nr; j++) {
            row *r = &e->rows[j];
            double tot = e->rows[r->fric_of].lambda;
            if (tot > 0) { r->lo = -r->mu * tot; r->hi = r->mu * tot; RESOLVE(r); }
        }
    }
#undef RESOLVE
    for (int k = 0; k < nu; k++) e->last_dv[k] = dv[k];
    add_u(e, dv, 1.0);
    for (int c = 0; c < e->nct; c++) e->warm[e->ct[c].slot] = e->rows[nrm0 + c].lambda;
}

/* btMultiBody::stepPositionsMultiDof quaternion update (exponential map of omega*h) */
static void integrate_quat(double quat[4], const double omega[3], double h) {
    double w = v3norm(omega), ax[3];
    if (w * h > 0.25 * M_PI) w = 0.5 * (0.5 * M_PI) / h;
    double sc;
    if (w < 0.001) sc = 0.5 * h - h * h * h * 0.020833333333 * w * w; else sc = sin(0.5 * w * h) / w;
    for (int k = 0; k < 3; k++) ax[k] = omega[k] * sc;
    double dq[4] = {ax[0], ax[1], ax[2], cos(0.5 * w * h)}, qn[4];
    qmul(qn, dq, quat);
    double nn = sqrt(qn[0] * qn[0] + qn[1] * qn[1] + qn[2] * qn[2] + qn[3] * qn[3]);
    for (int k = 0; k < 4; k++) quat[k] = qn[k] / nn;
}

static void integrate(orc_env *e, double h) {
    if (e->m.floating) {
        for (int k = 0; k < 3; k++) e->bpos[k] += e->bvel[k] * h;
        integrate_quat(e->bquat, e->bomega, h);
    }
    if (e->m.cube) {
        for (int k = 0; k < 3; k++) e->xp[k] += e->xv[k] * h;
        integrate_quat(e->xq, e->xw, h);
    }
    for (int d = 0; d < e->nd; d++) e->q[d] += e->qd[d] * h;
}

static void substep(orc_env *e) {
    double h = e->m.dt_sub;
    fk(e);
    velocities(e);
    collide(e);
    articulated_inertias(e);
    get_u(e, e->last_u0);
    forward_dynamics(e, h);
    get_u(e, e->last_ufree);
    velocities(e);
    int packed = build_rows(e, h);
    solve_rows(e, packed);
    integrate(e, h);
}

void orc_physics_step(orc_env *e, const double *action) {
    const orc_model *m = &e->m;
    for (int d = 0; d < e->nd; d++) {
        int l = e->link_of_dof[d];
        e->tau[d] = -m->damping[l] * e->qd[d];     /* C3.3: joint damping torque, set once per stepSimulation */
    }
    for (int n = 0; n < m->nact; n++) {
        double a = action[n]; a = a < -1 ? -1 : (a > 1 ? 1 : a);
        e->tau[e->dof_of_link[m->act_link[n]]] += m->act_torque[n] * a;
    }
    for (int s = 0; s < m->nsub; s++) substep(e);
    for (int d = 0; d < e->nd; d++) e->tau[d] = 0;
}

/* ------------------------------------------------------------------ task layer */
static void euler_from_quat(const double q[4], double rpy[3]) {
    /* pybullet getEulerFromQuaternion (rs/robot_bases.py:216-217), SURVEY A1 */
    double x = q[0], y = q[1], z = q[2], w = q[3];
    double sarg = -2.0 * (x * z - w * y);
    rpy[0] = atan2(2 * (y * z + w * x), w * w - x * x - y * y + z * z);
    rpy[1] = sarg <= -1.0 ? -0.5 * M_PI : (sarg >= 1.0 ? 0.5 * M_PI : asin(sarg));
    rpy[2] = atan2(2 * (x * y + w * z), w * w + x * x - y * y - z * z);
}
static void mat_to_quat(m3 m, double q[4]) {
    double t = m[0][0] + m[1][1] + m[2][2];
    if (t > 0) { double s = sqrt(t + 1.0) * 2; q[3] = 0.25 * s; q[0] = (m[2][1] - m[1][2]) / s; q[1] = (m[0][2] - m[2][0]) / s; q[2] = (m[1][0] - m[0][1]) / s; }
    else {
        int i = 0; if (m[1][1] > m[0][0]) i = 1; if (m[2][2] > m[i][i]) i = 2;
        int j = (i + 1) % 3, k = (i + 2) % 3;
        double s = sqrt(m[i][i] - m[j][j] - m[k][k] + 1.0) * 2;
        q[i] = 0.25 * s; q[j] = (m[j][i] + m[i][j]) / s; q[k] = (m[k][i] + m[i][k]) / s; q[3] = (m[k][j] - m[j][k]) / s;
    }
}

/* WalkerBase robots; the MuJoCo-style Ant / Humanoid (pybulletgym/envs/mujoco/robot_locomotors.py:222-319) run the same
 * WalkerBase.calc_state and only re-pack the observation */
static int is_mjfloat(int kind) { return kind == ORC_KIND_ANT_MJ || kind == ORC_KIND_HUMANOID_MJ; }
static int is_walker(int kind) { return (kind >= ORC_KIND_HOPPER && kind <= ORC_KIND_FLAGRUN_HARDER) || is_mjfloat(kind); }
static double potential_leak(const orc_env *e);

static void dmarg(orc_env *e, double v) { v = fabs(v); if (v < e->done_margin) e->done_margin = v; }

static double alive_bonus(orc_env *e, double z, double pitch) {
    switch (e->m.kind) {
    case ORC_KIND_HOPPER: case ORC_KIND_WALKER2D: dmarg(e, z - 0.8); dmarg(e, fabs(pitch) - 1.0); break;
    case ORC_KIND_HALFCHEETAH: dmarg(e, fabs(pitch) - 1.0); break;
    case ORC_KIND_ANT: case ORC_KIND_ANT_MJ: dmarg(e, z - 0.26); break;
    case ORC_KIND_FLAGRUN_HARDER: dmarg(e, z - 0.8); break;
    default: dmarg(e, z - 0.78); break;
    }
    switch (e->m.kind) {
    case ORC_KIND_HOPPER: case ORC_KIND_WALKER2D: return (z > 0.8 && fabs(pitch) < 1.0) ? 1 : -1;
    case ORC_KIND_HALFCHEETAH:
        return (fabs(pitch) < 1.0 && !e->feet_contact[1] && !e->feet_contact[2] && !e->feet_contact[4] && !e->feet_contact[5]) ? 1 : -1;
    case ORC_KIND_ANT: case ORC_KIND_ANT_MJ: return z > 0.26 ? 1 : -1;
    case ORC_KIND_FLAGRUN_HARDER:
        /* rs/robot_locomotors.py:250-273: every 30 frames after frame 100 the cube is thrown at the spot the
         * robot will be at when it arrives */
        if (e->frame % 30 == 0 && e->frame > 100 && e->on_ground == 0) {
            double angle = draw(e, 2, 5u * e->attacks, -3.14f, 3.14f), speed = draw(e, 2, 5u * e->attacks + 1u, 20.f, 30.f);
            double ttt = 4.0 / speed, tgt[3], pos[3], vel[3], nrm = 0;
            for (int k = 0; k < 3; k++) tgt[k] = e->body_xyz[k] + e->torso_speed[k] * ttt;
            pos[0] = tgt[0] + 4.0 * cos(angle); pos[1] = tgt[1] + 4.0 * sin(angle); pos[2] = tgt[2] + 1.0;
            for (int k = 0; k < 3; k++) { vel[k] = tgt[k] - pos[k]; nrm += vel[k] * vel[k]; }
            nrm = sqrt(nrm);
            for (int k = 0; k < 3; k++) vel[k] = vel[k] * (speed / nrm) + draw(e, 2, 5u * e->attacks + 2u + k, -1.f, 1.f);
            e->attacks++;
            if (e->m.cube) for (int k = 0; k < 3; k++) { e->xp[k] = pos[k]; e->xv[k] = vel[k]; e->xw[k] = 0; }   /* orientation kept */
        }
        if (z < 0.8) e->on_ground++; else if (e->on_ground > 0) e->on_ground--;
        e->frame++;
        return e->on_ground < 170 ? potential_leak(e) : -1;
    default: return z > 0.78 ? 2 : -1;
    }
}

static int is_flagrun(int kind) { return kind == ORC_KIND_FLAGRUN || kind == ORC_KIND_FLAGRUN_HARDER; }

/* HumanoidFlagrun.flag_reposition (rs/robot_locomotors.py:204-218); float32 arithmetic like the device */
static void flag_reposition(orc_env *e) {
    const orc_model *m = &e->m;
    int taped = e->tape_pos < e->tape_n;
    double x = draw(e, 1, 2u * e->flag_count, -(float)m->stadium_halflen, (float)m->stadium_halflen);
    double y = draw(e, 1, 2u * e->flag_count + 1u, -(float)m->stadium_halfwidth, (float)m->stadium_halfwidth);
    if (taped) { e->walk_target_x = 0.5 * x; e->walk_target_y = 0.5 * y; }      /* reference: float64 */
    else { e->walk_target_x = (double)(0.5f * (float)x); e->walk_target_y = (double)(0.5f * (float)y); }
    e->flag_count++;
    e->flag_timeout = 600.0 / m->nsub;
}
static double potential_leak(const orc_env *e) {
    double z = e->body_xyz[2]; z = z < 0 ? 0 : (z > 0.8 ? 0.8 : z);
    return z / 0.8 + 1.0;
}
/* HumanoidFlagrunHarder.calc_potential (rs/robot_locomotors.py:280-302): has side effects */
static double harder_potential(orc_env *e) {
    double frp = -e->walk_target_dist / e->m.dt_scene;
    if (e->body_xyz[2] < 0.8) {
        if (!e->crawl_has) { e->crawl_start = frp - e->crawl_ign; e->crawl_has = 1; }
        e->crawl_ign = frp - e->crawl_start;
        frp = e->crawl_start;
    } else { frp -= e->crawl_ign; e->crawl_has = 0; }
    return frp + potential_leak(e) * 100.0;
}

/* WalkerBase.calc_state (rs/robot_locomotors.py:31-64) on the current physics state */
static void walker_calc_state(orc_env *e, double *obs) {
    const orc_model *m = &e->m;
    fk(e); velocities(e);
    int nA = m->nact;
    float j[2 * MAXD];
    e->joints_at_limit = 0;
    for (int n = 0; n < nA; n++) {
        int l = m->act_link[n], d = e->dof_of_link[l];
        double pos = e->q[d], vel = e->qd[d];
        if (m->lower[l] < m->upper[l]) pos = 2 * (pos - 0.5 * (m->lower[l] + m->upper[l])) / (m->upper[l] - m->lower[l]);
        vel *= (m->jtype[l] == ORC_JT_REVOLUTE) ? 0.1 : 0.5;
        j[2 * n] = (float)pos; j[2 * n + 1] = (float)vel;
        e->joint_speeds[n] = j[2 * n + 1];
        if (fabsf(j[2 * n]) > 0.99f) e->joints_at_limit++;
    }
    double sx = 0, sy = 0; int np = 0;
    for (int i = 0; i < m->nl; i++) if (m->in_parts[i]) { sx += e->c[i][0]; sy += e->c[i][1]; np++; }
    if (e->floor_in_parts) np++;        /* quirk Q1: the floor (0,0,0) is averaged in */
    int tl = m->torso_link;
    e->body_xyz[0] = sx / np; e->body_xyz[1] = sy / np; e->body_xyz[2] = e->c[tl][2];
    double tq[4]; mat_to_quat((double(*)[3])e->R[tl], tq);
    euler_from_quat(tq, e->body_rpy);
    double z = e->body_xyz[2];
    if (!e->have_initial_z) { e->initial_z = z; e->have_initial_z = 1; }
    double yaw = e->body_rpy[2];
    double ty = e->walk_target_y - e->body_xyz[1], tx = e->walk_target_x - e->body_xyz[0];
    double theta = atan2(ty, tx);
    e->walk_target_dist = sqrt(ty * ty + tx * tx);
    if (is_flagrun(m->kind)) {
        /* HumanoidFlagrun.calc_state (rs/robot_locomotors.py:220-227) */
        e->flag_timeout -= 1;
        if (e->walk_target_dist < 1 || e->flag_timeout <= 0) {
            flag_reposition(e);
            ty = e->walk_target_y - e->body_xyz[1]; tx = e->walk_target_x - e->body_xyz[0];
            theta = atan2(ty, tx);
            e->walk_target_dist = sqrt(ty * ty + tx * tx);
            if (m->kind == ORC_KIND_FLAGRUN_HARDER) (void)harder_potential(e);   /* robot.potential: unused by the env (Q5) */
        }
    }
    double ang = theta - yaw;
    v3 sp; link_com_vel(e, tl, sp);
    v3cpy(e->torso_speed, sp);
    double cy = cos(-yaw), sy_ = sin(-yaw);
    double vx = cy * sp[0] - sy_ * sp[1], vy = sy_ * sp[0] + cy * sp[1], vz = sp[2];
    float more[8] = {(float)(z - e->initial_z), (float)sin(ang), (float)cos(ang), (float)(0.3 * vx), (float)(0.3 * vy),
                     (float)(0.3 * vz), (float)e->body_rpy[0], (float)e->body_rpy[1]};
    int o = 0;
    for (int k = 0; k < 8; k++) obs[o++] = more[k];
    for (int k = 0; k < 2 * nA; k++) obs[o++] = j[k];
    for (int k = 0; k < m->nfeet; k++) obs[o++] = (float)e->feet_contact[k];
    for (int k = 0; k < o; k++) { if (obs[k] < -5) obs[k] = -5; if (obs[k] > 5) obs[k] = 5; }
}

/* Reacher.calc_state (rs/robot_manipulators.py:28-47): dofs joint0, joint1, target_x, target_y */
static void reacher_calc_state(orc_env *e, double *obs) {
    const orc_model *m = &e->m;
    fk(e);
    double theta = e->q[0], theta_dot = 0.1 * e->qd[0];        /* joint0 is unlimited: raw position */
    int l1 = e->link_of_dof[1];
    double gamma = 2 * (e->q[1] - 0.5 * (m->lower[l1] + m->upper[l1])) / (m->upper[l1] - m->lower[l1]), gamma_dot = 0.1 * e->qd[1];
    const double *ft = e->c[m->aux_link[0]], *tg = e->c[m->aux_link[1]];
    for (int k = 0; k < 3; k++) e->body_xyz[k] = ft[k] - tg[k];       /* to_target_vec */
    e->joint_speeds[0] = theta_dot; e->joint_speeds[1] = gamma_dot; e->body_rpy[0] = gamma;
    obs[0] = e->q[2]; obs[1] = e->q[3]; obs[2] = e->body_xyz[0]; obs[3] = e->body_xyz[1];
    obs[4] = cos(theta); obs[5] = sin(theta); obs[6] = theta_dot; obs[7] = gamma; obs[8] = gamma_dot;
}
static double reacher_potential(const orc_env *e) { return -100.0 * v3norm(e->body_xyz); }

/* MuJoCo-style Hopper / Walker2D (pybulletgym/envs/mujoco/robot_locomotors.py:86-165): qpos[1:] ++ clip(qvel, +-10) over
 * all dofs (root joints included), float32 storage */
static int is_mjwalker(int kind) { return kind == ORC_KIND_HOPPER_MJ || kind == ORC_KIND_WALKER2D_MJ || kind == ORC_KIND_HALFCHEETAH_MJ; }
static void mjwalker_calc_state(orc_env *e, double *obs) {
    int nd = e->nd, o = 0;
    for (int k = 1; k < nd; k++) obs[o++] = (double)(float)e->q[k];
    /* mujoco/robot_locomotors.py:183-190: the HalfCheetah concatenates qvel unclipped */
    int clipv = e->m.kind != ORC_KIND_HALFCHEETAH_MJ;
    for (int k = 0; k < nd; k++) { float v = (float)e->qd[k]; if (clipv) v = v < -10.f ? -10.f : (v > 10.f ? 10.f : v); obs[o++] = v; }
}
static double mjwalker_body_x(orc_env *e) { fk(e); return e->c[e->m.torso_link][0]; }   /* robot_body.get_pose()[0] */

/* observation of a WalkerBase robot.  MuJoCo-style Ant / Humanoid (pybulletgym/envs/mujoco/robot_locomotors.py:222-319): run
 * WalkerBase.calc_state for its side effects, then qpos[2:] ++ qvel ++ zero padding (cinert, cvel, qfrc_actuator, cfrc_ext
 * are "TODO: FIND" zeros in the reference); float64 throughout */
static void walker_obs(orc_env *e, double *obs) {
    const orc_model *m = &e->m;
    if (!is_mjfloat(m->kind)) { walker_calc_state(e, obs); return; }
    double tmp[64]; walker_calc_state(e, tmp);
    int o = 0, nA = m->nact;
    obs[o++] = e->bpos[2];
    for (int k = 0; k < 4; k++) obs[o++] = e->bquat[k];
    for (int n = 0; n < nA; n++) obs[o++] = e->q[e->dof_of_link[m->act_link[n]]];
    for (int k = 0; k < 3; k++) obs[o++] = e->bvel[k];
    for (int k = 0; k < 3; k++) obs[o++] = e->bomega[k];
    for (int n = 0; n < nA; n++) obs[o++] = e->qd[e->dof_of_link[m->act_link[n]]];
    while (o < m->obs_dim) obs[o++] = 0.0;
}

static void pendulum_calc_state(orc_env *e, double *obs) {
    /* rs/robot_pendula.py:27-51: slider = dof 0, hinge = dof 1 */
    double x = e->q[0], vx = e->qd[0], th = e->q[1], thd = e->qd[1];
    if (e->m.kind == ORC_KIND_DOUBLE_PENDULUM || e->m.kind == ORC_KIND_DOUBLE_PENDULUM_MJ) {
        /* InvertedDoublePendulum.calc_state (rs/robot_pendula.py:75-87): pole2 = last link, its COM is the
         * middle of the second pole (rs/gym_pendulum_envs.py:73-74) */
        fk(e);
        double ga = e->q[2], gad = e->qd[2];
        e->body_xyz[0] = e->c[e->m.nl - 1][0]; e->body_xyz[2] = e->c[e->m.nl - 1][2];
        if (e->m.kind == ORC_KIND_DOUBLE_PENDULUM_MJ) {
            /* pybulletgym/envs/mujoco/robot_pendula.py:73-88: [x, sin, sin, cos, cos, clip(qvel, +-10), qfrc_constraint = 0] */
            double v[3] = {vx, thd, gad};
            obs[0] = x; obs[1] = sin(th); obs[2] = sin(ga); obs[3] = cos(th); obs[4] = cos(ga);
            for (int k = 0; k < 3; k++) { obs[5 + k] = v[k] < -10 ? -10 : (v[k] > 10 ? 10 : v[k]); obs[8 + k] = 0; }
            return;
        }
        obs[0] = x; obs[1] = vx; obs[2] = e->body_xyz[0]; obs[3] = cos(th); obs[4] = sin(th); obs[5] = thd;
        obs[6] = cos(ga); obs[7] = sin(ga); obs[8] = gad;
        return;
    }
    obs[0] = x; obs[1] = vx; obs[2] = cos(th); obs[3] = sin(th); obs[4] = thd;
}

static double calc_potential(orc_env *e) {
    if (e->m.kind == ORC_KIND_FLAGRUN_HARDER) return harder_potential(e);
    return -e->walk_target_dist / e->m.dt_scene;
}

static void update_feet_contact(orc_env *e) {
    const orc_model *m = &e->m;
    for (int f = 0; f < m->nfeet; f++) e->feet_contact[f] = e->feet_touch[f];
}

int orc_observe(orc_env *e, const double *a, double *obs, double *reward, double *terms) {
    const orc_model *m = &e->m;
    double t5[5] = {0, 0, 0, 0, 0};
    int done = 0;
    e->done_margin = 1e30;
    if (is_mjwalker(m->kind)) {
        /* HopperMuJoCoEnv._step / Walker2DMuJoCoEnv._step (pybulletgym/envs/mujoco/gym_locomotion_envs.py:121-206) */
        double x = mjwalker_body_x(e);
        t5[0] = (x - e->potential) / m->dt_scene;          /* calc_potential(): (pos_after - pos_before) / dt */
        e->potential = x;
        t5[1] = 1.0;
        double ss = 0; for (int n = 0; n < m->nact; n++) ss += a[n] * a[n];
        t5[2] = -1e-3 * ss;
        mjwalker_calc_state(e, obs);
        if (m->kind == ORC_KIND_HALFCHEETAH_MJ) {
            /* HalfCheetahMuJoCoEnv._step (mujoco/gym_locomotion_envs.py:216-244): rewards = [potential, power_cost], done = False */
            t5[1] = -0.1 * ss; t5[2] = 0.0;
            *reward = t5[0] + t5[1];
            if (terms) memcpy(terms, t5, sizeof(t5));
            return 0;
        }
        double height = obs[0], ang = obs[1];
        int ok = 1;
        for (int k = 0; k < m->obs_dim; k++) if (!isfinite(obs[k])) ok = 0;
        for (int k = 2; k < m->obs_dim; k++) if (!(fabs(obs[k]) < 100)) ok = 0;
        if (m->kind == ORC_KIND_HOPPER_MJ) { ok = ok && height > -0.3 && fabs(ang) < 0.2; dmarg(e, height + 0.3); dmarg(e, fabs(ang) - 0.2); }
        else { ok = ok && 1.0 > height && height > -0.2 && -1.0 < ang && ang < 1.0; dmarg(e, height - 1.0); dmarg(e, height + 0.2); dmarg(e, fabs(ang) - 1.0); }
        for (int k = 2; k < m->obs_dim; k++) dmarg(e, fabs(obs[k]) - 100);
        done = !ok;
        *reward = t5[0] + t5[1] + t5[2];
    } else if (m->kind == ORC_KIND_REACHER) {
        /* ReacherBulletEnv._step (rs/gym_manipulator_envs.py:15-32): never done */
        reacher_calc_state(e, obs);
        double pold = e->potential;
        e->potential = reacher_potential(e);
        double td = e->joint_speeds[0], gd = e->joint_speeds[1], gamma = e->body_rpy[0];
        t5[0] = e->potential - pold;
        t5[1] = -0.10 * (fabs(a[0] * td) + fabs(a[1] * gd)) - 0.01 * (fabs(a[0]) + fabs(a[1]));
        t5[2] = fabs(fabs(gamma) - 1) < 0.01 ? -0.1 : 0.0;
        *reward = t5[0] + t5[1] + t5[2];
    } else if (!is_walker(m->kind)) {
        pendulum_calc_state(e, obs);
        double th = e->q[1];
        if (m->kind == ORC_KIND_DOUBLE_PENDULUM || m->kind == ORC_KIND_DOUBLE_PENDULUM_MJ) {
            /* InvertedDoublePendulumBulletEnv._step (rs/gym_pendulum_envs.py:69-83); the MuJoCo-style variant
             * (pybulletgym/envs/mujoco/gym_pendulum_envs.py:56-69) adds the velocity penalty */
            double px = e->body_xyz[0], py = e->body_xyz[2];
            t5[0] = 10.0; t5[1] = -(0.01 * px * px + (py + 0.3 - 2) * (py + 0.3 - 2)); t5[2] = -0.0;
            if (m->kind == ORC_KIND_DOUBLE_PENDULUM_MJ) t5[2] = -(1e-3 * e->qd[1] * e->qd[1] + 5e-3 * e->qd[2] * e->qd[2]);
            done = py + 0.3 <= 1; dmarg(e, py + 0.3 - 1);
            *reward = t5[0] + t5[1] + t5[2];
        } else {
        if (m->kind == ORC_KIND_PENDULUM_SWINGUP) { t5[0] = cos(th); done = 0; }
        else { t5[0] = 1.0; done = fabs(th) > 0.2; dmarg(e, fabs(th) - 0.2); }
        *reward = t5[0];
        }
    } else {
        double zz;
        walker_obs(e, obs);
        if (is_mjfloat(m->kind)) zz = obs[0] + e->initial_z;   /* WalkerBaseMuJoCoEnv._step: state[0] = torso z stands in for z - initial_z */
        else if (m->initial_z >= 0) zz = (double)((float)obs[0] + (float)e->initial_z);   /* np.float32 + python float */
        else zz = obs[0] + e->initial_z;                                            /* np.float32 + np.float64 */
        double alive = alive_bonus(e, zz, e->body_rpy[1]);
        done = alive < 0;
        for (int k = 0; k < m->obs_dim; k++) if (!isfinite(obs[k])) done = 1;
        double pold = e->potential;
        e->potential = calc_potential(e);
        double progress = e->potential - pold;
        update_feet_contact(e);                   /* quirk Q2: after calc_state / alive */
        double se = 0, ss = 0;
        for (int n = 0; n < m->nact; n++) { se += fabs(a[n] * e->joint_speeds[n]); ss += a[n] * a[n]; }
        double elec = m->elec_cost * (se / m->nact) + m->stall_cost * (ss / m->nact);
        double lim = m->limit_cost * e->joints_at_limit;
        t5[0] = alive; t5[1] = progress; t5[2] = elec; t5[3] = lim; t5[4] = 0.0;
        if (is_mjfloat(m->kind)) { t5[2] = lim; t5[3] = 0.0; }     /* [alive, progress, joints_at_limit_cost, feet_collision_cost] */
        *reward = t5[0] + t5[1] + t5[2] + t5[3] + t5[4];
    }
    if (terms) memcpy(terms, t5, sizeof(t5));
    return done;
}

int orc_step(orc_env *e, const double *action, double *obs, double *reward, double *terms) {
    orc_physics_step(e, action);
    e->steps++;
    return orc_observe(e, action, obs, reward, terms);
}

static void reset_common(orc_env *e, const double *noise, int floor_in_parts, double *obs) {
    const orc_model *m = &e->m;
    memset(e->q, 0, sizeof(e->q)); memset(e->qd, 0, sizeof(e->qd)); memset(e->tau, 0, sizeof(e->tau));
    memset(e->warm, 0, sizeof(e->warm));
    memset(e->mf, 0, sizeof(e->mf));   /* removeBody / restoreState: the manifolds start empty */
    if (m->floating) {
        /* snapshot pose = MJCF pose: base inertial frame */
        m3 R0; q2m(R0, &m->quat[0]);
        v3 t; m3mulv(t, R0, &m->com[0]);
        for (int k = 0; k < 3; k++) { e->bpos[k] = m->pos[k] + t[k]; e->bomega[k] = 0; e->bvel[k] = 0; }
        for (int k = 0; k < 4; k++) e->bquat[k] = m->quat[k];
    }
    if (is_walker(m->kind)) {
        for (int n = 0; n < m->nact; n++) e->q[e->dof_of_link[m->act_link[n]]] = noise[n];
    } else if (is_mjwalker(m->kind)) {
        for (int n = 0; n < e->nd; n++) e->q[n] = noise[n];       /* every ordered joint, root joints included */
    } else if (m->kind == ORC_KIND_REACHER) {
        /* rs/robot_manipulators.py:12-21: target_x, target_y, joint0, joint1 in this draw order */
        e->q[2] = noise[0]; e->q[3] = noise[1]; e->q[0] = noise[2]; e->q[1] = noise[3];
    } else {
        e->q[1] = noise[0] + (m->kind == ORC_KIND_PENDULUM_SWINGUP ? 3.1415 : 0.0);
        if (m->kind == ORC_KIND_DOUBLE_PENDULUM || m->kind == ORC_KIND_DOUBLE_PENDULUM_MJ) e->q[2] = noise[1];       /* rs/robot_pendula.py:66-68 */
    }
    if (m->cube) {
        /* restoreState + resetBasePositionAndOrientation(cube, [-1.5,0,0.05], identity) (rs/robot_locomotors.py:240-243) */
        for (int k = 0; k < 3; k++) { e->xp[k] = m->cube_pos0[k]; e->xw[k] = 0; e->xv[k] = 0; e->xq[k] = 0; }
        e->xq[3] = 1.0;
    }
    for (int f = 0; f < MAXFEET; f++) { e->feet_contact[f] = 0; e->feet_touch[f] = 0; }
    e->steps = 0; e->nct = 0;
    e->floor_in_parts = floor_in_parts;
    e->walk_target_x = m->walk_target_x; e->walk_target_y = m->walk_target_y;
    if (m->initial_z >= 0) { e->initial_z = m->initial_z; e->have_initial_z = 1; } else e->have_initial_z = 0;
    e->flag_count = 0; e->flag_timeout = 0; e->frame = 0; e->on_ground = 0; e->crawl_has = 0; e->crawl_start = 0; e->crawl_ign = 0; e->attacks = 0;
    if (is_flagrun(m->kind)) flag_reposition(e);
    if (is_walker(m->kind)) { walker_obs(e, obs); e->potential = calc_potential(e); }
    else if (m->kind == ORC_KIND_REACHER) { reacher_calc_state(e, obs); e->potential = reacher_potential(e); }
    else if (is_mjwalker(m->kind)) { mjwalker_calc_state(e, obs); e->potential = mjwalker_body_x(e); }
    else pendulum_calc_state(e, obs);
    /* quirk Q1: the env adds the floor to robot.parts right after this first calc_state
     * (rs/gym_locomotion_envs.py:30-31), so every later calc_state of the episode averages it in */
    e->floor_in_parts = 1;
}

void orc_reset_with(orc_env *e, const double *noise, int floor_in_parts, double *obs) {
    e->episode++;
    reset_common(e, noise, floor_in_parts, obs);
}

void orc_reset(orc_env *e, int floor_in_parts, double *obs) {
    double noise[MAXD];
    e->episode++;
    int n = is_walker(e->m.kind) ? e->m.nact : ((e->m.kind == ORC_KIND_DOUBLE_PENDULUM || e->m.kind == ORC_KIND_DOUBLE_PENDULUM_MJ) ? 2 : 1);
    for (int k = 0; k < n; k++) noise[k] = rng_uniform(e->seed, e->env_index, e->episode, (uint32_t)k, -0.1f, 0.1f);
    if (is_mjwalker(e->m.kind))
        for (int k = 0; k < e->nd; k++) noise[k] = rng_uniform(e->seed, e->env_index, e->episode, (uint32_t)k, -0.1f, 0.1f);
    if (e->m.kind == ORC_KIND_REACHER)
        for (int k = 0; k < 4; k++) { float r = k < 2 ? 0.27f : 3.14f; noise[k] = rng_uniform(e->seed, e->env_index, e->episode, (uint32_t)k, -r, r); }
    reset_common(e, noise, floor_in_parts, obs);
}

/* physics step driven by raw joint torques (dof order), as setJointMotorControl2(TORQUE_CONTROL)
 * would: used by the fake-pybullet backend that generates the task-layer golden vectors */
void orc_physics_step_torque(orc_env *e, const double *tau) {
    const orc_model *m = &e->m;
    for (int d = 0; d < e->nd; d++) e->tau[d] = tau[d] - m->damping[e->link_of_dof[d]] * e->qd[d];
    for (int s = 0; s < m->nsub; s++) substep(e);
    for (int d = 0; d < e->nd; d++) e->tau[d] = 0;
}
/* per link: COM position(3), orientation quaternion xyzw(4), COM linear velocity(3) */
void orc_link_state(orc_env *e, double *out) {
    fk(e); velocities(e);
    for (int i = 0; i < e->m.nl; i++) {
        double q[4]; mat_to_quat((double(*)[3])e->R[i], q);
        v3 v; link_com_vel(e, i, v);
        for (int k = 0; k < 3; k++) out[10 * i + k] = e->c[i][k];
        for (int k = 0; k < 4; k++) out[10 * i + 3 + k] = q[k];
        for (int k = 0; k < 3; k++) out[10 * i + 7 + k] = v[k];
    }
}
/* contact points of the last substep: link A, link B (-1 = floor), distance */
int orc_get_contacts(const orc_env *e, int32_t *la, int32_t *lb, double *dist) {
    for (int c = 0; c < e->nct; c++) { la[c] = e->ct[c].la; lb[c] = e->ct[c].lb; dist[c] = e->ct[c].dist; }
    return e->nct;
}
void orc_set_joint(orc_env *e, int dof, double q, double qd) { e->q[dof] = q; e->qd[dof] = qd; }
void orc_get_joint(const orc_env *e, int dof, double *q, double *qd) { *q = e->q[dof]; *qd = e->qd[dof]; }
void orc_set_cube(orc_env *e, const double *pos, const double *quat, const double *omega, const double *vel) {
    for (int k = 0; k < 3; k++) { if (pos) e->xp[k] = pos[k]; if (omega) e->xw[k] = omega[k]; if (vel) e->xv[k] = vel[k]; }
    if (quat) for (int k = 0; k < 4; k++) e->xq[k] = quat[k];
}
void orc_get_cube(const orc_env *e, double *pos, double *quat, double *omega, double *vel) {
    for (int k = 0; k < 3; k++) { if (pos) pos[k] = e->xp[k]; if (omega) omega[k] = e->xw[k]; if (vel) vel[k] = e->xv[k]; }
    if (quat) for (int k = 0; k < 4; k++) quat[k] = e->xq[k];
}

/* rows of the last substep: out = [nlim, nct, then per row rhs, dinv, lambda]; returns the row count */
int orc_get_rows(const orc_env *e, double *out) {
    int nlim = e->last_nlim, nr = e->last_nr;
    out[0] = nlim; out[1] = e->nct;
    for (int i = 0; i < nr; i++) { out[2 + 3 * i] = e->rows[i].rhs; out[3 + 3 * i] = e->rows[i].dinv; out[4 + 3 * i] = e->rows[i].lambda; }
    return nr;
}
void orc_get_debug(const orc_env *e, double *out) { for (int k = 0; k < e->nu; k++) { out[k] = e->last_u0[k]; out[64 + k] = e->last_ufree[k]; out[128 + k] = e->last_dv[k]; } }
int orc_num_contacts(const orc_env *e) { return e->nct; }
void orc_feet_contact(const orc_env *e, double *out) { for (int f = 0; f < e->m.nfeet; f++) out[f] = e->feet_contact[f]; }
void orc_link_com(orc_env *e, double *out) { fk(e); for (int i = 0; i < e->m.nl; i++) for (int k = 0; k < 3; k++) out[3 * i + k] = e->c[i][k]; }

double orc_energy(orc_env *e) {
    fk(e); velocities(e);
    double E = 0;
    for (int i = 0; i < e->m.nl; i++) {
        sm I; spatial_inertia(e, i, I);
        sv IV; sm_mulv(IV, I, e->V[i]);
        E += 0.5 * sv_dot(e->V[i], IV) + e->m.mass[i] * e->m.gravity * e->c[i][2];
    }
    return E;
}

void orc_mass_matrix_inv(orc_env *e, double *Minv) {
    fk(e); velocities(e); articulated_inertias(e);
    int nu = e->nu;
    for (int k = 0; k < nu; k++) {
        double f[MAXU], du[MAXU];
        for (int i = 0; i < MAXU; i++) f[i] = 0;
        f[k] = 1.0;
        impulse_response(e, f, du);
        for (int i = 0; i < nu; i++) Minv[i * nu + k] = du[i];
    }
}

long orc_rollout(orc_env *e, long steps, uint64_t action_seed, double *ret_sum, long *episodes) {
    double obs[512], act[MAXD], rew, rs = 0; long eps = 0;
    orc_reset(e, 1, obs);
    for (long t = 0; t < steps; t++) {
        for (int n = 0; n < e->m.nact; n++)
            act[n] = rng_uniform(action_seed, e->env_index, (uint32_t)t, (uint32_t)n, -1.0f, 1.0f);
        int done = orc_step(e, act, obs, &rew, NULL);
        rs += rew;
        if (done || e->steps >= e->m.max_episode_steps) { eps++; orc_reset(e, 1, obs); }
    }
    if (ret_sum) *ret_sum = rs;
    if (episodes) *episodes = eps;
    return steps;
}

/* Whole random-policy episodes for the statistical parity tier (T4): `n` episodes in a row (each: reset from the counter RNG,
 * U(-1,1) actions keyed by (action_seed, env_index, episode * 4096 + t, n), until done or `cap` steps); per episode the return
 * and the length.  Returns the number of env steps taken. */
long orc_episodes(orc_env *e, int n, int cap, uint64_t action_seed, double *returns, int32_t *lengths) {
    double obs[512], act[MAXD], rew; long total = 0;
    for (int ep = 0; ep < n; ep++) {
        orc_reset(e, 1, obs);
        double rs = 0; int t = 0;
        for (; t < cap; ) {
            for (int k = 0; k < e->m.nact; k++)
                act[k] = rng_uniform(action_seed, e->env_index, (uint32_t)(ep * 4096 + t), (uint32_t)k, -1.0f, 1.0f);
            int done = orc_step(e, act, obs, &rew, NULL);
            rs += rew; t++;
            if (done) break;
        }
        returns[ep] = rs; lengths[ep] = t; total += t;
    }
    return total;
}

int orc_cap_overflows(const orc_env *e) { return e->cap_overflow; }
void orc_set_max_rows(orc_env *e, int max_rows) { e->max_rows = max_rows; }
double orc_done_margin(const orc_env *e) { return e->done_margin; }
void orc_feet_margin(const orc_env *e, double *out) { for (int f = 0; f < e->m.nfeet; f++) out[f] = e->feet_margin[f]; }
