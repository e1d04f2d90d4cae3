// C ABI of libpbg_b200.so (see include/pbg.h): host-side model flattening, state allocation and
// kernel dispatch.  One handle = one env kind x num_envs worlds on one device.
#include "../../include/pbg.h"
#include "pbg_cfgs.cuh"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

using namespace pbg;

static bool kernel_for_kind(int kind, KernelInfo *out) {
    switch (kind) {
    case PBG_KIND_PENDULUM: case PBG_KIND_PENDULUM_SWINGUP: *out = info_Pendulum(); return true;
    case PBG_KIND_DOUBLE_PENDULUM: *out = info_DoublePendulum(); return true;
    case PBG_KIND_DOUBLE_PENDULUM_MJ: *out = info_DoublePendulumMJ(); return true;
    case PBG_KIND_REACHER: *out = info_Reacher(); return true;
    case PBG_KIND_HOPPER: *out = info_Hopper(); return true;
    case PBG_KIND_WALKER2D: *out = info_Walker(); return true;
    case PBG_KIND_HOPPER_MJ: *out = info_HopperMJ(); return true;
    case PBG_KIND_WALKER2D_MJ: *out = info_WalkerMJ(); return true;
    case PBG_KIND_HALFCHEETAH: *out = info_Cheetah(); return true;
    case PBG_KIND_HALFCHEETAH_MJ: *out = info_CheetahMJ(); return true;
    case PBG_KIND_ANT: *out = info_Ant(); return true;
    case PBG_KIND_ANT_MJ: *out = info_AntMJ(); return true;
    case PBG_KIND_HUMANOID_MJ: *out = info_HumanoidMJ(); return true;
    case PBG_KIND_HUMANOID: case PBG_KIND_FLAGRUN: *out = info_Humanoid(); return true;
    case PBG_KIND_FLAGRUN_HARDER: *out = info_Harder(); return true;
    default: return false;
    }
}

struct pbg_handle {
    int device = 0;
    int E = 0;
    int kind = 0;
    KernelInfo k{};
    DevModel *dmodel = nullptr;
    float *state = nullptr;
    unsigned long long *stats = nullptr;
    // staging for pbg_step_host
    float *d_act = nullptr, *d_obs = nullptr, *d_rew = nullptr;
    uint8_t *d_done = nullptr;
    cudaStream_t hstream = nullptr;
    unsigned long long seed = 0, env_offset = 0;
    int auto_reset = 1;
    int debug_env = 0;
    float *policy_buf = nullptr;   // fused-policy weights (pbg_set_policy)
    PolicyDev policy{};
    int policy_tc = 0;          // pbg_set_policy_tensor_cores
    int zero_copy = 1;          // pbg_step_host: let the kernel read / write mapped pinned host buffers directly
    int last_host_path = 0;     // 1: zero-copy, 2: staged copies
    int64_t launches = 0;
    // stream-ordering bookkeeping: the stream of the last stream-ordered call.  pbg_step_host (private stream) and
    // pbg_stats (blocking copy) order themselves after it; see order_after_last()
    cudaStream_t last_stream = nullptr;
    bool have_last = false;
    cudaEvent_t order_ev = nullptr;
    bool ready = false;         // a reset / set_state / restore has initialised the state
    float *d_cand = nullptr;    // [E, nslot] contact-candidate distances of the last step (pbg_enable_contact_export)
    std::string err;
};

static thread_local std::string g_create_err;

static int fail(pbg_handle *h, int code, const std::string &msg) {
    if (h) h->err = msg; else g_create_err = msg;
    return code;
}
#define CUDA_TRY(h, expr)                                                                             \
    do {                                                                                              \
        cudaError_t _e = (expr);                                                                      \
        if (_e != cudaSuccess)                                                                        \
            return fail(h, PBG_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));         \
    } while (0)

static void quat_to_mat(const double *q, float *R) {
    double x = q[0], y = q[1], z = q[2], w = q[3];
    double n = std::sqrt(x * x + y * y + z * z + w * w);
    x /= n; y /= n; z /= n; w /= n;
    R[0] = float(1 - 2 * (y * y + z * z)); R[1] = float(2 * (x * y - w * z)); R[2] = float(2 * (x * z + w * y));
    R[3] = float(2 * (x * y + w * z)); R[4] = float(1 - 2 * (x * x + z * z)); R[5] = float(2 * (y * z - w * x));
    R[6] = float(2 * (x * z - w * y)); R[7] = float(2 * (y * z + w * x)); R[8] = float(1 - 2 * (x * x + y * y));
}

// flatten the caller's tables into the device model; returns "" or an error message
static std::string build_dev_model(const pbg_model *pm, const KernelInfo &k, DevModel *d) {
    char buf[256];
    memset(d, 0, sizeof(DevModel));
    if (pm->nb != k.nb - k.hasx || pm->nj != k.nj || pm->floating != k.floating) {
        snprintf(buf, sizeof buf, "model has nb=%d nj=%d floating=%d, kernel for kind %d expects %d/%d/%d", pm->nb, pm->nj,
                 pm->floating, pm->kind, k.nb - k.hasx, k.nj, k.floating);
        return buf;
    }
    if ((pm->torsional_friction != 0) != (k.tors != 0)) return "the kernel of this env kind and the model disagree about torsional friction rows";
    if (k.tors && (!pm->geom_spin || !pm->geom_roll)) return "torsional friction needs geom_spin / geom_roll";
    if ((pm->cube != 0) != (k.hasx != 0)) return "the kernel of this env kind and the model disagree about the cube";
    if (pm->action_dim != k.nact || pm->obs_dim != k.obs || pm->nfeet != k.nfeet) return "action/obs/feet dims do not match the kernel";
    if (pm->ns > MSUB) return "too many Bullet links";
    if (pm->max_contacts != k.maxc) {
        snprintf(buf, sizeof buf, "max_contacts=%d but the kernel keeps %d", pm->max_contacts, k.maxc);
        return buf;
    }
    const int F = pm->floating ? 6 : 0;
    d->nb = pm->nb + k.hasx; d->nj = pm->nj; d->nd = pm->nj + F + 6 * k.hasx; d->floating = pm->floating;
    d->nact = pm->action_dim; d->nfeet = pm->nfeet; d->obs_dim = pm->obs_dim; d->kind = pm->kind;
    int jidx = 0, nlim = 0;
    for (int b = 0; b < pm->nb; ++b) {
        d->parent[b] = pm->parent[b]; d->jtype[b] = pm->jtype[b];
        if (pm->parent[b] >= b) return "bodies are not in depth-first order";
        d->depth[b] = pm->parent[b] < 0 ? 0 : d->depth[pm->parent[b]] + 1;
        if (d->depth[b] > d->maxdepth) d->maxdepth = d->depth[b];
        quat_to_mat(pm->q0 + 4 * b, d->q0m[b]);
        if (k.q0id) {
            const double *q = pm->q0 + 4 * b;
            if (std::fabs(q[0]) > 1e-12 || std::fabs(q[1]) > 1e-12 || std::fabs(q[2]) > 1e-12 || std::fabs(std::fabs(q[3]) - 1.0) > 1e-12)
                return "this env kind's kernel assumes identity rest rotations (q0) for every body";
        }
        for (int i = 0; i < 3; ++i) {
            d->anchor_p[b][i] = (float)pm->anchor_p[3 * b + i];
            d->com_off[b][i] = (float)pm->com_off[3 * b + i];
        }
        if (k.q0id == 2) {
            // planar kinds (KCfg::PLANAR): hinges about +-y, slides in the xz plane, unit axes, fixed base
            const double *ax = pm->axis + 3 * b;
            const bool ok = pm->jtype[b] == PBG_JT_REVOLUTE
                                ? (std::fabs(ax[0]) < 1e-12 && std::fabs(ax[2]) < 1e-12 && std::fabs(std::fabs(ax[1]) - 1.0) < 1e-12)
                                : (pm->jtype[b] == PBG_JT_PRISMATIC && std::fabs(ax[1]) < 1e-12 &&
                                   std::fabs(ax[0] * ax[0] + ax[2] * ax[2] - 1.0) < 1e-12);
            if (!ok || pm->floating) return "this env kind's kernel assumes a planar robot: fixed base, unit hinge axes along +-y, unit slide axes in the xz plane";
        }
        {
            // Bullet keeps a non-unit MJCF joint axis as written: rotation about its direction, motion subspace with its length
            const double *ax = pm->axis + 3 * b;
            const double len = std::sqrt(ax[0] * ax[0] + ax[1] * ax[1] + ax[2] * ax[2]);
            d->axis_len[b] = len > 0 ? (float)len : 1.f;
            for (int i = 0; i < 3; ++i) d->axis[b][i] = len > 0 ? (float)(ax[i] / len) : 0.f;
        }
        d->mass[b] = (float)pm->mass[b];
        for (int i = 0; i < 6; ++i) d->inertia[b][i] = (float)pm->inertia[6 * b + i];
        if (pm->jtype[b] == PBG_JT_FREE) {
            if (b != 0) return "floating joint on a non-root body";
            d->dof[b] = 0;
            for (int i = 0; i < 3; ++i) d->base_pos0[i] = (float)pm->anchor_p[i];
            for (int i = 0; i < 4; ++i) d->base_quat0[i] = (float)pm->q0[i];
        } else if (pm->jtype[b] == PBG_JT_REVOLUTE || pm->jtype[b] == PBG_JT_PRISMATIC) {
            d->dof[b] = F + jidx;
            d->jbody[jidx] = b;
            d->jrev[jidx] = pm->jtype[b] == PBG_JT_REVOLUTE;
            d->jlo[jidx] = (float)pm->jnt_lower[jidx]; d->jhi[jidx] = (float)pm->jnt_upper[jidx];
            d->jlimited[jidx] = pm->jnt_lower[jidx] < pm->jnt_upper[jidx];
            nlim += d->jlimited[jidx];
            d->jdamp[jidx] = (float)pm->jnt_damping[jidx];
            d->jact[jidx] = pm->jnt_act[jidx];
            d->jtorque[jidx] = (float)pm->jnt_torque[jidx];
            if (pm->jnt_act[jidx] >= 0) {
                if (pm->jnt_act[jidx] >= pm->action_dim) return "action index out of range";
                d->act_joint[pm->jnt_act[jidx]] = jidx;
            }
            ++jidx;
        } else return "unsupported joint type in the reduced tree";
    }
    if (jidx != pm->nj) return "joint count mismatch";
    if (nlim > k.nlim) return "more limited joints than the kernel has limit rows";
    d->nlim = nlim;
    // dof relations
    for (int b = 0; b < pm->nb; ++b) {
        unsigned mask = 0;
        for (int a = b; a >= 0; a = pm->parent[a]) {
            if (pm->jtype[a] == PBG_JT_FREE) mask |= 0x3fu;
            else mask |= 1u << d->dof[a];
        }
        d->anc[b] = mask;
    }
    const int XD0 = pm->nj + F, XB = pm->nb;       // first cube dof / the cube's body index
    for (int kk = 0; kk < XD0; ++kk) {
        const int body = (pm->floating && kk < 6) ? 0 : d->jbody[kk - F];
        d->up[kk] = d->anc[body];
    }
    if (k.hasx) {
        // the cube: a second root.  Inertia of the collision box (isotropic), identity rest frame.
        d->parent[XB] = -1; d->jtype[XB] = 4; d->depth[XB] = 0; d->dof[XB] = XD0;
        d->q0m[XB][0] = d->q0m[XB][4] = d->q0m[XB][8] = 1.f;
        d->mass[XB] = (float)pm->cube_mass;
        d->inertia[XB][0] = d->inertia[XB][1] = d->inertia[XB][2] = (float)pm->cube_inertia;
        d->anc[XB] = 0x3fu << XD0;
        for (int c = 0; c < 6; ++c) d->up[XD0 + c] = d->anc[XB];
        for (int i = 0; i < 3; ++i) d->cube_pos0[i] = (float)pm->cube_pos0[i];
    }
    // the kernel's factorisation skips structural zeros by a compile-time topology: it must be this model's
    for (int kk = 0; kk < d->nd; ++kk)
        if ((d->up[kk] & ((1u << kk) - 1u)) != k.low[kk]) return "the model's kinematic tree does not match the kernel's compile-time topology";
    for (int kk = 0; kk < d->nd; ++kk) {
        unsigned dn = 0;
        for (int l = 0; l < d->nd; ++l) if ((d->up[l] >> kk) & 1u) dn |= 1u << l;
        d->down[kk] = dn;
    }
    // Bullet links: parts membership and damping entries grouped by body
    std::vector<std::vector<int>> by_body(d->nb);
    for (int s = 0; s < pm->ns; ++s) {
        const int b = pm->sub_body[s];
        if (b < 0 || b >= pm->nb) return "sub_body out of range";
        if (pm->sub_in_parts[s]) {
            d->part_cnt[b] += 1.f;
            for (int i = 0; i < 3; ++i) d->part_sum[b][i] += (float)pm->sub_off[3 * s + i];
        }
        if (pm->sub_mass[s] > 0) by_body[b].push_back(s);
    }
    int nds = 0;
    for (int b = 0; b < pm->nb; ++b) {
        d->ds_begin[b] = nds;
        for (int s : by_body[b]) {
            for (int i = 0; i < 3; ++i) { d->ds_off[nds][i] = (float)pm->sub_off[3 * s + i]; d->ds_inertia[nds][i] = (float)pm->sub_inertia[3 * s + i]; }
            d->ds_mass[nds] = (float)pm->sub_mass[s];
            ++nds;
        }
    }
    if (k.hasx) {
        if (nds >= MSUB) return "too many damping entries";
        d->ds_begin[XB] = nds;
        d->ds_mass[nds] = (float)pm->cube_mass;
        for (int i = 0; i < 3; ++i) d->ds_inertia[nds][i] = (float)pm->cube_inertia;
        ++nds;
    }
    for (int b = d->nb; b <= MB; ++b) d->ds_begin[b] = nds;
    if (pm->torso_sub < 0 || pm->torso_sub >= pm->ns) return "torso_sub out of range";
    d->torso_body = pm->sub_body[pm->torso_sub];
    for (int i = 0; i < 3; ++i) d->torso_off[i] = (float)pm->sub_off[3 * pm->torso_sub + i];
    for (int a = 0; a < 2; ++a) {
        d->aux_body[a] = 0;
        if (pm->aux_sub[a] >= 0) {
            if (pm->aux_sub[a] >= pm->ns) return "aux_sub out of range";
            d->aux_body[a] = pm->sub_body[pm->aux_sub[a]];
            for (int i = 0; i < 3; ++i) d->aux_off[a][i] = (float)pm->sub_off[3 * pm->aux_sub[a] + i];
        }
    }
    if (pm->floating) {
        // the canonical base state is the Bullet base-link COM; the kernel integrates the merged root body
        double o = 0;
        for (int i = 0; i < 3; ++i) o += std::fabs(pm->sub_off[i]);
        if (o > 1e-9) return "root body COM differs from the Bullet base-link COM (not supported)";
    }
    // contact candidates: sphere -> centre, capsule -> both end spheres
    int nc = 0;
    std::vector<int> first_cand(pm->ng, -1);
    for (int g = 0; g < pm->ng; ++g) {
        if (pm->geom_type[g] == PBG_G_BOX) return "box geoms are not supported by this kernel";
        const int npt = pm->geom_type[g] == PBG_G_CAPSULE ? 2 : 1;
        if (!pm->geom_ground[g]) continue;
        first_cand[g] = nc;
        for (int e = 0; e < npt; ++e) {
            if (nc >= MCAND || nc >= k.ncand) return "too many ground contact candidates for the kernel";
            d->c_body[nc] = pm->geom_body[g];
            d->c_foot[nc] = pm->geom_foot[g];
            const double *p = e ? pm->geom_p1 + 3 * g : pm->geom_p0 + 3 * g;
            for (int i = 0; i < 3; ++i) d->c_p[nc][i] = (float)p[i];
            d->c_rad[nc] = (float)pm->geom_radius[g];
            d->c_thr[nc] = (float)pm->geom_threshold[g];
            d->c_mu[nc] = (float)(pm->geom_friction[g] * pm->ground_friction);
            if (k.tors) {
                // btManifoldResult::calculateCombinedSpinning/RollingFriction: spinA * fricB + spinB * fricA
                d->c_spin[nc] = (float)(pm->geom_spin[g] * pm->ground_friction + pm->ground_spinning_friction * pm->geom_friction[g]);
                d->c_roll[nc] = (float)(pm->geom_roll[g] * pm->ground_friction + pm->ground_rolling_friction * pm->geom_friction[g]);
                if (!(d->c_spin[nc] > 0.f) || !(d->c_roll[nc] > 0.f)) return "torsional friction rows need positive spinning and rolling coefficients on every ground contact";
            }
            ++nc;
        }
    }
    if (k.hasx) {
        // cube corners against the floor, slots right after the robot's candidates (oracle.c collide)
        for (int c = 0; c < 8; ++c) {
            if (nc >= MCAND || nc >= k.ncand) return "too many ground contact candidates for the kernel";
            d->c_body[nc] = XB; d->c_foot[nc] = -1;
            for (int i = 0; i < 3; ++i) d->c_p[nc][i] = (float)(((c >> i) & 1) ? pm->cube_half : -pm->cube_half);
            d->c_rad[nc] = 0.f;
            d->c_thr[nc] = (float)pm->cube_threshold;
            d->c_mu[nc] = (float)(pm->cube_friction * pm->ground_friction);
            ++nc;
        }
    }
    d->ncand = nc;
    const int nxp = k.hasx ? pm->ng : 0;
    if (pm->npair + nxp > k.npair || pm->npair + nxp > MPAIR) return "too many geom pairs for the kernel";
    d->npair = pm->npair + nxp;
    d->ng = 0;
    if (d->npair > 0) {
        if (pm->ng > MGEOM) return "too many collision geoms for the pair tables";
        d->ng = pm->ng;
        for (int g = 0; g < pm->ng; ++g) {
            d->g_body[g] = pm->geom_body[g];
            for (int i = 0; i < 3; ++i) { d->g_p0[g][i] = (float)pm->geom_p0[3 * g + i]; d->g_p1[g][i] = (float)pm->geom_p1[3 * g + i]; }
        }
    }
    for (int g = 0; g < nxp; ++g) {
        const int p = pm->npair + g;
        d->p_ga[p] = g; d->p_gb[p] = -1;
        d->p_ba[p] = pm->geom_body[g]; d->p_bb[p] = XB; d->p_box[p] = 1;
        d->p_half[p] = (float)pm->cube_half;
        d->p_ra[p] = (float)pm->geom_radius[g]; d->p_rb[p] = 0.f;
        d->p_thr[p] = (float)std::fmin(pm->geom_threshold[g], pm->cube_threshold);
        d->p_mu[p] = (float)(pm->geom_friction[g] * pm->cube_friction);
    }
    for (int p = 0; p < pm->npair; ++p) {
        const int a = pm->pair_a[p], b = pm->pair_b[p];
        d->p_ga[p] = a; d->p_gb[p] = b;
        d->p_ba[p] = pm->geom_body[a]; d->p_bb[p] = pm->geom_body[b];
        d->p_ra[p] = (float)pm->geom_radius[a]; d->p_rb[p] = (float)pm->geom_radius[b];
        d->p_thr[p] = (float)std::fmin(pm->geom_threshold[a], pm->geom_threshold[b]);
        d->p_mu[p] = (float)(pm->geom_friction[a] * pm->geom_friction[b]);
    }
    d->gravity = (float)pm->gravity; d->h = (float)pm->timestep; d->erp_contact = (float)pm->contact_erp;
    d->erp_limit = (float)pm->erp; d->slop = (float)pm->linear_slop; d->warm = (float)pm->warmstarting_factor;
    d->kdamp = (float)pm->link_damping; d->maxvel = (float)pm->max_coordinate_velocity;
    d->limit_max_imp = (float)pm->limit_max_impulse; d->split_thr = (float)pm->split_impulse_threshold;
    d->nsub = pm->frame_skip; d->niter = pm->num_solver_iterations; d->limit_split = pm->limit_split_impulse;
    d->max_contacts = pm->max_contacts;
    d->initial_z = (float)pm->initial_z; d->elec_cost = (float)pm->electricity_cost; d->stall_cost = (float)pm->stall_torque_cost;
    d->limit_cost = (float)pm->joints_at_limit_cost; d->walk_tx = (float)pm->walk_target_x; d->walk_ty = (float)pm->walk_target_y;
    d->halflen = (float)pm->stadium_halflen; d->halfwidth = (float)pm->stadium_halfwidth;
    d->dt_scene = pm->timestep * pm->frame_skip;
    d->inv_dt_scene = 1.0 / d->dt_scene;
    d->max_steps = pm->max_episode_steps;
    return "";
}

extern "C" {

int pbg_version(void) { return PBG_VERSION; }

// development / documentation helper: {shared memory per CTA, envs per CTA, threads per CTA, state floats per env}
int pbg_dev_kernel_geometry(int kind, int32_t *out4) {
    KernelInfo k;
    if (!kernel_for_kind(kind, &k)) return PBG_ERR_UNSUPPORTED;
    out4[0] = (int32_t)k.smem; out4[1] = k.epb; out4[2] = k.threads; out4[3] = k.sstride;
    return PBG_OK;
}

int pbg_max_contacts(int kind) {
    KernelInfo k;
    if (!kernel_for_kind(kind, &k)) return PBG_ERR_UNSUPPORTED;
    return k.maxc;
}

int pbg_max_rows(int kind) {
    KernelInfo k;
    if (!kernel_for_kind(kind, &k)) return PBG_ERR_UNSUPPORTED;
    return k.maxr;
}

const char *pbg_last_error(const pbg_handle *h) { return h ? h->err.c_str() : g_create_err.c_str(); }

int pbg_create(const pbg_model *model, int32_t num_envs, int32_t device, uint64_t seed, uint64_t env_offset, pbg_handle **out) {
    if (!model || !out || num_envs <= 0) return fail(nullptr, PBG_ERR_INVALID, "pbg_create: bad arguments");
    KernelInfo k;
    if (!kernel_for_kind(model->kind, &k)) return fail(nullptr, PBG_ERR_UNSUPPORTED, "pbg_create: no kernel for this env kind");
    DevModel hm;
    std::string e = build_dev_model(model, k, &hm);
    if (!e.empty()) return fail(nullptr, PBG_ERR_INVALID, "pbg_create: " + e);
    pbg_handle *h = new pbg_handle();
    h->device = device; h->E = num_envs; h->kind = model->kind; h->k = k; h->seed = seed; h->env_offset = env_offset;
#define CREATE_TRY(expr)                                                                  \
    do {                                                                                  \
        cudaError_t _e = (expr);                                                          \
        if (_e != cudaSuccess) {                                                          \
            g_create_err = std::string(#expr) + ": " + cudaGetErrorString(_e);            \
            pbg_destroy(h);                                                               \
            return PBG_ERR_CUDA;                                                          \
        }                                                                                 \
    } while (0)
    CREATE_TRY(cudaSetDevice(device));
    CREATE_TRY(k.prepare());
    CREATE_TRY(cudaMalloc(&h->dmodel, sizeof(DevModel)));
    CREATE_TRY(cudaMemcpy(h->dmodel, &hm, sizeof(DevModel), cudaMemcpyHostToDevice));
    const size_t sbytes = size_t(num_envs) * k.sstride * sizeof(float);
    CREATE_TRY(cudaMalloc(&h->state, sbytes));
    CREATE_TRY(cudaMemset(h->state, 0, sbytes));
    CREATE_TRY(cudaMalloc(&h->stats, 8 * sizeof(unsigned long long)));
    CREATE_TRY(cudaMemset(h->stats, 0, 8 * sizeof(unsigned long long)));
    CREATE_TRY(cudaMalloc(&h->d_act, size_t(num_envs) * k.nact * sizeof(float)));
    CREATE_TRY(cudaMalloc(&h->d_obs, size_t(num_envs) * k.obs * sizeof(float)));
    CREATE_TRY(cudaMalloc(&h->d_rew, size_t(num_envs) * sizeof(float)));
    CREATE_TRY(cudaMalloc(&h->d_done, size_t(num_envs)));
    CREATE_TRY(cudaStreamCreateWithFlags(&h->hstream, cudaStreamNonBlocking));
    CREATE_TRY(cudaEventCreateWithFlags(&h->order_ev, cudaEventDisableTiming));
#undef CREATE_TRY
    *out = h;
    return PBG_OK;
}

int pbg_destroy(pbg_handle *h) {
    if (!h) return PBG_OK;
    cudaSetDevice(h->device);
    cudaFree(h->dmodel); cudaFree(h->state); cudaFree(h->stats);
    cudaFree(h->d_act); cudaFree(h->d_obs); cudaFree(h->d_rew); cudaFree(h->d_done); cudaFree(h->policy_buf); cudaFree(h->d_cand);
    if (h->hstream) cudaStreamDestroy(h->hstream);
    if (h->order_ev) cudaEventDestroy(h->order_ev);
    delete h;
    return PBG_OK;
}

int pbg_num_envs(const pbg_handle *h) { return h ? h->E : PBG_ERR_INVALID; }
int pbg_obs_dim(const pbg_handle *h) { return h ? h->k.obs : PBG_ERR_INVALID; }
int pbg_action_dim(const pbg_handle *h) { return h ? h->k.nact : PBG_ERR_INVALID; }
int pbg_state_dim(const pbg_handle *h) { return h ? h->k.canon : PBG_ERR_INVALID; }
int pbg_noise_dim(const pbg_handle *h) { return h ? h->k.nnoise : PBG_ERR_INVALID; }
int64_t pbg_launch_count(const pbg_handle *h) { return h ? h->launches : 0; }

// Makes stream `s` wait for everything the handle's previous stream-ordered call enqueued on another stream (reset on
// torch's stream followed by pbg_step_host on the private stream, ...).  The event is recorded lazily, here, on the
// previous stream: it covers all work enqueued there so far and costs nothing on the common same-stream path.
// A stream that is being captured into a CUDA graph is left alone (recording into the graph would tie the event to it).
static int order_after_last(pbg_handle *h, cudaStream_t s) {
    if (!h->have_last || h->last_stream == s) return PBG_OK;
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(h->last_stream, &cs) != cudaSuccess) { cudaGetLastError(); return PBG_OK; }
    if (cs != cudaStreamCaptureStatusNone) return PBG_OK;
    if (cudaStreamIsCapturing(s, &cs) != cudaSuccess) { cudaGetLastError(); return PBG_OK; }
    if (cs != cudaStreamCaptureStatusNone) return PBG_OK;      // a capture may not wait on an event from outside it
    CUDA_TRY(h, cudaEventRecord(h->order_ev, h->last_stream));
    CUDA_TRY(h, cudaStreamWaitEvent(s, h->order_ev, 0));
    return PBG_OK;
}

static int launch(pbg_handle *h, int mode, StepBuffers &b, int floor_in_parts, void *stream) {
    if (!h) return PBG_ERR_INVALID;
    CUDA_TRY(h, cudaSetDevice(h->device));
    const bool needs_state = mode == MODE_STEP || mode == MODE_PHYSICS || mode == MODE_OBSERVE || mode >= 100;
    if (needs_state && !h->ready)
        return fail(h, PBG_ERR_INVALID, "the handle has no state yet: call pbg_reset / pbg_reset_with (or pbg_set_state / pbg_restore) before stepping");
    if (mode == MODE_RESET || mode == MODE_SET) h->ready = true;
    {
        int rc = order_after_last(h, (cudaStream_t)stream);
        if (rc != PBG_OK) return rc;
    }
    h->last_stream = (cudaStream_t)stream; h->have_last = true;
    b.state = h->state;
    b.stats = h->stats;
    b.cand_out = h->d_cand;
    LaunchArgs la;
    la.E = h->E; la.mode = mode; la.auto_reset = h->auto_reset; la.floor_in_parts = floor_in_parts;
    la.seed = h->seed; la.env_offset = h->env_offset; la.debug_env = h->debug_env; la.nsteps = 1;
    if (mode >= 100) { la.nsteps = mode - 100; la.mode = mode = MODE_POLICY; b.policy = h->policy; }
    h->k.launch(h->dmodel, b, la, (cudaStream_t)stream);
    CUDA_TRY(h, cudaGetLastError());
    h->launches++;
    return PBG_OK;
}

int pbg_set_auto_reset(pbg_handle *h, int32_t enabled) {
    if (!h) return PBG_ERR_INVALID;
    h->auto_reset = enabled ? 1 : 0;
    return PBG_OK;
}

int pbg_reset(pbg_handle *h, const uint8_t *mask_dev, int32_t floor_in_parts, float *obs_dev, void *stream) {
    StepBuffers b{};
    b.mask = mask_dev; b.obs = obs_dev;
    return launch(h, MODE_RESET, b, floor_in_parts, stream);
}

int pbg_reset_with(pbg_handle *h, const float *joint_noise_dev, int32_t floor_in_parts, float *obs_dev, void *stream) {
    if (!joint_noise_dev) return fail(h, PBG_ERR_INVALID, "pbg_reset_with: joint_noise_dev is NULL");
    StepBuffers b{};
    b.noise = joint_noise_dev; b.obs = obs_dev;
    return launch(h, MODE_RESET, b, floor_in_parts, stream);
}

int pbg_step(pbg_handle *h, const float *actions_dev, float *obs_dev, float *reward_dev, uint8_t *done_dev,
             float *reward_terms_dev, float *final_obs_dev, uint8_t *truncated_dev, void *stream) {
    if (!h || !actions_dev) return fail(h, PBG_ERR_INVALID, "pbg_step: actions_dev is NULL");
    StepBuffers b{};
    b.actions = actions_dev; b.obs = obs_dev; b.reward = reward_dev; b.done = done_dev; b.terms = reward_terms_dev;
    b.final_obs = final_obs_dev; b.truncated = truncated_dev;
    return launch(h, MODE_STEP, b, 1, stream);
}

// device-visible alias of a pinned (cudaHostAlloc / cudaHostRegister, UVA-mapped) host buffer, or nullptr
static void *mapped_alias(const void *p) {
    if (!p) return nullptr;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return a.type == cudaMemoryTypeHost ? a.devicePointer : nullptr;
}

int pbg_set_zero_copy(pbg_handle *h, int32_t enabled) {
    if (!h) return PBG_ERR_INVALID;
    h->zero_copy = enabled ? 1 : 0;
    return PBG_OK;
}
int pbg_last_host_path(const pbg_handle *h) { return h ? h->last_host_path : PBG_ERR_INVALID; }

int pbg_step_host(pbg_handle *h, const float *actions_host, float *obs_host, float *reward_host, uint8_t *done_host) {
    if (!h || !actions_host) return fail(h, PBG_ERR_INVALID, "pbg_step_host: actions_host is NULL");
    CUDA_TRY(h, cudaSetDevice(h->device));
    cudaStream_t s = h->hstream;
    const size_t E = h->E;
    if (!h->ready) return fail(h, PBG_ERR_INVALID, "pbg_step_host: call pbg_reset first");
    if (h->zero_copy) {
        // Pinned host buffers are mapped into the device address space (UVA): the kernel reads the 32 B of actions per env
        // over PCIe at its start and posts obs / reward / done straight into host memory at its end -- no copy
        // engine round trips, one launch and one stream synchronisation per step.
        void *da = mapped_alias(actions_host), *dobs = mapped_alias(obs_host), *drew = mapped_alias(reward_host),
             *ddone = mapped_alias(done_host);
        if (da && (dobs || !obs_host) && (drew || !reward_host) && (ddone || !done_host)) {
            int rc = pbg_step(h, (const float *)da, (float *)dobs, (float *)drew, (uint8_t *)ddone, nullptr, nullptr, nullptr, s);
            if (rc != PBG_OK) return rc;
            CUDA_TRY(h, cudaStreamSynchronize(s));
            h->last_host_path = 1;
            return PBG_OK;
        }
    }
    h->last_host_path = 2;
    CUDA_TRY(h, cudaMemcpyAsync(h->d_act, actions_host, E * h->k.nact * sizeof(float), cudaMemcpyHostToDevice, s));
    int rc = pbg_step(h, h->d_act, h->d_obs, h->d_rew, h->d_done, nullptr, nullptr, nullptr, s);
    if (rc != PBG_OK) return rc;
    if (obs_host) CUDA_TRY(h, cudaMemcpyAsync(obs_host, h->d_obs, E * h->k.obs * sizeof(float), cudaMemcpyDeviceToHost, s));
    if (reward_host) CUDA_TRY(h, cudaMemcpyAsync(reward_host, h->d_rew, E * sizeof(float), cudaMemcpyDeviceToHost, s));
    if (done_host) CUDA_TRY(h, cudaMemcpyAsync(done_host, h->d_done, E, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(h, cudaStreamSynchronize(s));
    return PBG_OK;
}

int pbg_set_policy(pbg_handle *h, int32_t h1, int32_t h2, const float *w1, const float *b1, const float *w2, const float *b2,
                   const float *w3, const float *b3) {
    if (!h || !w1 || !b1 || !w2 || !b2 || !w3 || !b3 || h1 <= 0 || h2 <= 0) return fail(h, PBG_ERR_INVALID, "pbg_set_policy: bad arguments");
    // the hidden activations live in the env's kinematics / Delassus scratch (dead between two steps)
    if (h1 + h2 > h->k.ysz) return fail(h, PBG_ERR_UNSUPPORTED, "pbg_set_policy: hidden layers do not fit this env kind's shared-memory scratch");
    CUDA_TRY(h, cudaSetDevice(h->device));
    const size_t D = h->k.obs, A = h->k.nact;
    const size_t n = D * h1 + h1 + size_t(h1) * h2 + h2 + size_t(h2) * A + A;
    cudaFree(h->policy_buf); h->policy_buf = nullptr;
    CUDA_TRY(h, cudaMalloc(&h->policy_buf, n * sizeof(float)));
    float *p = h->policy_buf;
    const float *src[6] = {w1, b1, w2, b2, w3, b3};
    const size_t cnt[6] = {D * h1, size_t(h1), size_t(h1) * h2, size_t(h2), size_t(h2) * A, A};
    const float *dst[6];
    for (int i = 0; i < 6; ++i) {
        CUDA_TRY(h, cudaMemcpy(p, src[i], cnt[i] * sizeof(float), cudaMemcpyHostToDevice));
        dst[i] = p; p += cnt[i];
    }
    h->policy = PolicyDev{dst[0], dst[1], dst[2], dst[3], dst[4], dst[5], h1, h2, h->policy_tc};
    return PBG_OK;
}

int pbg_set_policy_tensor_cores(pbg_handle *h, int32_t enabled) {
    if (!h) return PBG_ERR_INVALID;
    h->policy_tc = enabled ? 1 : 0;
    h->policy.tc = h->policy_tc;
    return PBG_OK;
}

int pbg_rollout_policy(pbg_handle *h, int32_t nsteps, float *obs_dev, float *reward_sum_dev, uint8_t *done_any_dev, void *stream) {
    if (!h || !obs_dev || nsteps <= 0 || nsteps > 100000) return fail(h, PBG_ERR_INVALID, "pbg_rollout_policy: bad arguments");
    if (!h->policy_buf) return fail(h, PBG_ERR_INVALID, "pbg_rollout_policy: no policy set (pbg_set_policy)");
    StepBuffers b{};
    b.obs = obs_dev; b.reward = reward_sum_dev; b.done = done_any_dev;
    return launch(h, 100 + nsteps, b, 1, stream);
}

int pbg_get_state(pbg_handle *h, float *state_dev, void *stream) {
    if (!h || !state_dev) return fail(h, PBG_ERR_INVALID, "pbg_get_state: NULL buffer");
    StepBuffers b{};
    b.canon = state_dev;
    return launch(h, MODE_GET, b, 1, stream);
}

int pbg_set_state(pbg_handle *h, const float *state_dev, void *stream) {
    if (!h || !state_dev) return fail(h, PBG_ERR_INVALID, "pbg_set_state: NULL buffer");
    StepBuffers b{};
    b.canon = const_cast<float *>(state_dev);
    return launch(h, MODE_SET, b, 1, stream);
}

static const uint32_t SNAP_MAGIC = 0x50424753u;   // "PBGS"
struct SnapHeader { uint32_t magic, version; int32_t kind, E, sstride, pad; unsigned long long seed, env_offset; };

int64_t pbg_snapshot_bytes(const pbg_handle *h) {
    if (!h) return PBG_ERR_INVALID;
    return int64_t(sizeof(SnapHeader)) + int64_t(h->E) * h->k.sstride * sizeof(float) + 8 * sizeof(unsigned long long);
}

int pbg_snapshot(pbg_handle *h, void *buf, void *stream) {
    if (!h || !buf) return fail(h, PBG_ERR_INVALID, "pbg_snapshot: NULL buffer");
    CUDA_TRY(h, cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    SnapHeader hd{SNAP_MAGIC, (uint32_t)pbg_version(), h->kind, h->E, h->k.sstride, 0, h->seed, h->env_offset};
    const size_t sb = size_t(h->E) * h->k.sstride * sizeof(float);
    { int rc = order_after_last(h, st); if (rc != PBG_OK) return rc; }
    h->last_stream = st; h->have_last = true;
    char *p = static_cast<char *>(buf);
    CUDA_TRY(h, cudaMemcpyAsync(p, &hd, sizeof hd, cudaMemcpyDefault, st));
    CUDA_TRY(h, cudaStreamSynchronize(st));       // hd lives on this stack frame
    CUDA_TRY(h, cudaMemcpyAsync(p + sizeof hd, h->state, sb, cudaMemcpyDefault, st));
    CUDA_TRY(h, cudaMemcpyAsync(p + sizeof hd + sb, h->stats, 8 * sizeof(unsigned long long), cudaMemcpyDefault, st));
    return PBG_OK;
}

int pbg_restore(pbg_handle *h, const void *buf, void *stream) {
    if (!h || !buf) return fail(h, PBG_ERR_INVALID, "pbg_restore: NULL buffer");
    CUDA_TRY(h, cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    SnapHeader hd{};
    const char *p = static_cast<const char *>(buf);
    CUDA_TRY(h, cudaMemcpyAsync(&hd, p, sizeof hd, cudaMemcpyDefault, st));
    CUDA_TRY(h, cudaStreamSynchronize(st));
    if (hd.magic != SNAP_MAGIC || hd.version != (uint32_t)pbg_version())
        return fail(h, PBG_ERR_INVALID, "pbg_restore: not a snapshot of this library version");
    if (hd.kind != h->kind || hd.E != h->E || hd.sstride != h->k.sstride)
        return fail(h, PBG_ERR_INVALID, "pbg_restore: snapshot of a different env kind or batch size");
    if (hd.seed != h->seed || hd.env_offset != h->env_offset)
        return fail(h, PBG_ERR_INVALID, "pbg_restore: snapshot of a handle with another seed / env_offset (its reset RNG streams differ)");
    const size_t sb = size_t(h->E) * h->k.sstride * sizeof(float);
    { int rc = order_after_last(h, st); if (rc != PBG_OK) return rc; }
    CUDA_TRY(h, cudaMemcpyAsync(h->state, p + sizeof hd, sb, cudaMemcpyDefault, st));
    CUDA_TRY(h, cudaMemcpyAsync(h->stats, p + sizeof hd + sb, 8 * sizeof(unsigned long long), cudaMemcpyDefault, st));
    h->last_stream = st; h->have_last = true; h->ready = true;
    return PBG_OK;
}

int pbg_physics_step(pbg_handle *h, const float *actions_dev, void *stream) {
    if (!h || !actions_dev) return fail(h, PBG_ERR_INVALID, "pbg_physics_step: actions_dev is NULL");
    StepBuffers b{};
    b.actions = actions_dev;
    return launch(h, MODE_PHYSICS, b, 1, stream);
}

int pbg_observe(pbg_handle *h, const float *actions_dev, float *obs_dev, float *reward_dev, uint8_t *done_dev,
                float *reward_terms_dev, void *stream) {
    if (!h || !actions_dev) return fail(h, PBG_ERR_INVALID, "pbg_observe: actions_dev is NULL");
    StepBuffers b{};
    b.actions = actions_dev; b.obs = obs_dev; b.reward = reward_dev; b.done = done_dev; b.terms = reward_terms_dev;
    return launch(h, MODE_OBSERVE, b, 1, stream);
}

// small helper kernels for diagnostics
__global__ void gather_kernel(const float *state, int sstride, int off, int n, float *out, int E) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < E * n) out[i] = state[size_t(i / n) * sstride + off + i % n];
}

int pbg_get_feet_contact(pbg_handle *h, float *out_dev, void *stream) {
    if (!h || !out_dev) return fail(h, PBG_ERR_INVALID, "pbg_get_feet_contact: NULL buffer");
    if (h->k.nfeet == 0) return PBG_OK;
    CUDA_TRY(h, cudaSetDevice(h->device));
    // feet flags are the last NFEET floats before the padded end of the state row
    const int off = h->k.off_feet;
    const int n = h->E * h->k.nfeet;
    { int rc = order_after_last(h, (cudaStream_t)stream); if (rc != PBG_OK) return rc; }
    h->last_stream = (cudaStream_t)stream; h->have_last = true;
    gather_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(h->state, h->k.sstride, off, h->k.nfeet, out_dev, h->E);
    CUDA_TRY(h, cudaGetLastError());
    h->launches++;
    return PBG_OK;
}

// development hook (not part of include/pbg.h): physics step that dumps the constraint rows of env
// `env` in its last substep: [nl, nc, then per row rhs, 1/A_ii, lambda, residual]
int pbg_dev_physics_step_rows(pbg_handle *h, const float *actions_dev, int32_t env, float *dbg_dev, void *stream) {
    if (!h || !actions_dev) return PBG_ERR_INVALID;
    StepBuffers b{};
    b.actions = actions_dev; b.debug = dbg_dev;
    h->debug_env = env;
    return launch(h, MODE_PHYSICS, b, 1, stream);
}

int pbg_physics_step_counts(pbg_handle *h, const float *actions_dev, int32_t *ncontact_dev, void *stream) {
    if (!h || !actions_dev) return fail(h, PBG_ERR_INVALID, "pbg_physics_step_counts: actions_dev is NULL");
    StepBuffers b{};
    b.actions = actions_dev; b.ncontact_out = ncontact_dev;
    return launch(h, MODE_PHYSICS, b, 1, stream);
}

// FP32 CUDA-core peak of the device the way the roofline needs it: dependent-free FFMA chains,
// 8 per thread, full occupancy.  2 flops per FFMA.
__global__ void __launch_bounds__(256) fma_peak_kernel(float *out, int iters, float seed) {
    float a0 = seed, a1 = seed + 1, a2 = seed + 2, a3 = seed + 3, a4 = seed + 4, a5 = seed + 5, a6 = seed + 6, a7 = seed + 7;
    const float b = 1.0000001f, c = 1e-9f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            a0 = fmaf(a0, b, c); a1 = fmaf(a1, b, c); a2 = fmaf(a2, b, c); a3 = fmaf(a3, b, c);
            a4 = fmaf(a4, b, c); a5 = fmaf(a5, b, c); a6 = fmaf(a6, b, c); a7 = fmaf(a7, b, c);
        }
    }
    const float r = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (r == 12345.678f) out[0] = r;
}

int pbg_measure_fp32_peak(int32_t device, double *tflops_out) {
    if (!tflops_out) return PBG_ERR_INVALID;
    if (cudaSetDevice(device) != cudaSuccess) return PBG_ERR_CUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return PBG_ERR_CUDA;
    float *buf = nullptr;
    if (cudaMalloc(&buf, 16) != cudaSuccess) return PBG_ERR_CUDA;
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 4096;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    double best = 0;
    for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(e0);
        fma_peak_kernel<<<blocks, threads>>>(buf, iters, 0.5f);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        const double flops = 2.0 * 64.0 * iters * double(blocks) * threads;
        if (rep > 0 && ms > 0) best = std::fmax(best, flops / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(buf);
    if (cudaGetLastError() != cudaSuccess) return PBG_ERR_CUDA;
    *tflops_out = best;
    return PBG_OK;
}

int pbg_set_seed(pbg_handle *h, uint64_t seed) {
    if (!h) return PBG_ERR_INVALID;
    h->seed = seed;
    return PBG_OK;
}

// task bookkeeping of every env as doubles (the potential is kept in fp64)
__global__ void task_view_kernel(const float *state, int sstride, int off_task, double *out, int E) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    const float *T = state + size_t(e) * sstride + off_task;
    double *o = out + size_t(e) * PBG_TASK_VIEW_DIM;
    o[0] = __hiloint2double(__float_as_int(T[T_POT_HI]), __float_as_int(T[T_POT_LO]));
    o[1] = T[T_TX]; o[2] = T[T_TY]; o[3] = T[T_FLAGTIMEOUT];
    o[4] = (double)__float_as_int(T[T_FRAME]); o[5] = (double)__float_as_int(T[T_ONGROUND]);
    o[6] = (double)__float_as_int(T[T_STEPS]); o[7] = T[T_RETURN];
    o[8] = T[T_INITZ]; o[9] = (double)__float_as_int(T[T_EPISODE]);
    o[10] = (double)__float_as_int(T[T_ATTACKS]); o[11] = (double)__float_as_int(T[T_FLAGCNT]);
}

int pbg_get_task_view(pbg_handle *h, double *out_dev, void *stream) {
    if (!h || !out_dev) return fail(h, PBG_ERR_INVALID, "pbg_get_task_view: NULL buffer");
    CUDA_TRY(h, cudaSetDevice(h->device));
    { int rc = order_after_last(h, (cudaStream_t)stream); if (rc != PBG_OK) return rc; }
    h->last_stream = (cudaStream_t)stream; h->have_last = true;
    task_view_kernel<<<(h->E + 127) / 128, 128, 0, (cudaStream_t)stream>>>(h->state, h->k.sstride, h->k.off_task, out_dev, h->E);
    CUDA_TRY(h, cudaGetLastError());
    h->launches++;
    return PBG_OK;
}

// development only (not in pbg.h): per-phase cycle counters of builds made with -DPBG_PHASE_CLOCKS
extern "C" int pbg_debug_phases(pbg_handle *h, unsigned long long *out32, int reset) {
    if (!h || !out32) return PBG_ERR_INVALID;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    h->k.phases(out32, reset);
    return PBG_OK;
}

int pbg_num_contact_slots(const pbg_handle *h) { return h ? h->k.nslot : PBG_ERR_INVALID; }

int pbg_enable_contact_export(pbg_handle *h, int32_t enabled) {
    if (!h) return PBG_ERR_INVALID;
    CUDA_TRY(h, cudaSetDevice(h->device));
    if (enabled && !h->d_cand && h->k.nslot > 0) {
        const size_t n = size_t(h->E) * h->k.nslot;
        CUDA_TRY(h, cudaMalloc(&h->d_cand, n * sizeof(float)));
        std::vector<float> inf(n, INFINITY);
        CUDA_TRY(h, cudaMemcpy(h->d_cand, inf.data(), n * sizeof(float), cudaMemcpyHostToDevice));
    } else if (!enabled && h->d_cand) {
        CUDA_TRY(h, cudaDeviceSynchronize());
        cudaFree(h->d_cand); h->d_cand = nullptr;
    }
    return PBG_OK;
}

int pbg_get_contact_candidates(pbg_handle *h, float *out_dev, void *stream) {
    if (!h || !out_dev) return fail(h, PBG_ERR_INVALID, "pbg_get_contact_candidates: NULL buffer");
    if (!h->d_cand) return fail(h, PBG_ERR_INVALID, "pbg_get_contact_candidates: call pbg_enable_contact_export(h, 1) first");
    CUDA_TRY(h, cudaSetDevice(h->device));
    { int rc = order_after_last(h, (cudaStream_t)stream); if (rc != PBG_OK) return rc; }
    h->last_stream = (cudaStream_t)stream; h->have_last = true;
    CUDA_TRY(h, cudaMemcpyAsync(out_dev, h->d_cand, size_t(h->E) * h->k.nslot * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return PBG_OK;
}

int pbg_stats(pbg_handle *h, pbg_episode_stats *out, int32_t reset) {
    if (!h || !out) return fail(h, PBG_ERR_INVALID, "pbg_stats: NULL argument");
    CUDA_TRY(h, cudaSetDevice(h->device));
    unsigned long long raw[8];
    // blocking: ordered after the last stream-ordered call of this handle (torch side streams do not synchronise with
    // the legacy default stream a plain cudaMemcpy runs on)
    cudaStream_t s = h->have_last ? h->last_stream : (cudaStream_t)0;
    {
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(s, &cs) != cudaSuccess) { cudaGetLastError(); s = 0; }
        else if (cs != cudaStreamCaptureStatusNone) return fail(h, PBG_ERR_INVALID, "pbg_stats: the handle's stream is being captured into a CUDA graph");
    }
    CUDA_TRY(h, cudaMemcpyAsync(raw, h->stats, sizeof raw, cudaMemcpyDeviceToHost, s));
    if (reset) CUDA_TRY(h, cudaMemsetAsync(h->stats, 0, sizeof raw, s));
    CUDA_TRY(h, cudaStreamSynchronize(s));
    out->episodes = (int64_t)raw[0];
    out->length_sum = (double)raw[1];
    memcpy(&out->return_sum, &raw[2], sizeof(double));
    out->truncated = (int64_t)raw[3];
    out->nonfinite = (int64_t)raw[4];
    out->steps = (int64_t)raw[5];
    out->contact_overflow = (int64_t)raw[6];
    return PBG_OK;
}

}  // extern "C"
