"""ctypes binding of the CPU oracle (oracle/oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this
module.  The physics part of the oracle is **parity unpinned** (see oracle/oracle.h); the task layer
is pinned by tests/golden/task_*.json, generated from the reference's own Python.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional

import numpy as np

from pybullet_gym_b200.mjcf import compiler as mj
from pybullet_gym_b200.spec import SPECS, EnvSpec

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liborc.so")

_pd = C.POINTER(C.c_double)
_pi = C.POINTER(C.c_int32)


class OrcModel(C.Structure):
    _fields_ = [
        ("nl", C.c_int32), ("floating", C.c_int32),
        ("parent", _pi), ("jtype", _pi),
        ("axis", _pd), ("pos", _pd), ("quat", _pd), ("com", _pd),
        ("mass", _pd), ("inertia", _pd),
        ("lower", _pd), ("upper", _pd), ("damping", _pd),
        ("in_parts", _pi),
        ("ng", C.c_int32),
        ("g_link", _pi), ("g_type", _pi), ("g_ground", _pi),
        ("g_radius", _pd), ("g_p0", _pd), ("g_p1", _pd), ("g_friction", _pd), ("g_threshold", _pd),
        ("npair", C.c_int32),
        ("pair_a", _pi), ("pair_b", _pi),
        ("gravity", C.c_double), ("dt_sub", C.c_double),
        ("nsub", C.c_int32), ("niter", C.c_int32),
        ("erp_contact", C.c_double), ("erp_limit", C.c_double), ("linear_slop", C.c_double),
        ("warmstart", C.c_double), ("link_damping", C.c_double), ("max_coord_vel", C.c_double),
        ("ground_friction", C.c_double), ("limit_max_impulse", C.c_double), ("split_impulse_threshold", C.c_double),
        ("limit_split_impulse", C.c_int32), ("max_contacts", C.c_int32),
        ("kind", C.c_int32), ("nact", C.c_int32), ("nfeet", C.c_int32), ("torso_link", C.c_int32),
        ("obs_dim", C.c_int32),
        ("act_link", _pi), ("act_torque", _pd), ("foot_link", _pi),
        ("initial_z", C.c_double),
        ("elec_cost", C.c_double), ("stall_cost", C.c_double), ("limit_cost", C.c_double), ("dt_scene", C.c_double),
        ("walk_target_x", C.c_double), ("walk_target_y", C.c_double),
        ("max_episode_steps", C.c_int32),
        ("stadium_halflen", C.c_double), ("stadium_halfwidth", C.c_double),
        ("torsional", C.c_int32), ("g_spin", _pd), ("g_roll", _pd), ("ground_spin", C.c_double), ("ground_roll", C.c_double),
        ("cube", C.c_int32),
        ("cube_half", C.c_double), ("cube_mass", C.c_double), ("cube_inertia", C.c_double),
        ("cube_friction", C.c_double), ("cube_threshold", C.c_double), ("cube_pos0", C.c_double * 3),
        ("friction_cone", C.c_int32),
        ("aux_link", C.c_int32 * 2),
        ("ground_manifold", C.c_int32),
    ]


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "liborc.so"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.POINTER(OrcModel), C.c_uint64, C.c_uint64]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_reset.argtypes = [C.c_void_p, C.c_int, _pd]
        L.orc_reset_with.argtypes = [C.c_void_p, _pd, C.c_int, _pd]
        L.orc_step.argtypes = [C.c_void_p, _pd, _pd, _pd, _pd]
        L.orc_step.restype = C.c_int
        L.orc_observe.argtypes = [C.c_void_p, _pd, _pd, _pd, _pd]
        L.orc_observe.restype = C.c_int
        L.orc_physics_step.argtypes = [C.c_void_p, _pd]
        L.orc_get_state.argtypes = [C.c_void_p, _pd]
        L.orc_set_state.argtypes = [C.c_void_p, _pd]
        L.orc_state_size.argtypes = [C.POINTER(OrcModel)]
        L.orc_num_dofs.argtypes = [C.POINTER(OrcModel)]
        L.orc_num_contacts.argtypes = [C.c_void_p]
        L.orc_feet_contact.argtypes = [C.c_void_p, _pd]
        L.orc_link_com.argtypes = [C.c_void_p, _pd]
        L.orc_energy.argtypes = [C.c_void_p]
        L.orc_energy.restype = C.c_double
        L.orc_mass_matrix_inv.argtypes = [C.c_void_p, _pd]
        L.orc_physics_step_torque.argtypes = [C.c_void_p, _pd]
        L.orc_link_state.argtypes = [C.c_void_p, _pd]
        L.orc_get_contacts.argtypes = [C.c_void_p, _pi, _pi, _pd]
        L.orc_get_contacts.restype = C.c_int
        L.orc_set_joint.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double]
        L.orc_get_joint.argtypes = [C.c_void_p, C.c_int, _pd, _pd]
        L.orc_set_cube.argtypes = [C.c_void_p, _pd, _pd, _pd, _pd]
        L.orc_get_cube.argtypes = [C.c_void_p, _pd, _pd, _pd, _pd]
        L.orc_get_rows.argtypes = [C.c_void_p, _pd]
        L.orc_get_rows.restype = C.c_int
        L.orc_get_debug.argtypes = [C.c_void_p, _pd]
        L.orc_set_tape.argtypes = [C.c_void_p, _pd, C.c_int]
        L.orc_rollout.argtypes = [C.c_void_p, C.c_long, C.c_uint64, _pd, C.POINTER(C.c_long)]
        L.orc_rollout.restype = C.c_long
        L.orc_episodes.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_uint64, _pd, _pi]
        L.orc_episodes.restype = C.c_long
        L.orc_cap_overflows.argtypes = [C.c_void_p]
        L.orc_set_max_rows.argtypes = [C.c_void_p, C.c_int]
        L.orc_feet_margin.argtypes = [C.c_void_p, _pd]
        L.orc_done_margin.argtypes = [C.c_void_p]
        L.orc_done_margin.restype = C.c_double
        _lib = L
    return _lib


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def random_policy_episodes(env_id_or_spec, n_episodes: int, cap: int, seed: int = 0, threads: Optional[int] = None, **kw):
    """n_episodes whole random-policy episodes of the oracle on `threads` host threads (ctypes releases the GIL inside
    orc_episodes): (returns, lengths).  Thread k owns RNG stream env_index = k."""
    import threading
    threads = threads or min(os.cpu_count() or 1, 32)
    per = [n_episodes // threads + (1 if k < n_episodes % threads else 0) for k in range(threads)]
    envs = [OracleEnv(env_id_or_spec, seed=seed, env_index=k, **kw) for k in range(threads)]
    out = [None] * threads

    def work(k):
        out[k] = envs[k].episodes(per[k], cap, action_seed=seed + 1) if per[k] else (np.zeros(0), np.zeros(0, np.int32))

    th = [threading.Thread(target=work, args=(k,)) for k in range(threads)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    return np.concatenate([o[0] for o in out]), np.concatenate([o[1] for o in out])


def _i(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class OracleModel:
    """Flat copy of a BulletModel + EnvSpec in the layout oracle.c expects."""

    def __init__(self, spec: EnvSpec, bm: Optional[mj.BulletModel] = None, max_contacts: int = 0, **overrides):
        self.spec = spec
        self.bm = bm = bm or mj.parse_mjcf(spec.xml)
        L = bm.links
        nl = len(L)
        part_names = set(bm.part_names())
        keep = self._keep = {}
        keep["parent"] = _i([l.parent for l in L])
        keep["jtype"] = _i([l.jtype for l in L])
        keep["axis"] = _d([l.axis for l in L])
        keep["pos"] = _d([l.pos for l in L])
        keep["quat"] = _d([l.quat for l in L])
        keep["com"] = _d([l.com for l in L])
        keep["mass"] = _d([l.mass for l in L])
        keep["inertia"] = _d([l.inertia for l in L])
        keep["lower"] = _d([l.lower for l in L])
        keep["upper"] = _d([l.upper for l in L])
        keep["damping"] = _d([l.damping for l in L])
        keep["in_parts"] = _i([1 if (l.name in part_names and (i > 0 or bm.floating)) else 0 for i, l in enumerate(L)])
        g_link, g_type, g_ground, g_rad, g_p0, g_p1, g_fr, g_thr = [], [], [], [], [], [], [], []
        g_spin, g_roll = [], []
        filt = []
        for i, l in enumerate(L):
            for g in l.geoms:
                g_link.append(i); g_type.append(g.gtype); g_rad.append(g.radius)
                g_p0.append(g.p0); g_p1.append(g.p1); g_fr.append(g.friction); g_thr.append(l.contact_threshold)
                g_ground.append(1 if ((g.contype & ~2) or (2 & g.conaffinity)) else 0)
                g_spin.append(g.spin_friction); g_roll.append(g.roll_friction)
                filt.append((g.contype, g.conaffinity))
        ng = len(g_link)
        pa, pb = [], []
        for a in range(ng):
            for b in range(a + 1, ng):
                la, lb = g_link[a], g_link[b]
                if la == lb:
                    continue
                if not ((filt[a][0] & filt[b][1]) or (filt[b][0] & filt[a][1])):
                    continue
                if la in bm.ancestors(lb) or lb in bm.ancestors(la):
                    continue
                pa.append(a); pb.append(b)
        keep["g_link"], keep["g_type"], keep["g_ground"] = _i(g_link), _i(g_type), _i(g_ground)
        keep["g_radius"], keep["g_p0"], keep["g_p1"] = _d(g_rad), _d(g_p0).reshape(-1), _d(g_p1).reshape(-1)
        keep["g_friction"], keep["g_threshold"] = _d(g_fr), _d(g_thr)
        keep["g_spin"], keep["g_roll"] = _d(g_spin), _d(g_roll)
        keep["pair_a"], keep["pair_b"] = _i(pa), _i(pb)
        oj = bm.ordered_joints()
        names = [L[i].joint_name for i in oj]
        scale = spec.torque_scale(names)
        nact = spec.action_dim
        keep["act_link"] = _i(oj[:nact])
        keep["act_torque"] = _d(scale[:nact])
        keep["foot_link"] = _i([bm.link_index(f) for f in spec.foot_list])
        sc = spec.scene
        m = OrcModel()
        m.nl, m.floating = nl, int(bm.floating)
        m.ng, m.npair = ng, len(pa)
        for k, v in keep.items():
            setattr(m, k, v.ctypes.data_as(_pd if v.dtype == np.float64 else _pi))
        m.gravity, m.dt_sub, m.nsub, m.niter = sc.gravity, sc.timestep, sc.frame_skip, sc.num_solver_iterations
        m.erp_contact, m.erp_limit, m.linear_slop = sc.contact_erp, sc.erp, sc.linear_slop
        m.warmstart, m.link_damping, m.max_coord_vel = sc.warmstarting_factor, bm.rules.link_damping, sc.max_coordinate_velocity
        m.ground_friction, m.limit_max_impulse = sc.ground_friction, sc.limit_max_impulse
        m.split_impulse_threshold, m.limit_split_impulse = sc.split_impulse_threshold, int(sc.limit_split_impulse)
        m.max_contacts = max_contacts
        m.kind, m.nact, m.nfeet, m.obs_dim = spec.kind, nact, len(spec.foot_list), spec.obs_dim
        m.torso_link = bm.link_index(spec.robot_name)
        m.initial_z = -1.0 if spec.initial_z is None else spec.initial_z
        m.elec_cost, m.stall_cost, m.limit_cost = spec.electricity_cost, spec.stall_torque_cost, spec.joints_at_limit_cost
        m.dt_scene = sc.dt
        m.walk_target_x, m.walk_target_y = spec.walk_target
        m.max_episode_steps = spec.max_episode_steps
        m.stadium_halflen, m.stadium_halfwidth = sc.stadium_halflen, sc.stadium_halfwidth
        m.torsional = int(getattr(sc, 'torsional_friction', False))
        m.ground_spin, m.ground_roll = sc.ground_spinning_friction, sc.ground_rolling_friction
        for i in range(2):
            m.aux_link[i] = bm.link_index(spec.aux_links[i]) if i < len(spec.aux_links) else -1
        cube = spec.cube
        m.cube = 1 if cube is not None else 0
        if cube is not None:
            m.cube_half, m.cube_mass, m.cube_inertia = cube.half_extent, cube.mass, cube.inertia
            m.cube_friction = cube.friction
            m.cube_threshold = (cube.contact_threshold if bm.rules.relative_breaking_threshold else cube.breaking_threshold)
            for i in range(3):
                m.cube_pos0[i] = cube.pos0[i]
        for k, v in overrides.items():
            setattr(m, k, v)
        self.c = m
        self.nd = lib().orc_num_dofs(C.byref(m))
        self.state_size = lib().orc_state_size(C.byref(m))
        self.nu = self.nd + (6 if bm.floating else 0)


class OracleEnv:
    def __init__(self, env_id_or_spec, seed: int = 0, env_index: int = 0, max_contacts: int = 0, bm=None, max_rows: int = 0, **overrides):
        spec = SPECS[env_id_or_spec] if isinstance(env_id_or_spec, str) else env_id_or_spec
        self.model = OracleModel(spec, bm=bm, max_contacts=max_contacts, **overrides)
        self.spec = spec
        self._h = lib().orc_create(C.byref(self.model.c), seed, env_index)
        if max_rows:
            lib().orc_set_max_rows(self._h, int(max_rows))
        self.obs_dim, self.nact = spec.obs_dim, spec.action_dim

    def __del__(self):
        if getattr(self, "_h", None) and lib is not None:      # module globals may be gone at interpreter shutdown
            try:
                lib().orc_destroy(self._h)
            except Exception:
                pass
            self._h = None

    def reset(self, noise=None, floor_in_parts: bool = True):
        obs = np.zeros(self.obs_dim)
        if noise is None:
            lib().orc_reset(self._h, int(floor_in_parts), obs.ctypes.data_as(_pd))
        else:
            nz = _d(noise)
            lib().orc_reset_with(self._h, nz.ctypes.data_as(_pd), int(floor_in_parts), obs.ctypes.data_as(_pd))
        return obs

    def step(self, action):
        a = _d(action)
        obs, rew, terms = np.zeros(self.obs_dim), C.c_double(0), np.zeros(5)
        done = lib().orc_step(self._h, a.ctypes.data_as(_pd), obs.ctypes.data_as(_pd), C.byref(rew), terms.ctypes.data_as(_pd))
        return obs, rew.value, bool(done), terms

    def observe(self, action):
        a = _d(action)
        obs, rew, terms = np.zeros(self.obs_dim), C.c_double(0), np.zeros(5)
        done = lib().orc_observe(self._h, a.ctypes.data_as(_pd), obs.ctypes.data_as(_pd), C.byref(rew), terms.ctypes.data_as(_pd))
        return obs, rew.value, bool(done), terms

    def physics_step(self, action):
        a = _d(action)
        lib().orc_physics_step(self._h, a.ctypes.data_as(_pd))

    def get_state(self):
        s = np.zeros(self.model.state_size)
        lib().orc_get_state(self._h, s.ctypes.data_as(_pd))
        return s

    def set_state(self, s):
        s = _d(s)
        assert s.size == self.model.state_size
        lib().orc_set_state(self._h, s.ctypes.data_as(_pd))

    def physics_step_torque(self, tau):
        t = _d(tau)
        assert t.size == self.model.nd
        lib().orc_physics_step_torque(self._h, t.ctypes.data_as(_pd))

    def link_state(self):
        out = np.zeros(10 * self.model.c.nl)
        lib().orc_link_state(self._h, out.ctypes.data_as(_pd))
        return out.reshape(-1, 10)

    def contacts(self):
        la, lb, d = np.zeros(256, np.int32), np.zeros(256, np.int32), np.zeros(256)
        n = lib().orc_get_contacts(self._h, la.ctypes.data_as(_pi), lb.ctypes.data_as(_pi), d.ctypes.data_as(_pd))
        return la[:n].copy(), lb[:n].copy(), d[:n].copy()

    def set_joint(self, dof, q, qd):
        lib().orc_set_joint(self._h, int(dof), float(q), float(qd))

    def get_joint(self, dof):
        q, qd = C.c_double(0), C.c_double(0)
        lib().orc_get_joint(self._h, int(dof), C.byref(q), C.byref(qd))
        return q.value, qd.value

    def set_cube(self, pos=None, quat=None, omega=None, vel=None):
        arrs = [None if a is None else _d(a) for a in (pos, quat, omega, vel)]
        lib().orc_set_cube(self._h, *[None if a is None else a.ctypes.data_as(_pd) for a in arrs])

    def get_cube(self):
        p, q, w, v = np.zeros(3), np.zeros(4), np.zeros(3), np.zeros(3)
        lib().orc_get_cube(self._h, *[a.ctypes.data_as(_pd) for a in (p, q, w, v)])
        return p, q, w, v

    def rows(self):
        out = np.zeros(2 + 3 * 256)
        n = lib().orc_get_rows(self._h, out.ctypes.data_as(_pd))
        return int(out[0]), int(out[1]), out[2:2 + 3 * n].reshape(n, 3)

    def debug(self):
        out = np.zeros(192)
        lib().orc_get_debug(self._h, out.ctypes.data_as(_pd))
        n = self.model.nu
        return out[:n], out[64:64 + n], out[128:128 + n]

    def set_tape(self, values):
        t = _d(values)
        lib().orc_set_tape(self._h, t.ctypes.data_as(_pd), int(t.size))

    def num_contacts(self):
        return lib().orc_num_contacts(self._h)

    def feet_contact(self):
        out = np.zeros(max(1, len(self.spec.foot_list)))
        lib().orc_feet_contact(self._h, out.ctypes.data_as(_pd))
        return out[:len(self.spec.foot_list)]

    def link_com(self):
        out = np.zeros(3 * self.model.c.nl)
        lib().orc_link_com(self._h, out.ctypes.data_as(_pd))
        return out.reshape(-1, 3)

    def energy(self):
        return lib().orc_energy(self._h)

    def mass_matrix_inv(self):
        n = self.model.nu
        out = np.zeros(n * n)
        lib().orc_mass_matrix_inv(self._h, out.ctypes.data_as(_pd))
        return out.reshape(n, n)

    def episodes(self, n: int, cap: int, action_seed: int = 0):
        """n whole random-policy episodes (each at most `cap` steps): (returns[n], lengths[n])."""
        ret, ln = np.zeros(n), np.zeros(n, np.int32)
        lib().orc_episodes(self._h, n, cap, action_seed, ret.ctypes.data_as(_pd), ln.ctypes.data_as(_pi))
        return ret, ln

    def feet_margin(self):
        out = np.zeros(max(1, len(self.spec.foot_list)))
        lib().orc_feet_margin(self._h, out.ctypes.data_as(_pd))
        return out[:len(self.spec.foot_list)]

    def done_margin(self) -> float:
        return float(lib().orc_done_margin(self._h))

    def cap_overflows(self) -> int:
        return lib().orc_cap_overflows(self._h)

    def rollout(self, steps: int, action_seed: int = 0):
        rs, ep = C.c_double(0), C.c_long(0)
        n = lib().orc_rollout(self._h, steps, action_seed, C.byref(rs), C.byref(ep))
        return n, rs.value, ep.value
