"""ctypes binding of libpbg_b200.so (include/pbg.h).

The shared library is built in-tree by ``__graft_entry__.build()`` / ``build_extension()`` with
``nvcc -gencode arch=compute_100a,code=sm_100a``.  There is no CPU fallback: if the library is
missing or cannot be loaded, importing anything that steps physics raises ``BackendUnavailable``.
"""
from __future__ import annotations

import ctypes as C
import os
import sys
import subprocess
from typing import Optional

import numpy as np

from .mjcf import compiler as mj
from .spec import EnvSpec

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PBG_LIB") or os.path.join(_PKG, "libpbg_b200.so")
CSRC = os.path.join(_PKG, "csrc")
INCLUDE = os.path.join(_PKG, "..", "include")


class BackendUnavailable(RuntimeError):
    pass


class PbgError(RuntimeError):
    pass


_pd = C.POINTER(C.c_double)
_pi = C.POINTER(C.c_int32)
_pf = C.POINTER(C.c_float)
_pu8 = C.POINTER(C.c_uint8)


class PbgModel(C.Structure):
    _fields_ = [
        ("nb", C.c_int32), ("nj", C.c_int32), ("floating", C.c_int32),
        ("parent", _pi), ("jtype", _pi),
        ("q0", _pd), ("anchor_p", _pd), ("com_off", _pd), ("axis", _pd), ("mass", _pd), ("inertia", _pd),
        ("jnt_lower", _pd), ("jnt_upper", _pd), ("jnt_damping", _pd), ("jnt_act", _pi), ("jnt_torque", _pd),
        ("ns", C.c_int32),
        ("sub_body", _pi), ("sub_off", _pd), ("sub_mass", _pd), ("sub_inertia", _pd), ("sub_in_parts", _pi),
        ("torso_sub", C.c_int32),
        ("ng", C.c_int32),
        ("geom_body", _pi), ("geom_type", _pi), ("geom_ground", _pi), ("geom_foot", _pi),
        ("geom_radius", _pd), ("geom_p0", _pd), ("geom_p1", _pd), ("geom_friction", _pd), ("geom_threshold", _pd),
        ("npair", C.c_int32),
        ("pair_a", _pi), ("pair_b", _pi),
        ("gravity", C.c_double), ("timestep", C.c_double),
        ("frame_skip", C.c_int32), ("num_solver_iterations", C.c_int32),
        ("contact_erp", C.c_double), ("erp", C.c_double), ("linear_slop", C.c_double),
        ("warmstarting_factor", C.c_double), ("link_damping", C.c_double), ("max_coordinate_velocity", C.c_double),
        ("ground_friction", C.c_double), ("limit_max_impulse", C.c_double), ("split_impulse_threshold", C.c_double),
        ("limit_split_impulse", C.c_int32), ("max_contacts", C.c_int32),
        ("kind", C.c_int32), ("action_dim", C.c_int32), ("obs_dim", C.c_int32), ("nfeet", C.c_int32),
        ("max_episode_steps", C.c_int32),
        ("initial_z", C.c_double),
        ("electricity_cost", C.c_double), ("stall_torque_cost", C.c_double), ("joints_at_limit_cost", C.c_double),
        ("walk_target_x", C.c_double), ("walk_target_y", C.c_double),
        ("stadium_halflen", C.c_double), ("stadium_halfwidth", C.c_double),
        ("cube", C.c_int32),
        ("cube_half", C.c_double), ("cube_mass", C.c_double), ("cube_inertia", C.c_double),
        ("cube_friction", C.c_double), ("cube_threshold", C.c_double), ("cube_pos0", C.c_double * 3),
        ("aux_sub", C.c_int32 * 2),
        ("torsional_friction", C.c_int32), ("geom_spin", _pd), ("geom_roll", _pd),
        ("ground_spinning_friction", C.c_double), ("ground_rolling_friction", C.c_double),
    ]


class PbgEpisodeStats(C.Structure):
    _fields_ = [("return_sum", C.c_double), ("length_sum", C.c_double), ("episodes", C.c_int64),
                ("truncated", C.c_int64), ("nonfinite", C.c_int64), ("steps", C.c_int64), ("contact_overflow", C.c_int64)]


PBG_VERSION = 103          # include/pbg.h
TASK_VIEW_DIM = 12         # PBG_TASK_VIEW_DIM
TASK_VIEW_FIELDS = ("potential", "walk_target_x", "walk_target_y", "flag_timeout", "frame", "on_ground_frame_counter",
                    "episode_steps", "episode_return", "initial_z", "episode", "attacks", "flag_moves")


# -prec-div=false -prec-sqrt=false -ftz=true: single-precision divisions and square roots as reciprocal / rsqrt sequences
# (2 ulp instead of the IEEE sequence with its slow-path call) and denormals flushed: Ant +4.3 %, Humanoid +6.1 % (A/B on one box,
# profiles/r02_experiments.md); the parity suite runs against the oracle with these flags.  sinf / cosf / atan2f keep their
# accurate versions (no -use_fast_math: __sinf of a 1e-3 rad half-angle would cost the quaternion integration 4 digits).
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
              "-prec-div=false", "-prec-sqrt=false", "-ftz=true"]
BUILD_DIR = os.path.join(CSRC, "_build")


def build_extension(force: bool = False, verbose: bool = False, jobs: Optional[int] = None, extra_flags=()) -> str:
    """Compile csrc/*.cu into libpbg_b200.so for sm_100a (cross-compiles without a GPU).  One object per translation unit
    (pbg_abi.cu + one pbg_k_*.cu per kernel configuration group), compiled in parallel, objects cached under csrc/_build."""
    from concurrent.futures import ThreadPoolExecutor
    headers = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(".cuh")] + [os.path.join(INCLUDE, "pbg.h")]
    units = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(".cu")]
    os.makedirs(BUILD_DIR, exist_ok=True)
    hdr_time = max(os.path.getmtime(h) for h in headers)
    flags = NVCC_FLAGS + list(extra_flags) + (["-Xptxas", "-v"] if verbose else [])
    stamp = os.path.join(BUILD_DIR, "flags.txt")
    if (open(stamp).read() if os.path.exists(stamp) else None) != " ".join(flags):
        force = True

    def obj_of(u):
        return os.path.join(BUILD_DIR, os.path.basename(u)[:-3] + ".o")

    stale = [u for u in units if force or not os.path.exists(obj_of(u))
             or os.path.getmtime(obj_of(u)) < max(hdr_time, os.path.getmtime(u))]
    objs = [obj_of(u) for u in units]
    if not stale and os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(o) for o in objs):
        return LIB_PATH

    def compile_one(u):
        r = subprocess.run(["nvcc"] + flags + ["-c", "-o", obj_of(u), u], capture_output=True, text=True)
        return u, r

    # the big units first, so that the longest compile starts at once
    stale.sort(key=lambda u: -os.path.getsize(u) - (1 << 20) * any(k in u for k in ("harder", "humanoid")))
    with ThreadPoolExecutor(max_workers=jobs or min(len(stale) or 1, os.cpu_count() or 1)) as ex:
        for u, r in ex.map(compile_one, stale):
            if verbose and (r.stdout or r.stderr):
                sys.stderr.write(r.stdout + r.stderr)
            if r.returncode != 0:
                raise RuntimeError("nvcc failed on %s:\n%s" % (u, r.stdout + r.stderr))
    subprocess.check_call(["nvcc", "-shared", "-o", LIB_PATH] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"])
    open(stamp, "w").write(" ".join(flags))
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise BackendUnavailable(
            "libpbg_b200.so is not built (run `python -c 'import __graft_entry__ as g; g.build()'`); "
            "this backend has no CPU fallback")
    try:
        L = C.CDLL(LIB_PATH)
    except OSError as e:
        raise BackendUnavailable("cannot load %s: %s (no CPU fallback)" % (LIB_PATH, e)) from e
    vp = C.c_void_p
    L.pbg_version.restype = C.c_int
    L.pbg_max_contacts.argtypes = [C.c_int]
    L.pbg_max_rows.argtypes = [C.c_int]
    L.pbg_create.argtypes = [C.POINTER(PbgModel), C.c_int32, C.c_int32, C.c_uint64, C.c_uint64, C.POINTER(vp)]
    L.pbg_destroy.argtypes = [vp]
    L.pbg_last_error.argtypes = [vp]
    L.pbg_last_error.restype = C.c_char_p
    for f in ("pbg_num_envs", "pbg_obs_dim", "pbg_action_dim", "pbg_state_dim", "pbg_noise_dim"):
        getattr(L, f).argtypes = [vp]
    L.pbg_reset.argtypes = [vp, vp, C.c_int32, vp, vp]
    L.pbg_reset_with.argtypes = [vp, vp, C.c_int32, vp, vp]
    L.pbg_step.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp]
    L.pbg_step_host.argtypes = [vp, vp, vp, vp, vp]
    L.pbg_set_auto_reset.argtypes = [vp, C.c_int32]
    L.pbg_set_policy.argtypes = [vp, C.c_int32, C.c_int32, vp, vp, vp, vp, vp, vp]
    L.pbg_rollout_policy.argtypes = [vp, C.c_int32, vp, vp, vp, vp]
    L.pbg_set_policy_tensor_cores.argtypes = [vp, C.c_int32]
    L.pbg_set_zero_copy.argtypes = [vp, C.c_int32]
    L.pbg_last_host_path.argtypes = [vp]
    L.pbg_get_state.argtypes = [vp, vp, vp]
    L.pbg_set_state.argtypes = [vp, vp, vp]
    L.pbg_physics_step.argtypes = [vp, vp, vp]
    L.pbg_physics_step_counts.argtypes = [vp, vp, vp, vp]
    L.pbg_observe.argtypes = [vp, vp, vp, vp, vp, vp, vp]
    L.pbg_get_feet_contact.argtypes = [vp, vp, vp]
    L.pbg_stats.argtypes = [vp, C.POINTER(PbgEpisodeStats), C.c_int32]
    L.pbg_measure_fp32_peak.argtypes = [C.c_int32, C.POINTER(C.c_double)]
    L.pbg_launch_count.argtypes = [vp]
    L.pbg_launch_count.restype = C.c_int64
    L.pbg_snapshot_bytes.argtypes = [vp]
    L.pbg_snapshot_bytes.restype = C.c_int64
    L.pbg_snapshot.argtypes = [vp, vp, vp]
    L.pbg_restore.argtypes = [vp, vp, vp]
    L.pbg_set_seed.argtypes = [vp, C.c_uint64]
    L.pbg_get_task_view.argtypes = [vp, vp, vp]
    L.pbg_num_contact_slots.argtypes = [vp]
    L.pbg_enable_contact_export.argtypes = [vp, C.c_int32]
    L.pbg_get_contact_candidates.argtypes = [vp, vp, vp]
    if L.pbg_version() != PBG_VERSION:
        raise BackendUnavailable("%s is version %d, this package needs %d: rebuild it (__graft_entry__.build())"
                                 % (LIB_PATH, L.pbg_version(), PBG_VERSION))
    _lib = L
    return L


EXPORTS = ["pbg_version", "pbg_create", "pbg_destroy", "pbg_last_error", "pbg_num_envs", "pbg_obs_dim",
           "pbg_action_dim", "pbg_state_dim", "pbg_noise_dim", "pbg_reset", "pbg_reset_with", "pbg_step", "pbg_step_host",
           "pbg_set_auto_reset", "pbg_set_zero_copy", "pbg_last_host_path", "pbg_set_policy", "pbg_rollout_policy", "pbg_set_policy_tensor_cores", "pbg_get_state", "pbg_set_state", "pbg_physics_step", "pbg_physics_step_counts",
           "pbg_max_contacts", "pbg_max_rows", "pbg_measure_fp32_peak", "pbg_observe", "pbg_get_feet_contact", "pbg_stats", "pbg_launch_count",
           "pbg_snapshot_bytes", "pbg_snapshot", "pbg_restore", "pbg_set_seed", "pbg_get_task_view", "pbg_num_contact_slots",
           "pbg_enable_contact_export", "pbg_get_contact_candidates"]


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class ModelTables:
    """pbg_model built from an EnvSpec: MJCF -> BulletModel -> ReducedModel -> flat C arrays."""

    def __init__(self, spec: EnvSpec, rules: Optional[mj.ImporterRules] = None, max_contacts: Optional[int] = None):
        self.spec = spec
        self.bullet = bm = mj.parse_mjcf(spec.xml, rules)
        ordered = [bm.links[i].joint_name for i in bm.ordered_joints()]
        self.ordered_joint_names = ordered
        act_names = ordered[:spec.action_dim]
        self.reduced = rm = mj.reduce_model(bm, act_names)
        scale = dict(zip(ordered, spec.torque_scale(ordered)))
        k = self._keep = {}
        k["parent"], k["jtype"] = _i(rm.parent), _i(rm.jtype)
        k["q0"], k["anchor_p"], k["com_off"], k["axis"] = _d(rm.q0), _d(rm.anchor_p), _d(rm.com_off), _d(rm.axis)
        k["mass"], k["inertia"] = _d(rm.mass), _d(rm.inertia)
        k["jnt_lower"], k["jnt_upper"], k["jnt_damping"] = _d(rm.jnt_lower), _d(rm.jnt_upper), _d(rm.jnt_damping)
        k["jnt_act"] = _i(rm.jnt_act)
        k["jnt_torque"] = _d([scale.get(n, 0.0) if a >= 0 else 0.0 for n, a in zip(rm.jnt_names, rm.jnt_act)])
        k["sub_body"], k["sub_off"], k["sub_mass"] = _i(rm.sub_body), _d(rm.sub_off), _d(rm.sub_mass)
        k["sub_inertia"], k["sub_in_parts"] = _d(rm.sub_inertia), _i(rm.sub_in_parts)
        foot_of_link = {name: n for n, name in enumerate(spec.foot_list)}
        k["geom_foot"] = _i([foot_of_link.get(rm.sub_names[s], -1) for s in rm.geom_link])
        k["geom_body"], k["geom_type"], k["geom_ground"] = _i(rm.geom_body), _i(rm.geom_type), _i(rm.geom_ground)
        k["geom_radius"], k["geom_p0"], k["geom_p1"] = _d(rm.geom_radius), _d(rm.geom_p0), _d(rm.geom_p1)
        k["geom_friction"], k["geom_threshold"] = _d(rm.geom_friction), _d(rm.geom_threshold)
        k["geom_spin"], k["geom_roll"] = _d(rm.geom_spin), _d(rm.geom_roll)
        k["pair_a"], k["pair_b"] = _i(rm.pair_a), _i(rm.pair_b)
        m = PbgModel()
        m.nb, m.nj, m.floating = rm.nb, rm.nj, int(rm.floating)
        m.ns, m.ng, m.npair = len(rm.sub_body), len(rm.geom_body), len(rm.pair_a)
        m.torso_sub = rm.sub_names.index(spec.robot_name)
        for name, arr in k.items():
            setattr(m, name, arr.ctypes.data_as(_pd if arr.dtype == np.float64 else _pi))
        sc = spec.scene
        # torsional friction rows: only the kernel configurations built with them accept the switch (pbg_create checks)
        m.torsional_friction = int(sc.torsional_friction)
        m.ground_spinning_friction, m.ground_rolling_friction = sc.ground_spinning_friction, sc.ground_rolling_friction
        m.gravity, m.timestep, m.frame_skip, m.num_solver_iterations = sc.gravity, sc.timestep, sc.frame_skip, sc.num_solver_iterations
        m.contact_erp, m.erp, m.linear_slop, m.warmstarting_factor = sc.contact_erp, sc.erp, sc.linear_slop, sc.warmstarting_factor
        m.link_damping, m.max_coordinate_velocity = rm.link_damping, sc.max_coordinate_velocity
        m.ground_friction, m.limit_max_impulse = sc.ground_friction, sc.limit_max_impulse
        m.split_impulse_threshold, m.limit_split_impulse = sc.split_impulse_threshold, int(sc.limit_split_impulse)
        m.max_contacts = lib().pbg_max_contacts(spec.kind) if max_contacts is None else max_contacts
        m.kind, m.action_dim, m.obs_dim, m.nfeet = spec.kind, spec.action_dim, spec.obs_dim, len(spec.foot_list)
        m.max_episode_steps = spec.max_episode_steps
        m.initial_z = -1.0 if spec.initial_z is None else spec.initial_z
        m.electricity_cost, m.stall_torque_cost = spec.electricity_cost, spec.stall_torque_cost
        m.joints_at_limit_cost = spec.joints_at_limit_cost
        m.walk_target_x, m.walk_target_y = spec.walk_target
        m.stadium_halflen, m.stadium_halfwidth = sc.stadium_halflen, sc.stadium_halfwidth
        for i in range(2):
            m.aux_sub[i] = rm.sub_names.index(spec.aux_links[i]) if i < len(spec.aux_links) else -1
        cube = spec.cube
        m.cube = 1 if cube is not None else 0
        if cube is not None:
            m.cube_half, m.cube_mass, m.cube_inertia, m.cube_friction = cube.half_extent, cube.mass, cube.inertia, cube.friction
            m.cube_threshold = cube.contact_threshold if bm.rules.relative_breaking_threshold else cube.breaking_threshold
            for i in range(3):
                m.cube_pos0[i] = cube.pos0[i]
        self.c = m

    def contact_slots(self):
        """(link name, other) per contact-candidate slot, in the order include/pbg.h documents for
        pbg_get_contact_candidates; other is "floor", "cube" or the second link's name."""
        rm, spec = self.reduced, self.spec
        link = [rm.sub_names[s] for s in rm.geom_link]
        out = []
        for g in range(len(rm.geom_body)):
            if rm.geom_ground[g]:
                out += [(link[g], "floor")] * (2 if rm.geom_type[g] == mj.G_CAPSULE else 1)
        if spec.cube is not None:
            out += [("cube", "floor")] * 8
        out += [(link[a], link[b]) for a, b in zip(rm.pair_a, rm.pair_b)]
        if spec.cube is not None:
            out += [(link[g], "cube") for g in range(len(rm.geom_body))]
        return out


def solver_budget(kind: int) -> dict:
    """The kernel's solver budgets for an env kind, as keyword arguments of oracle.OracleEnv (tests compare like with like)."""
    return {"max_contacts": lib().pbg_max_contacts(kind), "max_rows": lib().pbg_max_rows(kind)}


def check(rc: int, handle=None):
    if rc != 0:
        msg = lib().pbg_last_error(handle)
        raise PbgError("libpbg_b200 error %d: %s" % (rc, msg.decode() if msg else "?"))
