"""Batched entry point: ``VectorEnv`` steps ``num_envs`` independent worlds of one env id on one GPU
and returns torch CUDA tensors.

Not in the reference (one pybullet client per process there); required by BASELINE.json's
north_star.  Everything numeric happens inside libpbg_b200.so (include/pbg.h); torch is used for
device memory and streams only.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from .spec import SPECS, UNBACKED_IDS, EnvSpec


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


class VectorEnv:
    """``reset() -> obs[E, D]``; ``step(a[E, nA]) -> (obs[E, D], reward[E], done[E], info)``.

    Semantics per env follow the reference's ``WalkerBaseBulletEnv`` / ``InvertedPendulumBulletEnv``
    (roboschool/gym_locomotion_envs.py:22-114, gym_pendulum_envs.py:16-39) wrapped in gym's
    ``TimeLimit(max_episode_steps)``.  With ``auto_reset`` (default) an env that finishes is reset
    inside the step kernel: ``obs`` then is the first observation of the next episode,
    ``info["final_obs"]`` the last one of the finished episode and ``info["truncated"]`` marks
    time-limit endings.
    """

    def __init__(self, env_id: str, num_envs: int, device: Optional[torch.device | int | str] = None, seed: int = 0,
                 env_offset: int = 0, auto_reset: bool = True, rules=None, spec: Optional[EnvSpec] = None):
        if env_id in UNBACKED_IDS:
            raise NotImplementedError("%s is registered by the reference but not implemented on this backend" % env_id)
        if env_id not in SPECS:
            raise KeyError("unknown env id %r" % env_id)
        if not torch.cuda.is_available():
            raise _lib.BackendUnavailable("VectorEnv needs a CUDA device (no CPU fallback)")
        self.spec: EnvSpec = spec if spec is not None else SPECS[env_id]
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.type != "cuda":
            raise _lib.BackendUnavailable("VectorEnv needs a CUDA device (no CPU fallback)")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.num_envs = int(num_envs)
        self.tables = _lib.ModelTables(self.spec, rules)
        self._L = _lib.lib()
        h = C.c_void_p()
        rc = self._L.pbg_create(C.byref(self.tables.c), self.num_envs, self.device.index, seed, env_offset, C.byref(h))
        _lib.check(rc, None)
        self._h = h
        self.obs_dim = self._L.pbg_obs_dim(h)
        self.action_dim = self._L.pbg_action_dim(h)
        self.state_dim = self._L.pbg_state_dim(h)
        self.noise_dim = self._L.pbg_noise_dim(h)
        self.auto_reset = auto_reset
        self._L.pbg_set_auto_reset(h, int(auto_reset))
        E, dev = self.num_envs, self.device
        self.obs = torch.zeros(E, self.obs_dim, device=dev)
        self.reward = torch.zeros(E, device=dev)
        self.done = torch.zeros(E, dtype=torch.uint8, device=dev)
        self.terms = torch.zeros(E, 5, device=dev)
        self.final_obs = torch.zeros(E, self.obs_dim, device=dev)
        self.truncated = torch.zeros(E, dtype=torch.uint8, device=dev)
        self._first_reset_done = False
        self._ready = False          # a reset / set_state / restore has initialised the state
        self._host_ok = {}           # step_host buffers that have passed the argument checks: id -> (address, numel, role, c_void_p)
        self._cand = None

    # ------------------------------------------------------------------ lifecycle
    def close(self):
        if getattr(self, "_h", None):
            self._L.pbg_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _act(self, actions: torch.Tensor) -> torch.Tensor:
        a = actions
        if a.device != self.device or a.dtype != torch.float32 or not a.is_contiguous():
            a = a.to(device=self.device, dtype=torch.float32).contiguous()
        if a.shape != (self.num_envs, self.action_dim):
            raise ValueError("actions must have shape %s" % ((self.num_envs, self.action_dim),))
        return a

    # ------------------------------------------------------------------ gym-like API
    def reset(self, mask: Optional[torch.Tensor] = None, joint_noise: Optional[torch.Tensor] = None,
              floor_in_parts: Optional[bool] = None) -> torch.Tensor:
        """Reset all envs (or those with mask != 0) and return the observations.

        The very first reset of the handle reproduces the reference's first ``env.reset()`` (floor
        not yet in ``robot.parts``, quirk Q1); later resets include the floor.
        """
        if floor_in_parts is None:
            floor_in_parts = self._first_reset_done
        with torch.cuda.device(self.device):
            if joint_noise is not None:
                nz = joint_noise.to(device=self.device, dtype=torch.float32).contiguous()
                if nz.shape != (self.num_envs, self.noise_dim):
                    raise ValueError("joint_noise must have shape %s" % ((self.num_envs, self.noise_dim),))
                rc = self._L.pbg_reset_with(self._h, _ptr(nz), int(floor_in_parts), _ptr(self.obs), self._stream())
            else:
                m = None if mask is None else mask.to(device=self.device, dtype=torch.uint8).contiguous()
                rc = self._L.pbg_reset(self._h, _ptr(m), int(floor_in_parts), _ptr(self.obs), self._stream())
        _lib.check(rc, self._h)
        self._first_reset_done = True
        self._ready = True
        return self.obs

    def seed(self, seed: int):
        """Re-key the device counter RNG (reset noise, Flagrun flag positions, cube attacks); effective from the next reset."""
        _lib.check(self._L.pbg_set_seed(self._h, int(seed) & 0xFFFFFFFFFFFFFFFF), self._h)

    def _check_io(self, t: torch.Tensor, shape, dtype, name: str, cuda: bool):
        """Cheap guards for the raw-pointer fast paths: a wrong device / dtype / stride / size would be silent garbage or an
        out-of-bounds access that poisons the CUDA context."""
        ok_dev = (t.device == self.device) if cuda else (t.device.type == "cpu")
        if not (ok_dev and t.dtype == dtype and t.is_contiguous() and tuple(t.shape) == tuple(shape)):
            raise ValueError("%s must be a contiguous %s tensor of shape %s on %s (got %s %s on %s, contiguous=%s)" % (
                name, dtype, tuple(shape), self.device if cuda else "the host", t.dtype, tuple(t.shape), t.device, t.is_contiguous()))
        if not self._ready:
            raise RuntimeError("step before reset(): the env state is uninitialised")

    def step(self, actions: torch.Tensor):
        a = self._act(actions)
        with torch.cuda.device(self.device):
            rc = self._L.pbg_step(self._h, _ptr(a), _ptr(self.obs), _ptr(self.reward), _ptr(self.done), _ptr(self.terms),
                                  _ptr(self.final_obs), _ptr(self.truncated), self._stream())
        _lib.check(rc, self._h)
        info = {"reward_terms": self.terms, "final_obs": self.final_obs, "truncated": self.truncated}
        return self.obs, self.reward, self.done, info

    def step_fast(self, actions: torch.Tensor):
        """step() without the optional outputs (reward terms, final obs, truncated); `actions` must already be a contiguous
        float32 [num_envs, action_dim] tensor on this env's device."""
        self._check_io(actions, (self.num_envs, self.action_dim), torch.float32, "actions", True)
        with torch.cuda.device(self.device):
            rc = self._L.pbg_step(self._h, _ptr(actions), _ptr(self.obs), _ptr(self.reward), _ptr(self.done), None, None,
                                  None, self._stream())
        if rc:
            _lib.check(rc, self._h)
        return self.obs, self.reward, self.done

    def step_host(self, actions_host: torch.Tensor, obs_host: torch.Tensor, reward_host: torch.Tensor,
                  done_host: torch.Tensor):
        """Host-buffer step through pbg_step_host: H2D actions, step, D2H obs/reward/done, synchronised.  Ordered after
        whatever this env enqueued before (reset() on torch's stream ...) by the library itself."""
        # The argument checks cost more host time than the call itself (~8 us of a ~120 us step): a buffer that has passed them is
        # remembered by object identity + storage address + element count, and only a new one is checked in full.
        E = self.num_envs
        cache = self._host_ok
        ptrs = []
        for t, shape, dtype, name in ((actions_host, (E, self.action_dim), torch.float32, "actions_host"),
                                      (obs_host, (E, self.obs_dim), torch.float32, "obs_host"),
                                      (reward_host, (E,), torch.float32, "reward_host"),
                                      (done_host, (E,), torch.uint8, "done_host")):
            if t is None:
                if name == "actions_host":
                    raise ValueError("actions_host is required")
                ptrs.append(None)
                continue
            p = t.data_ptr()
            ent = cache.get(id(t))
            if ent is None or ent[0] != p or ent[1] != t.numel() or ent[2] != name:
                self._check_io(t, shape, dtype, name, False)
                if len(cache) > 64:
                    cache.clear()
                ent = cache[id(t)] = (p, t.numel(), name, C.c_void_p(p))
            ptrs.append(ent[3])
        if not self._ready:
            raise RuntimeError("step before reset(): the env state is uninitialised")
        rc = self._L.pbg_step_host(self._h, ptrs[0], ptrs[1], ptrs[2], ptrs[3])
        if rc:
            _lib.check(rc, self._h)

    # ------------------------------------------------------------------ fused policy rollouts
    def set_policy(self, w1, b1, w2, b2, w3, b3, obs_shift=None):
        """Two-hidden-layer ReLU policy, weights [in, out] (numpy / torch, any device); evaluated inside the step kernel.
        obs_shift: optional constant added to the observation first (the reference's Humanoid demos do `ob[0] += -0.6`);
        it is folded into the first bias."""
        import numpy as np
        arrs = [np.ascontiguousarray(a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else a, dtype=np.float32)
                for a in (w1, b1, w2, b2, w3, b3)]
        if obs_shift is not None:
            sh = np.asarray(obs_shift.detach().cpu().numpy() if isinstance(obs_shift, torch.Tensor) else obs_shift, dtype=np.float64)
            arrs[1] = (arrs[1].astype(np.float64) + sh @ arrs[0].astype(np.float64)).astype(np.float32)
        h1, h2 = arrs[0].shape[1], arrs[2].shape[1]
        assert arrs[0].shape == (self.obs_dim, h1) and arrs[2].shape == (h1, h2) and arrs[4].shape == (h2, self.action_dim)
        assert arrs[1].shape == (h1,) and arrs[3].shape == (h2,) and arrs[5].shape == (self.action_dim,)
        with torch.cuda.device(self.device):
            rc = self._L.pbg_set_policy(self._h, h1, h2, *[C.c_void_p(a.ctypes.data) for a in arrs])
        _lib.check(rc, self._h)

    def rollout_policy(self, steps: int):
        """`steps` env steps in one launch with the policy set by set_policy(), starting from self.obs (the observation
        reset() / step() returned last).  Returns (obs, reward_sum[E], done_any[E])."""
        with torch.cuda.device(self.device):
            rc = self._L.pbg_rollout_policy(self._h, int(steps), _ptr(self.obs), _ptr(self.reward), _ptr(self.done), self._stream())
        _lib.check(rc, self._h)
        return self.obs, self.reward, self.done

    def set_policy_tensor_cores(self, enabled: bool):
        """Fused policy through mma.sync (TF32 inputs, FP32 accumulation) instead of scalar FP32: faster, actions differ by
        ~1e-3 (pbg_set_policy_tensor_cores)."""
        _lib.check(self._L.pbg_set_policy_tensor_cores(self._h, int(enabled)), self._h)

    def set_zero_copy(self, enabled: bool):
        """pbg_step_host transport: kernel reads / writes pinned host buffers directly (default) or staged copies."""
        _lib.check(self._L.pbg_set_zero_copy(self._h, int(enabled)), self._h)

    def last_host_path(self) -> str:
        return {0: "none", 1: "zero-copy", 2: "staged"}[self._L.pbg_last_host_path(self._h)]

    # ------------------------------------------------------------------ parity / diagnostics
    def get_state(self) -> torch.Tensor:
        s = torch.zeros(self.num_envs, self.state_dim, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self._L.pbg_get_state(self._h, _ptr(s), self._stream()), self._h)
        return s

    def set_state(self, state: torch.Tensor):
        s = state.to(device=self.device, dtype=torch.float32).contiguous()
        assert s.shape == (self.num_envs, self.state_dim)
        with torch.cuda.device(self.device):
            _lib.check(self._L.pbg_set_state(self._h, _ptr(s), self._stream()), self._h)
        self._ready = True

    def _io_tail(self):
        return [self.obs, self.reward, self.done, self.terms, self.final_obs, self.truncated]

    def snapshot(self, pinned_host: bool = False) -> torch.Tensor:
        """Opaque uint8 blob: the whole handle state (pbg_snapshot: physics, warm start, task bookkeeping, reset-RNG
        counters, episode statistics) followed by this object's output buffers (obs, reward, done, reward terms, final obs,
        truncated), so that restore() also brings back what the last step() returned.  On the device by default, in pinned
        host memory on request."""
        n = int(self._L.pbg_snapshot_bytes(self._h))
        tail = torch.cat([t.reshape(-1).view(torch.uint8) for t in self._io_tail()])
        total = n + tail.numel()
        buf = (torch.empty(total, dtype=torch.uint8).pin_memory() if pinned_host
               else torch.empty(total, dtype=torch.uint8, device=self.device))
        with torch.cuda.device(self.device):
            _lib.check(self._L.pbg_snapshot(self._h, _ptr(buf), self._stream()), self._h)
            buf[n:].copy_(tail, non_blocking=True)
            if pinned_host:
                torch.cuda.current_stream(self.device).synchronize()
        return buf

    def restore(self, blob: torch.Tensor, first_reset_done: bool = True):
        """Resume bit-identically from snapshot(); the blob may come from another VectorEnv of the same env id, batch size,
        seed and env_offset.  self.obs / reward / done ... hold what they held when the snapshot was taken.  (observe() is
        NOT a way to get the observation back: it is the task half of a step and has that step's side effects.)"""
        n = int(self._L.pbg_snapshot_bytes(self._h))
        tail_n = sum(t.numel() * t.element_size() for t in self._io_tail())
        if blob.dtype != torch.uint8 or not blob.is_contiguous() or blob.numel() != n + tail_n:
            raise ValueError("not a snapshot of this env (size %d expected)" % (n + tail_n))
        with torch.cuda.device(self.device):
            _lib.check(self._L.pbg_restore(self._h, _ptr(blob), self._stream()), self._h)
            off = n
            for t in self._io_tail():
                nb = t.numel() * t.element_size()
                t.reshape(-1).view(torch.uint8).copy_(blob[off:off + nb], non_blocking=True)
                off += nb
        self._first_reset_done = bool(first_reset_done)    # host-side quirk flag Q1 (floor in robot.parts), not in the blob
        self._ready = True

    def physics_step(self, actions: torch.Tensor, want_contacts: bool = False):
        a = self._act(actions)
        with torch.cuda.device(self.device):
            if want_contacts:
                n = torch.zeros(self.num_envs, dtype=torch.int32, device=self.device)
                _lib.check(self._L.pbg_physics_step_counts(self._h, _ptr(a), _ptr(n), self._stream()), self._h)
                return n
            _lib.check(self._L.pbg_physics_step(self._h, _ptr(a), self._stream()), self._h)

    def observe(self, actions: torch.Tensor):
        a = self._act(actions)
        obs = torch.zeros(self.num_envs, self.obs_dim, device=self.device)
        rew = torch.zeros(self.num_envs, device=self.device)
        done = torch.zeros(self.num_envs, dtype=torch.uint8, device=self.device)
        terms = torch.zeros(self.num_envs, 5, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self._L.pbg_observe(self._h, _ptr(a), _ptr(obs), _ptr(rew), _ptr(done), _ptr(terms), self._stream()),
                       self._h)
        return obs, rew, done, terms

    def feet_contact(self) -> torch.Tensor:
        n = len(self.spec.foot_list)
        out = torch.zeros(self.num_envs, max(n, 1), device=self.device)
        if n:
            with torch.cuda.device(self.device):
                _lib.check(self._L.pbg_get_feet_contact(self._h, _ptr(out), self._stream()), self._h)
        return out[:, :n]

    def task_view(self) -> dict:
        """Task bookkeeping per env (pbg_get_task_view) as float64 tensors [E]: potential, walk_target_x/y, flag_timeout,
        frame, on_ground_frame_counter, episode_steps, episode_return, initial_z, episode, attacks, flag_moves."""
        out = torch.zeros(self.num_envs, _lib.TASK_VIEW_DIM, dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self._L.pbg_get_task_view(self._h, _ptr(out), self._stream()), self._h)
        return {f: out[:, i] for i, f in enumerate(_lib.TASK_VIEW_FIELDS)}

    def enable_contact_export(self, enabled: bool = True):
        _lib.check(self._L.pbg_enable_contact_export(self._h, int(enabled)), self._h)
        self._cand = (torch.zeros(self.num_envs, max(1, self._L.pbg_num_contact_slots(self._h)), device=self.device)
                      if enabled else None)

    def contact_candidates(self) -> torch.Tensor:
        """float[E, slots]: distance of every contact candidate in the last step's final collision pass, +inf where it is
        not in contact; slot meaning: self.tables.contact_slots().  Needs enable_contact_export() before the step."""
        if self._cand is None:
            raise RuntimeError("call enable_contact_export() first")
        with torch.cuda.device(self.device):
            _lib.check(self._L.pbg_get_contact_candidates(self._h, _ptr(self._cand), self._stream()), self._h)
        return self._cand

    def stats(self, reset: bool = False) -> dict:
        st = _lib.PbgEpisodeStats()
        _lib.check(self._L.pbg_stats(self._h, C.byref(st), int(reset)), self._h)
        return {f: getattr(st, f) for f, _ in st._fields_}

    def launch_count(self) -> int:
        return int(self._L.pbg_launch_count(self._h))
