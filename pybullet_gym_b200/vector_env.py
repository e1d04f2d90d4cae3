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
        return self.obs

    def step(self, actions: torch.Tensor):
        a = self._act(actions)
        with torch.cuda.device(self.device):
            rc = self._L.pbg_step(self._h, _ptr(a), _ptr(self.obs), _ptr(self.reward), _ptr(self.done), _ptr(self.terms),
                                  _ptr(self.final_obs), _ptr(self.truncated), self._stream())
        _lib.check(rc, self._h)
        info = {"reward_terms": self.terms, "final_obs": self.final_obs, "truncated": self.truncated}
        return self.obs, self.reward, self.done, info

    def step_fast(self, actions: torch.Tensor):
        """step() without the optional outputs (reward terms, final obs, truncated)."""
        with torch.cuda.device(self.device):
            rc = self._L.pbg_step(self._h, _ptr(actions), _ptr(self.obs), _ptr(self.reward), _ptr(self.done), None, None,
                                  None, self._stream())
        if rc:
            _lib.check(rc, self._h)
        return self.obs, self.reward, self.done

    def step_host(self, actions_host: torch.Tensor, obs_host: torch.Tensor, reward_host: torch.Tensor,
                  done_host: torch.Tensor):
        """Host-buffer step through pbg_step_host: H2D actions, step, D2H obs/reward/done, synchronised."""
        rc = self._L.pbg_step_host(self._h, _ptr(actions_host), _ptr(obs_host), _ptr(reward_host), _ptr(done_host))
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

    def snapshot(self, pinned_host: bool = False) -> torch.Tensor:
        """Opaque uint8 blob with the whole handle state (pbg_snapshot): physics, warm start, task bookkeeping, reset-RNG
        counters, episode statistics.  On the device by default, in pinned host memory on request."""
        n = int(self._L.pbg_snapshot_bytes(self._h))
        buf = (torch.empty(n, dtype=torch.uint8).pin_memory() if pinned_host
               else torch.empty(n, dtype=torch.uint8, device=self.device))
        with torch.cuda.device(self.device):
            _lib.check(self._L.pbg_snapshot(self._h, _ptr(buf), self._stream()), self._h)
            if pinned_host:
                torch.cuda.current_stream(self.device).synchronize()
        return buf

    def restore(self, blob: torch.Tensor, first_reset_done: bool = True):
        """Resume bit-identically from snapshot(); the blob may come from another VectorEnv of the same env id, batch size,
        seed and env_offset.  The observation / reward buffers are not part of it: call observe() or step()."""
        if blob.dtype != torch.uint8 or not blob.is_contiguous() or blob.numel() != int(self._L.pbg_snapshot_bytes(self._h)):
            raise ValueError("not a snapshot of this env (size %d expected)" % int(self._L.pbg_snapshot_bytes(self._h)))
        with torch.cuda.device(self.device):
            _lib.check(self._L.pbg_restore(self._h, _ptr(blob), self._stream()), self._h)
        self._first_reset_done = bool(first_reset_done)    # host-side quirk flag Q1 (floor in robot.parts), not in the blob

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

    def stats(self, reset: bool = False) -> dict:
        st = _lib.PbgEpisodeStats()
        _lib.check(self._L.pbg_stats(self._h, C.byref(st), int(reset)), self._h)
        return {f: getattr(st, f) for f, _ in st._fields_}

    def launch_count(self) -> int:
        return int(self._L.pbg_launch_count(self._h))
