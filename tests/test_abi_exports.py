"""The C-ABI library loads on a CPU-only box and exports every symbol include/pbg.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "pbg.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(pbg_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    from pybullet_gym_b200 import _lib
    _lib.build_extension()
    L = ctypes.CDLL(_lib.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(L, s), "libpbg_b200.so does not export %s" % s
    assert L.pbg_version() == _lib.PBG_VERSION
    assert sorted(_lib.EXPORTS) == syms


def test_model_tables_fit_their_kernels():
    """pbg_create's host-side checks accept the tables of every backed env (create fails later, at the first
    CUDA call, on a box without a GPU -- with an error code, not a crash)."""
    from pybullet_gym_b200 import _lib
    from pybullet_gym_b200.spec import SPECS
    L = _lib.lib()
    for env_id, spec in SPECS.items():
        t = _lib.ModelTables(spec)
        h = ctypes.c_void_p()
        rc = L.pbg_create(ctypes.byref(t.c), 8, 0, 0, 0, ctypes.byref(h))
        if rc == 0:
            L.pbg_destroy(h)
        else:
            assert rc == -2, (env_id, rc, L.pbg_last_error(None))     # PBG_ERR_CUDA: no device here
        t.c.max_contacts += 1
        assert L.pbg_create(ctypes.byref(t.c), 8, 0, 0, 0, ctypes.byref(h)) == -1   # PBG_ERR_INVALID
        assert b"max_contacts" in L.pbg_last_error(None)


def test_no_cpu_fallback():
    import torch
    from pybullet_gym_b200 import _lib
    from pybullet_gym_b200.vector_env import VectorEnv
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.BackendUnavailable):
        VectorEnv("AntPyBulletEnv-v0", 4)


def test_unbacked_ids_raise():
    from pybullet_gym_b200.envs import make, registry
    assert "AntPyBulletEnv-v0" in registry and registry["AntPyBulletEnv-v0"]["max_episode_steps"] == 1000
    with pytest.raises(NotImplementedError):
        make("PusherPyBulletEnv-v0")


def test_argument_errors_need_no_device():
    """include/pbg.h error contract: bad arguments are rejected with PBG_ERR_INVALID / PBG_ERR_UNSUPPORTED before any CUDA
    call, NULL handles are errors and never dereferenced."""
    from pybullet_gym_b200 import _lib
    from pybullet_gym_b200.spec import SPECS
    L = _lib.lib()
    t = _lib.ModelTables(SPECS["HopperPyBulletEnv-v0"])
    h = ctypes.c_void_p()
    assert L.pbg_create(None, 8, 0, 0, 0, ctypes.byref(h)) == -1
    assert L.pbg_create(ctypes.byref(t.c), 8, 0, 0, 0, None) == -1
    for n in (0, -5):
        assert L.pbg_create(ctypes.byref(t.c), n, 0, 0, 0, ctypes.byref(h)) == -1
        assert b"bad arguments" in L.pbg_last_error(None)
    kind = t.c.kind
    t.c.kind = 99
    assert L.pbg_create(ctypes.byref(t.c), 8, 0, 0, 0, ctypes.byref(h)) == -3            # PBG_ERR_UNSUPPORTED
    t.c.kind = kind
    for f in (L.pbg_num_envs, L.pbg_obs_dim, L.pbg_action_dim, L.pbg_state_dim, L.pbg_noise_dim, L.pbg_last_host_path):
        assert f(None) == -1
    null = ctypes.c_void_p()
    assert L.pbg_step(None, null, null, null, null, null, null, null, null) == -1
    assert L.pbg_step_host(None, null, null, null, null) == -1
    assert L.pbg_reset(None, null, 0, null, null) == -1
    assert L.pbg_rollout_policy(None, 4, null, null, null, null) == -1
    assert L.pbg_stats(None, None, 0) == -1
    assert L.pbg_snapshot_bytes(None) == -1 and L.pbg_snapshot(None, null, null) == -1 and L.pbg_restore(None, null, null) == -1
    L.pbg_destroy(None)                                                                   # a no-op, not a crash
