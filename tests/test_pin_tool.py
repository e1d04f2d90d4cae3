"""Self-consistency of tools/pin_pybullet.py (the "pin day" tool of SURVEY.md Appendix C6).

`dump` is run on the stub pybullet client (tools/fake_pybullet.py: the reference's own env classes on top of the oracle's
physics), `compare` must then find our compiler's tables identical to the dumped ones and replay the tape.  This checks the
tool's plumbing (field order, state mapping, one-step replay) -- it is NOT a pin of the physics: that needs a real pybullet.
Needs /root/reference (the build container); skipped elsewhere.
"""
import os
import sys

import numpy as np
import pytest

REF = "/root/reference"
TOOLS = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools")


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference checkout is only present in the build container")
def test_dump_on_the_stub_client_compares_clean(tmp_path, oracle_lib):
    sys.path.insert(0, TOOLS)
    import fake_pybullet as fp
    import pin_pybullet as pin
    fp.install()
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import contextlib
    import importlib
    import io
    from pybullet_gym_b200.spec import SPECS

    def importer(env_id, mod, cls):
        fp.FakeBulletClient.current_spec = SPECS[env_id]
        fp.FakeBulletClient.max_contacts = 0
        m = importlib.import_module("pybulletgym.envs.roboschool." + mod)

        def make():
            with contextlib.redirect_stdout(io.StringIO()):
                return getattr(m, cls)()
        return make

    path = str(tmp_path / "pins.json")
    ids = ["InvertedPendulumPyBulletEnv-v0", "HopperPyBulletEnv-v0", "AntPyBulletEnv-v0"]
    with contextlib.redirect_stdout(io.StringIO()):
        pin.dump(path, envs=ids, steps=25, importer=importer)
    report, bad = pin.compare(path, verbose=False)
    assert bad == 0, {k: v["table_rows"] for k, v in report.items()}
    assert set(report) == set(ids)
    # contact-free: the one-step replay is exact; with contacts only the warm-start impulses (not part of the state) differ
    assert report["InvertedPendulumPyBulletEnv-v0"]["one_step_max"] < 1e-12
    assert report["AntPyBulletEnv-v0"]["one_step_median"] < 1e-2 and report["HopperPyBulletEnv-v0"]["one_step_median"] < 1e-2


def _real_pybullet():
    """True only for a real pybullet wheel (the stub of tools/fake_pybullet.py marks itself)."""
    try:
        import pybullet
        import gym  # noqa: F401
    except Exception:
        return False
    return not getattr(pybullet, "IS_PBG_STUB", False) and hasattr(pybullet, "connect")


@pytest.mark.skipif(not _real_pybullet(), reason="pin day: needs a real pybullet wheel + gym (not installable in the build image)")
def test_pin_day_compare_against_real_pybullet(tmp_path, oracle_lib):
    """The day `import pybullet` works this test IS the physics pin: dump the reference's own envs (link / joint / dynamics
    tables, engine parameters, a state-by-state rollout tape) and require (a) the MJCF compiler's tables to match what
    loadMJCF built, (b) the oracle's one-step replay from every dumped state to stay within the single-step tier."""
    sys.path.insert(0, TOOLS)
    for ref in (os.path.join(os.path.dirname(TOOLS), "baseline", "_ref"), REF):
        if os.path.isdir(ref) and ref not in sys.path:
            sys.path.insert(0, ref)
    import pin_pybullet as pin
    path = str(tmp_path / "pins.json")
    pin.dump(path)
    report, bad = pin.compare(path, verbose=True)
    assert bad == 0, {k: v["table_rows"] for k, v in report.items()}
    for env_id, rep in report.items():
        assert rep["one_step_median"] < 1e-3, (env_id, rep["one_step_median"], rep["one_step_max"])
