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
