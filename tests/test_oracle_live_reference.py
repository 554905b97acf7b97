"""Differential test of the env oracle -- and of the env step kernel's logic, compiled for the host (tests/emul) -- against the LIVE unmodified reference on random scenarios (1-5 radars, 1-4 jammers,
random geometry / gains / power ranges / radar types / reward bounds / episode limits): beyond the three committed golden
scenarios.  The reference runs in a subprocess (tests/golden/make_golden.py live ..., from baseline/_ref -- the git-ignored
copy __graft_entry__.build() makes where /root/reference exists), so its same-named top-level modules never enter this
process; skipped where no copy of the reference is present (the GPU box).  CPU only."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle.env_oracle import EnvOracle
from tests.conftest import ROOT
from tests.helpers import spec_for_golden
from tests.env_checks import check_env_against_golden
from tests.test_emul_env import make_env
from tests.test_oracle_env import check_step_sequences

REF = os.path.join(ROOT, "baseline", "_ref")
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "simulation")), reason="baseline/_ref is not present")


@pytest.mark.parametrize("seed0", [100, 200])
def test_oracle_vs_live_reference_on_random_scenarios(tmp_path, seed0):
    count = 6
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "golden", "make_golden.py"), "live", str(tmp_path), str(seed0), str(count)],
                       env=dict(os.environ, MAKE_GOLDEN_REF=REF), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    shapes = set()
    for i in range(count):
        g = np.load(os.path.join(tmp_path, f"env_live_{i}.npz"))
        cfg = json.loads(str(g["config_json"]))
        shapes.add((len(cfg["radars"]), len(cfg["jammers"])))
        env = EnvOracle(spec_for_golden(g, cfg))
        assert env.get_env_info() == json.loads(str(g["env_info_json"]))
        np.testing.assert_array_equal(env.reset()[0], g["state0"][0])
        np.testing.assert_array_equal(env.get_obs()[0], g["obs0"][0])
        np.testing.assert_array_equal(env.get_avail_actions()[0], g["avail0"][0])
        np.testing.assert_array_equal(env.gt[0], g["radar_gt_lin"])
        np.testing.assert_array_equal(env.pn[0], g["radar_pn_watts"])
        np.testing.assert_array_equal(env.gj[0], g["jammer_gj_lin"])
        np.testing.assert_allclose(env.echo_power()[0, :, 0], g["echo_ps"], rtol=1e-12)
        check_step_sequences(g, cfg)
        # the same recording through the product's env class and the step kernel compiled for the host (tests/emul):
        # integer outputs bit-exact, float64 reward 1e-10, float32 views to half an ulp (tests/env_checks.py)
        check_env_against_golden(make_env, (g, cfg))
        assert 0.2 < g["pd"].max() and g["terminated"].any() and (g["r_j"] > 0).any()       # the recording exercises the physics
    assert len(shapes) >= 3


@pytest.mark.parametrize("seed0", [7, 8])
def test_networks_vs_live_reference_at_random_dims(tmp_path, monkeypatch, seed0):
    """Agent / selector sequences, mixer forward + backward and whole QMixLearner.train steps recorded from the live reference
    at random network dims (1-6 agents, 2-16 actions, obs 5-60, GRU 64 / 128 / 192, actor 64 / 128, mixer embed 16-64,
    random batch shapes, ragged episodes, learning rates, target intervals): the eager oracle (tests/test_oracle_agent.py's
    checks) and the FP32 kernels compiled for the host (tests/agent_checks.py, tests/learner_checks.py) against them."""
    count = 3
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "golden", "make_golden.py"), "live_nets", str(tmp_path), str(seed0), str(count)],
                       env=dict(os.environ, MAKE_GOLDEN_REF=REF), capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    from tests import agent_checks as AC, learner_checks as LC, test_oracle_agent as TOA
    from tests.helpers import emul_lib
    for mod in (AC, LC, TOA):
        monkeypatch.setattr(mod, "GOLDEN", str(tmp_path))
    for i in range(count):
        name = f"live{i}"
        TOA.test_agent_and_selector(name)
        TOA.test_mixer_forward_and_gradients(name)
        TOA.test_learner_steps(name)
        AC.check_mac_against_golden(name, "cpu", emul_lib())
        AC.check_agent_outputs_against_golden(name, "cpu", emul_lib())
        LC.check_mixer_against_golden(name, "cpu", emul_lib())
        LC.check_learner_against_golden(name, "cpu", emul_lib())


def test_replay_ring_vs_live_reference_on_random_sequences(tmp_path, monkeypatch):
    """Random rings (capacity 2-11, episode limit 2-9, 1-4 agents, ragged episode lengths, up to three times the capacity
    stored, samples at random points incl. requests larger than the ring holds) through the live reference
    EpisodeReplayBuffer (utils/replay_buffer.py:78-214) and through the product ring on the host-emulated copy kernel:
    pointer and size after every store, gathered batches bit for bit, and sample() under the same np.random seed."""
    count = 8
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "golden", "make_golden.py"), "live_replay", str(tmp_path), "50", str(count)],
                       env=dict(os.environ, MAKE_GOLDEN_REF=REF), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    from tests import learner_checks as LC
    from tests.helpers import emul_lib
    monkeypatch.setattr(LC, "GOLDEN", str(tmp_path))
    for i in range(count):
        LC.check_replay_against_golden("cpu", emul_lib(), name=f"replay_live{i}")
