"""bench.py's CPU legs (no GPU): the all-cores leg on the unmodified reference classes steps the same workload as
a single-process loop over the same classes, and the `--impl reference` line has the contract's keys.  Needs
baseline/_ref (a git-ignored copy of /root/reference that __graft_entry__.build() makes); skipped without it."""
import json
import os
import subprocess
import sys

import pytest

from tests.conftest import ROOT

HAVE_REF = os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "simulation"))
pytestmark = pytest.mark.skipif(not HAVE_REF, reason="baseline/_ref is not present")


def test_all_cores_leg_runs_the_unmodified_reference():
    import bench
    cwd, path = os.getcwd(), list(sys.path)
    r = bench.cpu_reference_parallel_run(37, 4, 1)          # ragged share over the workers
    assert os.getcwd() == cwd and sys.path == path            # the leg puts the process back as it found it
    assert "unavailable" not in r, r
    assert r["kind"] == "reference" and r["unit"] == "env-agent steps/s"
    assert r["value"] > 0 and 1 <= r["cores"] <= 37
    assert abs(r["value"] - 37 * 2 * 1e3 / r["ms_per_step"]) < 1e-6 * r["value"]
    # the product package is untouched by the reference's same-named modules (core, simulation, utils, runners)
    assert not any(k.split(".")[0] in ("core", "simulation", "utils", "runners") for k in sys.modules)


def test_reference_arm_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "4", "--warmup", "1",
                          "--n-envs", "64"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "env_agent_steps_per_sec" and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["gpu_launches"] == 0
    assert d["port"]["kind"] == "port" and d["port"]["value"] > 0
    assert d["config"]["n_envs_per_gpu"] == 64


def test_other_ranks_of_the_reference_arm_exit_without_work():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""
