"""The reference's OWN test-suite (/root/reference/tests, 5 files) executed on the CPU oracle engine through the fake
`mujoco` / `gymnasium` modules of oracle/fake_mujoco.py: the unmodified reference package (PickPlaceGymEnv, IKController,
PickAndPlaceTask, randomization, pose utilities) runs on top of the engine restatement and must pass its own behavioural
bounds.  Only possible where the reference checkout exists (the build container); skipped elsewhere (GPU box)."""
import os
import subprocess
import sys

import pytest

REF = "/root/reference"
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "tests")), reason="reference checkout not present")
def test_reference_tests_pass_on_the_oracle_engine(tmp_path):
    env = dict(os.environ, PYTHONPATH=REPO + os.pathsep + os.environ.get("PYTHONPATH", ""), PYTHONDONTWRITEBYTECODE="1")
    cmd = [sys.executable, "-m", "pytest", os.path.join(REF, "tests"), "-p", "oracle.refshim_plugin", "-q", "-p", "no:cacheprovider",
           "--rootdir", str(tmp_path), "-c", os.devnull]
    out = subprocess.run(cmd, cwd=REF, env=env, capture_output=True, text=True, timeout=1500)  # the suite opens the scene XML by a path relative to the checkout
    tail = out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-400:]
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-1000:]
    assert " passed" in tail and "failed" not in tail, tail
    n = int(tail.split(" passed")[0].split()[-1])
    assert n >= 200, tail  # 201 test cases at the surveyed commit
