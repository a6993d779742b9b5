"""bench.py contract checks that need no GPU: the reference arm (the oracle port on the host cores)
prints one JSON line with the agreed keys on rank 0 and nothing on the other ranks."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def _run(extra_env):
    env = dict(os.environ, **extra_env)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus",
                           extra_env.get("WORLD_SIZE", "1"), "--steps", "1", "--warmup", "3"],
                          capture_output=True, text=True, env=env, timeout=600)


def test_reference_arm_prints_the_contract_line():
    r = _run({})
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "InfoNCE fwd+bwd pairs/s" and line["unit"] == "pairs/s"
    assert line["higher_is_better"] is True and line["value"] > 0 and line["n_gpus"] == 1
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["cpu_baseline"]["value"] == line["value"] == line["e2e"]["value"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["config"]["d"] == 256 and line["config"]["global_batch"] == 4096


def test_reference_arm_other_ranks_exit_quietly():
    r = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""
