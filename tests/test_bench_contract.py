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


def test_committed_gpu_line_carries_the_contract_keys():
    """The N = 1 line committed under profiles/ (the run the round's summary quotes) has every key the
    measurement contract names, and its roofline object is self-consistent."""
    with open(os.path.join(ROOT, "profiles", "r2_bench_n1.json")) as f:
        line = json.loads(f.read().strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "clocks", "cpu_baseline"):
        assert key in line, key
    assert line["n_gpus"] == 1 and line["gpu_launches"] > 0 and line["vs_baseline"] is None
    assert "workload" in line["config"] and "model" not in line["config"]
    roof = line["roofline"]
    for key in ("bound", "achieved", "peak", "unit", "frac", "traffic", "kernel_ms", "event_floor_ms", "kernel_ms_marginal"):
        assert key in roof, key
    assert roof["bound"] == "tensor" and abs(roof["frac"] - roof["achieved"] / roof["peak"]) < 1e-9
    assert abs(roof["achieved"] - roof["algorithmic_flops_per_launch"] / (roof["kernel_ms"] * 1e-3) / 1e12) < 1e-6 * roof["achieved"]
    e2e = line["e2e"]
    assert e2e["h2d_bytes_per_step"] == 2 * 4096 * 256 * 4 and e2e["d2h_bytes_per_step"] == 4 and e2e["value"] < line["value"]
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert not set(line["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    for extra in ("c1", "c3", "c5", "retrieval", "siglip", "n1"):
        assert extra in line, extra
    assert {"roofline", "cpu_baseline", "e2e"} <= set(line["retrieval"])
