"""bench.py's reference arm (the CPU port timed on the host cores): the JSON line carries every key the
bench contract names, and under a multi-rank launch only rank 0 prints.  No GPU involved."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env):
    env = dict(os.environ, MV_BENCH_REF_BUDGET_S="0.3", **extra_env)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus",
                        extra_env.get("WORLD_SIZE", "1"), "--steps", "2", "--warmup", "1"],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    return [json.loads(l) for l in r.stdout.splitlines() if l.startswith("{")]


def test_reference_arm_line_has_the_contract_keys():
    lines = _run({})
    assert len(lines) == 1
    d = lines[0]
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["data"] == "synthetic"
    assert d["metric"].startswith("frame-pairs/sec") and d["unit"] == "frame-pairs/s"
    assert d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == 1 and d["vs_baseline"] is None
    assert d["value"] > 0 and d["ms_per_step"] > 0 and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_print_nothing():
    assert _run({"WORLD_SIZE": "2", "RANK": "1", "LOCAL_RANK": "1"}) == []
