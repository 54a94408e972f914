"""bench.py contract on the CPU: the reference arm (the CPU implementation of the path on the host cores) prints one JSON line with the
keys the driver reads; under a multi-rank launch only rank 0 runs it."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(env_extra=None):
    env = dict(os.environ, **(env_extra or {}))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1", "--cells", "32"], capture_output=True, text=True,
                       env=env, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout.strip()


def test_reference_arm_prints_the_contract_line():
    out = run()
    line = json.loads(out.splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "DoFs/s" and line["higher_is_better"] is True
    assert line["steps"] == 2 and line["warmup"] == 3 and line["value"] > 0 and line["ms_per_step"] > 0
    assert line["dtype"] == "f64" and line["data"] == "synthetic" and line["vs_baseline"] is None and "workload" in line["config"]
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and "vectorised" in cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "DoFs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_without_work():
    assert run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}) == ""
