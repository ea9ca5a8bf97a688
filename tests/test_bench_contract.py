"""The driver-facing contract of bench.py that can be checked without a GPU: the reference arm (the reference's own
PocketFFT path from oracle/_ref on the host cores) prints ONE JSON line with the agreed keys."""
import json
import os
import subprocess
import sys

import pytest

import parity


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    if parity.reflib() is None:
        pytest.skip("oracle/_ref not built")
    r = subprocess.run([sys.executable, os.path.join(parity.ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=parity.ROOT)
    assert r.returncode == 0, r.stderr[-800:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "GFLOP/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["gpu_launches"] == 0
    assert d["config"]["workload"].startswith("cfg2") and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] == (os.cpu_count() or 1) and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_non_zero_ranks_of_the_reference_arm_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(parity.ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       capture_output=True, text=True, timeout=120, cwd=parity.ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""
