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
    r = subprocess.run([sys.executable, os.path.join(parity.ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                        "--images", "1"], capture_output=True, text=True, timeout=600, cwd=parity.ROOT)
    assert r.returncode == 0, r.stderr[-800:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "GFLOP/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["gpu_launches"] == 0
    assert d["config"]["workload"].startswith("cfg2") and "model" not in d["config"]
    assert d["steps"] == 2 and d["warmup"] == 1 and d["config"]["images_per_gpu"] == 1  # the arguments are honoured
    cb = d["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] == (os.cpu_count() or 1) and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_non_zero_ranks_of_the_reference_arm_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(parity.ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       capture_output=True, text=True, timeout=120, cwd=parity.ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_the_reference_arm_never_maps_the_product_library():
    """bench.py --impl reference must run on reference code alone: importing what it imports must not load
    librocketfft_b200.so into the process (the driver lists the shared objects each arm mapped)."""
    code = (
        "import sys, os; sys.path.insert(0, %r); os.chdir(%r)\n"
        "import bench\n"
        "ref, kind, cores = bench.reference_lib()\n"
        "maps = open('/proc/self/maps').read()\n"
        "assert 'librocketfft_b200' not in maps, 'product library mapped'\n"
        "assert 'rocket_fft_b200' not in sys.modules\n"
        "print(kind)\n" % (parity.ROOT, parity.ROOT)
    )
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-800:]
