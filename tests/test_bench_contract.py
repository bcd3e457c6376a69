"""bench.py on a machine without a GPU: the reference arm prints the contract's JSON line (the CPU
implementation of the reference timed on a bounded sample), the product arm refuses to run."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], cwd=ROOT, capture_output=True, text=True,
                          timeout=900)


def test_reference_arm_prints_the_contract_line(oracle_lib):
    proc = _run("--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-structures", "6")
    assert proc.returncode == 0, proc.stderr[-2000:]
    line = json.loads(proc.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "template-structure pairs/sec" and line["unit"] == "pairs/s"
    assert line["n_gpus"] == 1 and line["steps"] == 1 and line["warmup"] == 0 and line["higher_is_better"] is True
    assert line["scaling"] == "weak" and line["vs_baseline"] is None and line["dtype"] == "f64" and line["data"] == "synthetic"
    assert line["config"]["workload"] == "cfg3" and line["config"]["templates"] == 1000
    assert line["config"]["structures_per_gpu"] == 1000 and line["sample_structures"] == 6   # the arm's config; the sample is stated
    assert line["engine_library_mapped"] is False          # inputs packed with numpy: the product's .so never enters this process
    assert line["value"] > 0 and line["ms_per_step"] > 0
    assert abs(line["value"] - 6000 / (line["ms_per_step"] / 1e3)) / line["value"] < 1e-6
    base = line["cpu_baseline"]
    assert base["kind"] in ("reference", "port") and base["cores"] >= 1 and base["value"] == line["value"] and base["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_product_arm_refuses_to_run_without_a_gpu():
    from pyjess_b200 import _capi
    if _capi.load_library().jb_device_count() > 0:
        pytest.skip("a GPU is present")
    proc = _run("--steps", "1", "--warmup", "1", "--no-cpu-baseline")
    assert proc.returncode != 0
    assert "no CPU fallback" in (proc.stderr + proc.stdout)
    assert not any(l.startswith("{") for l in proc.stdout.splitlines())        # no result line
