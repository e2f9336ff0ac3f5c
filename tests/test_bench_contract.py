"""bench.py contract checks that run without a GPU: the reference arm (CPU oracle) on a tiny configuration must print exactly one
JSON line with the keys the driver reads, and the precision constants of the header and of the ctypes binding must agree."""
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "1", "--warmup", "0",
           "--n-rows", "4096", "--d-in", "16", "--num-rf", "2", "--block", "64", "--classes", "5", "--cpu-rows", "256"]
    env = dict(os.environ, OMP_NUM_THREADS="1", RANK="0")   # what torchrun exports
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 1 and d["higher_is_better"] is True
    for key in ("metric", "value", "unit", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data", "config",
                "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["unit"] == "samples/s" and d["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and "sample" in cb and cb["value"] == d["value"]
    assert cb["per_row_seconds"] > 0 and cb["solve_seconds"] > 0 and cb["sample_rows"] <= 256
    assert cb["blas_threads"] == str(os.cpu_count())       # the CPU arm uses every host core even under torchrun (OMP_NUM_THREADS=1)
    assert d["config"]["workload"].startswith("C3 ")


def test_precision_constants_agree_between_header_and_binding():
    from keystone_b200 import _capi
    hdr = open(os.path.join(ROOT, "include", "keystone_b200.h")).read()
    consts = dict(re.findall(r"#define\s+(KS_PRECISION_[A-Z0-9]+)\s+\(?(-?\d+)\)?", hdr))
    assert set(consts) == {"KS_PRECISION_DEFAULT", "KS_PRECISION_TF32", "KS_PRECISION_F16", "KS_PRECISION_F16X2"}
    for name, val in consts.items():
        assert getattr(_capi, name) == int(val), name
