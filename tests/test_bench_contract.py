"""bench.py's reference arm runs on the CPU (numpy port of the reference's update): check the JSON-line contract and
that the product package never imports the oracle."""
import json
import os
import subprocess
import sys

from ts_testutil import ROOT


def test_reference_arm_prints_one_contract_line():
    env = dict(os.environ, TS_BENCH_CPU_ENVS="8")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, env=env, timeout=600, check=True).stdout.strip().splitlines()
    line = json.loads(out[-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["impl"] == "reference" and line["value"] > 0 and line["unit"] == "transitions/s"
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert "workload" in line["config"]


def test_reference_arm_is_silent_on_other_ranks():
    env = dict(os.environ, TS_BENCH_CPU_ENVS="8", RANK="1", WORLD_SIZE="2")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, env=env, timeout=120, check=True)
    assert res.stdout.strip() == ""


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "tianshou_b200")
    offenders = []
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(d, f)).read()
                if "import oracle" in src or "from oracle" in src:
                    offenders.append(os.path.join(d, f))
    assert not offenders, offenders
