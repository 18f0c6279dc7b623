"""bench.py --impl reference runs on the host alone (oracle/_ref/minialign): check its JSON line here, and that the
other ranks of a torchrun launch leave without work or output."""
import json
import os
import subprocess
import sys

import pytest

import refh

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(not os.path.exists(refh.BIN), reason="oracle/_ref/minialign not built")


def run(env_extra, tmp_path):
    env = dict(os.environ, MAB_BENCH_DIR=str(tmp_path), **env_extra)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", "--batch-reads", "96"],
                          capture_output=True, text=True, env=env, timeout=600)


def test_reference_arm_json_line(tmp_path):
    p = run({}, tmp_path)
    assert p.returncode == 0, p.stderr[-500:]
    lines = [l for l in p.stdout.split("\n") if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "Mbases aligned/sec" and d["unit"] == "Mbases/s" and d["higher_is_better"] is True
    assert d["steps"] == 1 and d["warmup"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "Mbases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["batch_reads"] == 96 and "workload" in d["config"]


def test_reference_arm_other_ranks_do_nothing(tmp_path):
    p = run({"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"}, tmp_path)
    assert p.returncode == 0 and p.stdout.strip() == ""
