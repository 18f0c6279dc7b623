"""bench.py's own arm on a small batch: the JSON line carries every key of the contract and the numbers are self-consistent."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_bench_json_line_contract(tmp_path):
    env = dict(os.environ, MAB_BENCH_DIR=str(tmp_path))
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "4", "--warmup", "3", "--batch-reads", "768", "--contexts", "2", "--no-cpu-baseline"],
                       capture_output=True, text=True, env=env, timeout=900)
    assert p.returncode == 0, p.stderr[-800:]
    lines = [l for l in p.stdout.split("\n") if l.strip()]
    assert len(lines) == 1, p.stdout[-400:]
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
              "config", "clocks", "e2e", "gpu_launches", "roofline"):
        assert k in d, k
    assert d["metric"] == "Mbases aligned/sec" and d["unit"] == "Mbases/s" and d["n_gpus"] == 1 and d["steps"] == 4 and d["warmup"] == 3
    assert d["scaling"] == "weak" and d["vs_baseline"] is None and d["higher_is_better"] is True and d["data"] == "synthetic"
    assert d["value"] > 0 and d["e2e"]["value"] > 0 and d["e2e"]["h2d_bytes_per_step"] > 768 * 10000 and d["e2e"]["d2h_bytes_per_step"] > 0
    assert d["gpu_launches"] >= 4 * 10                                       # reader, scan, expand, sortchain, extend, post, SAM kernels per step
    assert d["e2e"]["d2h_bytes_per_step"] >= d["e2e"]["sam_bytes_per_step"] > d["e2e"]["h2d_bytes_per_step"]      # text out (with the bases) > text in
    r = d["roofline"]
    assert r["bound"] == "int-alu" and r["unit"] == "GCUPS" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["ms_per_launch"] > 0
    assert 0 < r["frac"] <= 1.0 and r["peak"] > 300 and r["hbm"]["unit"] == "GB/s" and 0 < r["hbm"]["frac"] < 1
    assert d["parity"]["timed_equals_fresh_from_second_read"] is True and d["parity"]["ok"] in (True, None)
    assert d["config"]["batch_reads"] == 768 and d["config"]["contexts_per_gpu"] == 2
    assert "sm_mhz" in d["clocks"] and "reasons" in d["clocks"]
