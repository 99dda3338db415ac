"""The reference arm of bench.py (`--impl reference`: the CPU oracle timed on the host cores) runs without a GPU, so its
contract is checked here: one JSON line with the keys the driver reads, the metric / config of the B200 arm, and a
`cpu_baseline` describing the run."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    d = json.loads(lines[0])
    with open(os.path.join(ROOT, "BASELINE.json")) as f:
        base = json.load(f)
    assert d["impl"] == "reference" and d["metric"] == base["metric"]
    assert d["unit"] == "molecules/s" and d["higher_is_better"] is True and d["n_gpus"] == 1
    assert d["steps"] == 1 and d["warmup"] == 1 and d["ms_per_step"] > 0 and d["value"] > 0
    assert d["gpu_launches"] == 0 and d["vs_baseline"] is None
    assert "workload" in d["config"] and "configs[2]" in d["config"]["workload"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gin"]["value"] > 0 and d["gin"]["unit"] == "graphs/s"
