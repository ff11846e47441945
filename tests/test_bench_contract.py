"""The reference arm of bench.py on this host's CPU (no GPU needed): one JSON line on stdout with the keys the driver
reads, the same metric / unit / workload as our own arm, and the unmodified reference (oracle/_ref) or the port behind
it — never a library of this repo (`gpu_launches` 0)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines                      # stdout carries exactly the JSON line
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "vq_latents_per_sec" and d["unit"] == "latents/s"
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["dtype"] == "f32" and d["data"] == "synthetic" and d["gpu_launches"] == 0
    assert d["config"]["latents_per_step_per_gpu"] == 18432 + 76800 and d["config"]["codebook"] == [32, 128]
    assert d["value"] > 0 and abs(d["value"] - 95232 / (d["ms_per_step"] * 1e-3)) < 1e-6 * d["value"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
