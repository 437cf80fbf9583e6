"""bench.py prints exactly one JSON line on stdout with the keys of the contract (checked here on the CPU arm, which needs
no GPU: `--impl reference` times the oracle port on the host cores)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    proc = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1"],
                          capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert proc.returncode == 0, proc.stderr[-2000:]
    lines = [ln for ln in proc.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, proc.stdout[:500]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("512x512 frames/sec") and d["value"] > 0 and d["n_gpus"] == 1
    assert d["steps"] == 2 and d["warmup"] == 1                       # --steps / --warmup are honoured as given
    # same workload description as the product arm prints (bench.config_dict), so the driver can pair the two lines
    sys.path.insert(0, ROOT)
    import bench
    import types
    args = types.SimpleNamespace(variant="large", batch=32, height=512, width=512, mode="parity")
    assert d["config"] == bench.config_dict(args, 1)
    assert d["cpu_baseline"]["host"]["usable_cpus"] >= 1 and "sample" in d["cpu_baseline"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_product_arm_refuses_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    proc = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], capture_output=True, text=True,
                          timeout=600, cwd=ROOT)
    assert proc.returncode != 0 and "no CPU path" in (proc.stderr + proc.stdout)
    assert proc.stdout.strip() == ""
