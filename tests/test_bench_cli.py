"""CPU: the reference arm of bench.py (`--impl reference`) runs the reference's own code on the host and prints ONE JSON line
with the contract's keys, for a cheap configuration; and every configuration of the CUDA arm is declared consistently."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line_for_ssd300():
    env = dict(os.environ)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "ssd300", "--steps", "1",
                        "--warmup", "0", "--batch", "4"], capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "images/s" and d["value"] > 0 and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert d["config"]["workload"] == "ssd300_coco_bs32"


def test_every_baseline_configuration_is_a_bench_config():
    sys.path.insert(0, ROOT)
    import bench
    assert set(bench.CONFIGS) == {"headline", "cfg1", "cfg2", "ssd300", "retina800", "cfg4", "crowd512"}
    for name, cfg in bench.CONFIGS.items():
        assert cfg["kind"] in ("yolo", "prior", "targets") and cfg["metric"] and cfg["batch"] >= 1
        c = bench.base_config(cfg, cfg["batch"], 8)
        assert c["workload"] == cfg["name"] and c["global_batch"] == 8 * cfg["batch"]
    assert bench.CONFIGS["crowd512"]["batch"] * 8 == 512 and bench.CONFIGS["crowd512"]["conf_thres"] == 0.001
    assert bench.CONFIGS["headline"]["batch"] == 64 and bench.CONFIGS["headline"]["classes"] == 80
