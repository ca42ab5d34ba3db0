"""bench.py's CPU-runnable arm keeps the driver's contract: `--impl reference` prints ONE JSON line with the base
keys, `"impl": "reference"`, a cpu_baseline describing the run and an e2e object with zero transfer bytes; the
argument parser accepts the driver's flags. (The GPU arm's line is produced on the B200 box: profiles/r01_bench_*.json.)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "2", "--warmup", "3",
                        "--cpu-budget-s", "4"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config"):
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "rk_steps_per_sec" and d["dtype"] == "f64" and d["vs_baseline"] is None
    assert d["config"]["workload"] == "cfg2_dopri54_diag_8M" and d["steps"] == 2 and d["warmup"] == 3
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": "RK steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["value"] > 0 and d["gpu_launches"] == 0
    # both arms print the same `unit` and a `config` with identical keys AND values (the driver divides the two lines and
    # compares their configs): the workload only — library knobs and notes live outside `config`
    sys.path.insert(0, ROOT)
    import bench
    assert d["unit"] == bench.UNIT == "RK steps/s"
    assert d["config"] == bench.base_config("cfg2_dopri54_diag_8M", 23)
    assert set(d["config"]) == {"workload", "integrator", "rhs", "elems_per_gpu", "options"} and d["knobs"] == {}
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert src.count('"config": base_config(') == 3   # our line, the cfg3 / cfg4 objects, the reference line


def test_non_zero_ranks_of_the_reference_arm_exit_quietly():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "2", "--warmup", "3"],
                       capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_gpu_arm_fails_loudly_without_a_device():
    """No CPU fallback: without a CUDA device the product arm must raise, not print a number."""
    try:
        import torch
        if torch.cuda.is_available():
            return
    except Exception:
        pass
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "2", "--warmup", "3", "--no-cpu-baseline"], capture_output=True, text=True,
                       timeout=300, cwd=ROOT)
    assert r.returncode != 0 and not any(l.startswith('{"metric"') for l in r.stdout.splitlines())
