"""bench.py's output contract on the CPU arm (no GPU): exactly one JSON line on stdout, carrying the
keys the driver reads, also under torchrun's environment for a rank that must stay silent."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
        "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches")


def run_bench(extra_env=None, *flags):
    env = dict(os.environ)
    env.update(extra_env or {})
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "3", *flags], cwd=ROOT, env=env, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr[-2000:]
    return p.stdout


def test_reference_arm_prints_one_json_line():
    out = run_bench()
    lines = out.splitlines()
    assert len(lines) == 1, out
    d = json.loads(lines[0])
    for k in KEYS:
        assert k in d, k
    assert d["impl"] == "reference" and d["gpu_launches"] == 0
    assert d["higher_is_better"] is True and d["unit"] == "MSamples/s" and d["value"] > 0
    assert d["config"]["workload"].startswith("cfg2")
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"]
    assert cb["single_thread_value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": "MSamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_stay_silent():
    out = run_bench({"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"}, "--gpus", "2")
    assert out == ""


def test_library_chatter_on_fd1_does_not_reach_stdout():
    """Anything a library writes to file descriptor 1 while the arms run (NCCL's version banner does
    under torchrun) must not end up next to the JSON line."""
    code = ("import os, sys, json; sys.path.insert(0, %r); import bench\n"
            "real = bench.claim_stdout()\n"
            "os.write(1, b'NCCL version 0.0\\n')\n"
            "bench.emit(real, json.dumps({'ok': 1}) + '\\n')\n") % ROOT
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stderr
    assert p.stdout == '{"ok": 1}\n'
    assert "NCCL version 0.0" in p.stderr
