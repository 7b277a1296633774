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
    assert d["config"]["workload"].startswith("cfg3")       # the largest single-GPU config is the line's workload
    assert "l2" not in d["config"] and "input" not in d["config"]   # nothing arm-specific in `config`
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"]
    assert cb["single_thread_value"] > 0
    assert cb["seconds"] >= 2.0 and "synth.lattice_noise" in cb["sample"]      # no CPU sample shorter than 2 s
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


def test_reference_arm_does_not_map_the_product_library():
    """The CPU arm designs its taps with the oracle: libwebradio_b200.so never enters that process."""
    code = ("import sys, os; sys.path.insert(0, %r); sys.argv = ['bench.py', '--impl', 'reference', '--steps', '1']\n"
            "import bench\n"
            "bench.MIN_CPU_SECONDS = 0.05\n"
            "bench.reference_arm(bench.parse_args(), bench.synth.WORKLOADS['cfg2'], 'cfg2')\n"
            "maps = open('/proc/self/maps').read()\n"
            "assert 'libwebradio_b200' not in maps, 'product library mapped in the reference arm'\n"
            "assert 'libwr_ref' in maps or 'libwr_oracle' in maps\n") % ROOT
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]


def test_gpu_and_reference_arm_name_the_same_config():
    sys.path.insert(0, ROOT)
    import bench
    for name, w in bench.synth.WORKLOADS.items():
        c = bench.bench_config(w, name)
        assert set(c) == {"workload", "sample_rate", "frames_per_step", "n_receivers", "n_streams", "channel_fir",
                          "audio_fir", "modes", "parallelism"}
    # the CPU arms of a large bank run an evenly spaced subset of whole streams
    ids = bench.cpu_sample_receivers(bench.synth.WORKLOADS["cfg3"])
    assert len(ids) == 64 and ids[0] == 0 and ids[-1] == 1008 and len(set(ids)) == 64
    ids5 = bench.cpu_sample_receivers(bench.synth.WORKLOADS["cfg5"])
    assert ids5 == list(range(64))
    assert bench.cpu_sample_receivers(bench.synth.WORKLOADS["cfg2"]) == list(range(64))
    # SURVEY.md 8d algorithmic bytes
    assert bench.algorithmic_bytes(bench.synth.WORKLOADS["cfg3"]) == 8 * 102400 * 1024 + 4 * 1024 * 2048
    assert bench.algorithmic_bytes(bench.synth.WORKLOADS["cfg2"]) == 819200 + 524288


def test_in_run_parity_check_covers_every_receiver_and_refuses_a_wrong_block(wro):
    """bench.parity_check: the whole bank's block 0 against the oracle, spread over the host threads; a single
    wrong audio sample anywhere means no number is reported."""
    import numpy as np
    import pytest
    sys.path.insert(0, ROOT)
    import bench
    from webradio_b200 import synth
    w = synth.WORKLOADS["cfg2"]
    R, F = w["n_rx"], w["frames"]
    rng = np.random.default_rng(3)
    t1 = (rng.uniform(-1, 1, w["n1"]) / w["n1"]).astype(np.float32)
    t2 = (rng.uniform(-1, 1, w["n2"]) / w["n2"]).astype(np.float32)
    ifs, modes = synth.workload_ifs(w), synth.workload_modes(w)
    block = synth.lattice_noise(F, stream=11)
    audio = np.stack([wro.Rx(w["fs"], int(ifs[r]), t1, w["d1"], int(modes[r]), t2, w["d2"]).process(block) for r in range(R)])
    rec = bench.parity_check(w, audio, np.arange(R), 11, t1, t2)
    assert rec["receivers_checked"] == R == rec["receivers_in_bank"] and rec["bit_exact"] and rec["receivers_differing"] == 0
    audio[R - 1, 17] += np.float32(1e-3)
    with pytest.raises(SystemExit):
        bench.parity_check(w, audio, np.arange(R), 11, t1, t2)
