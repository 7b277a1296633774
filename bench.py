#!/usr/bin/env python
"""bench.py -- throughput of the per-receiver DSP hot path (NCO mix -> decimating FIR -> demod ->
audio FIR) on B200, next to the reference's own CPU path.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg3] [--subs cfg2,cfg4,cfg5]
                  [--impl reference]

The line's workload is cfg3 (1024 independent streams, 255 taps: the largest single-GPU config of
BASELINE.json and the HBM-roofline one) at every N -- weak scaling, each rank owns its
own tuners.  One "step" = one tuner block of the workload through every receiver of this rank's bank.
Prints ONE JSON line (rank 0).  Keys beyond the base contract:
  roofline     dominant kernel: ALGORITHMIC bytes of SURVEY.md 8d / its CUDA-event time vs the measured HBM peak
  cpu_baseline the reference's CPU chain (oracle/_ref, else the C port) on the host cores, same inputs
  e2e          the same metric through the synchronous C-ABI call the DspBlock drop-in makes, HOST buffers
               (H2D + kernels + D2H inside every call); pipelined / raw-byte / plug-in figures beside it
  parity       one output block of this run compared with the oracle inside the run
  configs      the same record for cfg2, cfg4 (spectrum) and this GPU's share of cfg5
Receivers are independent: ranks never exchange data (weak scaling, no collective on the path).
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "harness"))

import numpy as np  # noqa: E402

from webradio_b200 import synth  # noqa: E402

METRIC = "input IQ MSamples/s through downconvert→FIR→demod; achieved HBM GB/s vs peak"
L2_BYTES = 126 * 1024 * 1024
HBM_FALLBACK_GBS = 6650.0
MIN_CPU_SECONDS = 2.0          # no CPU sample shorter than this is reported
CPU_SAMPLE_RX = 64             # receivers of a large workload the CPU arms run (stated in `sample`)
TRAFFIC_FILE = os.path.join("profiles", "traffic.json")


def parse_args(argv=None):
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=None)
    p.add_argument("--warmup", type=int, default=None)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--workload", default="cfg3", choices=sorted(synth.WORKLOADS) + ["cfg4"])
    p.add_argument("--subs", default=None,
                   help="comma-separated workloads reported under `configs` (default: cfg2,cfg4,cfg5 next to cfg3; "
                        "'none' for a single record)")
    p.add_argument("--variant", type=int, default=0, help="kernel family: 0 auto, 1 v1, 2 v2, 3 v3, 4 v4")
    p.add_argument("--input", default="f32", choices=["f32", "u8"],
                   help="tuner block format of the DEVICE-timed legs: interleaved float IQ (the DspBlock convention) "
                        "or raw RTL-SDR bytes converted inside the channel kernel's load (SURVEY.md 8f-1); "
                        "e2e reports both")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--cpu-seconds", type=float, default=5.0, help="target wall time of each CPU baseline sample")
    return p.parse_args(argv)


# ------------------------------------------------------------------ workload ----

def workload_taps(w, design=None):
    """Tap values are an input of the FIR (the reference's own design collapses to all-zero for the
    narrow pass-bands of cfg2/3/5: maxbin = N*passband/Fs/2 = 0, lowpass.cxx:167).  64-tap stages use
    the reference design where it is not degenerate; the others a Hamming windowed-sinc of the same
    length.  `design` is the routine that restates LowPass::recalculate: the product's on the GPU
    arm, the oracle's on the reference arm (so that arm never maps the product library)."""
    if design is None:
        from webradio_b200 import capi
        design = capi.lowpass_design
    t1 = design(64, w["pb1"], w["fs"]) if w["n1"] == 64 else np.zeros(0, np.float32)
    if not np.any(t1):
        t1 = synth.windowed_sinc(w["n1"], w["pb1"] / w["fs"])
    t2 = design(w["n2"], w["pb2"], w["fs"] // w["d1"])
    if not np.any(t2):
        t2 = synth.windowed_sinc(w["n2"], w["pb2"] / (w["fs"] // w["d1"]))
    return t1, t2


def algorithmic_bytes(w, frame_bytes=8):
    """B_alg = 8*F*T + 4*R*F/(D1*D2) per block (T = unique input streams); 2 bytes per
    frame instead of 8 when the tuner block arrives as raw RTL-SDR bytes."""
    F, T, R = w["frames"], w["n_streams"], w["n_rx"]
    return frame_bytes * F * T + 4 * R * (F // w["d1"] // w["d2"])


def chan_kernel_bytes(w, variant, frame_bytes=8):
    """What the dominant kernel itself moves: the tuner block(s) once and, per channel-rate sample
    per receiver, the demodulated float (v1: demod fused in) or the IQ pair (v2, v3, v4)."""
    F, T, R = w["frames"], w["n_streams"], w["n_rx"]
    return frame_bytes * F * T + (8 if variant >= 2 else 4) * R * (F // w["d1"])


def bench_config(w, wname):
    """Identical for both arms (the driver compares them)."""
    return {"workload": f"{wname}: {w['desc']}", "sample_rate": w["fs"], "frames_per_step": w["frames"],
            "n_receivers": w["n_rx"], "n_streams": w["n_streams"], "channel_fir": [w["n1"], w["d1"]],
            "audio_fir": [w["n2"], w["d2"]], "modes": w["modes"],
            "parallelism": "receivers sharded by tuner, no collective"}


def hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


def stored_traffic(key):
    """DRAM bytes per launch of the dominant kernel from a COMMITTED ncu --set full capture (never
    measured in this run; the file says which capture)."""
    path = os.path.join(ROOT, TRAFFIC_FILE)
    if not os.path.exists(path):
        return {}
    return json.load(open(path)).get(key, {})


# ------------------------------------------------------------------ clocks, placement ----

class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "50",
                 "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        out = ""
        if self.proc:
            self.proc.terminate()
            try:
                out, _ = self.proc.communicate(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()
                out, _ = self.proc.communicate()
            self.proc = None
        if not out.strip():
            try:
                out = subprocess.check_output(
                    ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-i", str(self.gpu)],
                    text=True, stderr=subprocess.DEVNULL)
            except (OSError, subprocess.CalledProcessError):
                return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        # "under load" = samples in the upper half of the observed range (idle samples bracket the run)
        hi = [x for x in sm if x >= 0.5 * max(sm)]
        return {"sm_mhz": statistics.median(hi), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def bind_near_gpu(local):
    """Pin this rank (and the pinned buffers it is about to allocate: first touch) to the host cores
    next to its GPU: eight ranks feeding eight GPUs through one NUMA node was what held the host
    path's scaling at 0.44 in round 1.  Returns a short description for the JSON line."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(local).pci_bus_id
        dom = getattr(torch.cuda.get_device_properties(local), "pci_domain_id", 0)
        dev = getattr(torch.cuda.get_device_properties(local), "pci_device_id", 0)
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0"
        node = int(open(path + "/numa_node").read())
        cpus = open(path + "/local_cpulist").read().strip()
        if node < 0 or not cpus:
            return {"numa_node": node, "bound": False}
        ids = set()
        for part in cpus.split(","):
            a, _, b = part.partition("-")
            ids.update(range(int(a), int(b or a) + 1))
        ids &= os.sched_getaffinity(0)
        if ids:
            os.sched_setaffinity(0, ids)
        return {"numa_node": node, "cpus": cpus, "bound": bool(ids)}
    except Exception as e:  # placement is best effort
        return {"bound": False, "why": str(e)[:80]}


# ------------------------------------------------------------------ CPU reference arm ----

def _cpu_threads(n_rx):
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    replicas = max(1, cores // n_rx)
    return min(cores, n_rx * replicas), replicas


def cpu_sample_receivers(w):
    """The receivers of a workload the CPU arms run: all of them for a small bank, an evenly spaced
    subset (whole streams, so every sampled receiver keeps its own tuner stream) for a large one.
    Throughput is per receiver-frame and receivers are independent, so the figure scales linearly."""
    R, T = w["n_rx"], w["n_streams"]
    if R <= CPU_SAMPLE_RX:
        return list(range(R))
    per = R // T
    if per >= CPU_SAMPLE_RX:
        return list(range(CPU_SAMPLE_RX))            # the first tuner's receivers
    nstreams = max(1, CPU_SAMPLE_RX // per)
    step = max(1, T // nstreams)
    return [s * per + k for s in range(0, step * nstreams, step) for k in range(per)]


def cpu_reference_run(w, seconds, threads=None, flavour="ref", steps=None, warmup=1, first_stream=0):
    """The reference's CPU chain on the host cores, fed the SAME synthetic blocks as the GPU arm
    (synth.lattice_noise of the receiver's global stream).  Each worker thread owns an independent
    graph (tuner source + its share of the receivers), exactly how the reference would be scaled out
    -- its own Radio::run visits receivers sequentially on one thread (radio.cxx:56-59), which is
    what `threads=1` times (single thread next to all cores).
    Runs `steps` blocks if given, else as many as fill `seconds`; never less than MIN_CPU_SECONDS.
    Returns dict(value MS/s, seconds, steps, cores, kind, sample)."""
    import graphlib as G
    from oracle import wro
    t1, t2 = workload_taps(w, wro.lowpass_design)
    ifs, modes = synth.workload_ifs(w), synth.workload_modes(w)
    F, R, T = w["frames"], w["n_rx"], w["n_streams"]
    per = R // T
    rx_ids = cpu_sample_receivers(w)
    nthreads, replicas = _cpu_threads(len(rx_ids)) if threads is None else (threads, 1)
    total_rx = len(rx_ids) * replicas
    streams = sorted({r // per for r in rx_ids})
    iq = {s: synth.lattice_noise(F, stream=first_stream + s) for s in streams}
    use_ref = G.have(flavour)
    assign = [[] for _ in range(nthreads)]
    for i in range(total_rx):
        assign[i % nthreads].append(rx_ids[i % len(rx_ids)])

    if use_ref:
        graphs = []
        for th in range(nthreads):
            # one graph per (thread, stream) so every receiver is fed its own tuner stream
            per_stream = {}
            for r in assign[th]:
                per_stream.setdefault(r // per, []).append(r)
            gl = []
            for s, rxs in per_stream.items():
                g = G.Graph(flavour, w["fs"], F)
                for r in rxs:
                    g.add_receiver(if_hz=int(ifs[r]), ch_passband=w["pb1"], ch_rate=0, ch_decim=w["d1"],
                                   mode=int(modes[r]), au_passband=w["pb2"], au_rate=0, au_decim=w["d2"], capture=0)
                assert g.start()
                for k in range(len(rxs)):
                    g.set_taps(k, 0, t1)
                    g.set_taps(k, 1, t2)
                gl.append((g, iq[s]))
            graphs.append(gl)

        def work(th, n):
            for _ in range(n):
                for g, x in graphs[th]:
                    g.run(x)
    else:
        rxs = [[wro.Rx(w["fs"], int(ifs[r]), t1, w["d1"], int(modes[r]), t2, w["d2"]) for r in assign[th]]
               for th in range(nthreads)]

        def work(th, n):
            for _ in range(n):
                for j, rx in enumerate(rxs[th]):
                    rx.process(iq[assign[th][j] // per])

    def run_all(n):
        ths = [threading.Thread(target=work, args=(t, n)) for t in range(nthreads)]
        t0 = time.perf_counter()
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        return time.perf_counter() - t0

    run_all(max(1, warmup))
    probe = max(run_all(1), 1e-6)
    want = max(seconds, MIN_CPU_SECONDS)
    n = max(1, int(np.ceil(want / probe))) if steps is None else max(steps, int(np.ceil(1.15 * MIN_CPU_SECONDS / probe)))
    secs = run_all(n)
    while secs < MIN_CPU_SECONDS:          # a noisy probe under-sized the sample: time a longer one
        n = int(np.ceil(n * 1.3 * MIN_CPU_SECONDS / max(secs, 1e-6)))
        secs = run_all(n)
    if use_ref:
        for gl in graphs:
            for g, _ in gl:
                g.close()
    frames = total_rx * F * n
    what = f"{len(rx_ids)} of the workload's {R} receivers" if len(rx_ids) < R else f"all {R} receivers"
    return {
        "value": frames / secs / 1e6, "seconds": secs, "steps": n, "cores": nthreads,
        "kind": "reference" if use_ref else "port",
        "sample": f"{n} blocks of {F} frames x {total_rx} receivers ({what}, {replicas} replica(s)) on {nthreads} threads, "
                  f"{secs:.1f} s, synth.lattice_noise blocks (the GPU arm's), "
                  + (("oracle/_ref (unmodified reference, g++ -O0, the reference's stock flags)" if flavour == "ref_O0" else
                      "oracle/_ref (unmodified reference, g++ -O2 -ffp-contract=off)") if use_ref else "oracle port"),
        "host_cores": os.cpu_count(),
    }


def cpu_baseline_record(w, seconds, stock=True):
    r = cpu_reference_run(w, seconds)
    cpu = {"value": r["value"], "unit": "MSamples/s", "cores": r["cores"], "kind": r["kind"],
           "sample": r["sample"], "host_cores": r["host_cores"], "seconds": r["seconds"]}
    # the reference as shipped: ONE DSP thread visits every receiver (radio.cxx:56-59)
    r1 = cpu_reference_run(w, MIN_CPU_SECONDS, threads=1)
    cpu["single_thread_value"] = r1["value"]
    cpu["single_thread_sample"] = r1["sample"]
    if stock:
        import graphlib as G
        if G.have("ref_O0"):
            # the same chain built the way the reference's stock ./configure builds it (-O0, oracle/Makefile)
            r0 = cpu_reference_run(w, MIN_CPU_SECONDS, flavour="ref_O0")
            cpu["stock_O0_value"] = r0["value"]
            cpu["stock_O0_sample"] = r0["sample"]
    return cpu


def reference_arm(args, w, wname):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = args.steps if args.steps is not None else 20
    warmup = args.warmup if args.warmup is not None else 3
    r = cpu_reference_run(w, args.cpu_seconds, steps=steps, warmup=warmup)
    r1 = cpu_reference_run(w, MIN_CPU_SECONDS, threads=1)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": "MSamples/s",
        "n_gpus": args.gpus, "steps": r["steps"], "warmup": warmup,
        "ms_per_step": 1e3 * r["seconds"] / r["steps"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": bench_config(w, wname),
        "cpu_baseline": {"value": r["value"], "unit": "MSamples/s", "cores": r["cores"], "kind": r["kind"],
                         "sample": r["sample"], "host_cores": r["host_cores"], "seconds": r["seconds"],
                         "single_thread_value": r1["value"], "single_thread_sample": r1["sample"]},
        "e2e": {"value": r["value"], "unit": "MSamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------ GPU arm ----

class Ctx:
    """What the records of one run share: ranks, the device, the process group."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.args = torch, dist, args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU arm)")
        torch.cuda.set_device(self.local)
        self.placement = bind_near_gpu(self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, values):
        if self.world == 1:
            return [float(v) for v in values]
        t = self.torch.tensor(list(values), device="cuda", dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(x) for x in t]

    def close(self):
        if self.world > 1:
            self.dist.barrier()
            self.dist.destroy_process_group()


def parity_check(w, bank_audio, rx_ids, first_stream, t1, t2):
    """One output block of THIS run (block 0 of EVERY receiver of the bank, fresh state) against the
    oracle on the same synth block: bit-exact for AM/USB/LSB and -- on a glibc box -- for FM.  The oracle
    runs one receiver per host thread at a time (plain C calls; ctypes drops the interpreter lock)."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import wro
    ifs, modes = synth.workload_ifs(w), synth.workload_modes(w)
    per = w["n_rx"] // w["n_streams"]
    F = w["frames"]
    rx_ids = np.asarray(rx_ids)
    streams = sorted({int(r) // per for r in rx_ids})

    def one(r):
        r = int(r)
        want = wro.Rx(w["fs"], int(ifs[r]), t1, w["d1"], int(modes[r]), t2, w["d2"]).process(blocks[r // per])
        got = bank_audio[r]
        if np.array_equal(got.view(np.uint32), want.view(np.uint32)):
            return 0.0
        return max(float(np.max(np.abs(got - want))), 1e-45)

    with ThreadPoolExecutor(max(1, os.cpu_count() or 1)) as pool:
        blocks = dict(zip(streams, pool.map(lambda s: synth.lattice_noise(F, stream=first_stream + s), streams)))
        diffs = list(pool.map(one, rx_ids))
    bad = sum(1 for d in diffs if d > 0.0)
    worst = max(diffs) if diffs else 0.0
    if bad and (worst > 3e-7 or not np.any(modes[rx_ids] == synth.FM)):
        raise SystemExit(f"bench.py: parity check FAILED: {bad} of {len(rx_ids)} receivers differ from the oracle "
                         f"(max abs {worst:g}); no number is reported for a wrong result")
    return {"receivers_checked": len(rx_ids), "receivers_in_bank": int(w["n_rx"]), "block": 0,
            "oracle": "oracle/libwr_oracle.so (C port pinned to oracle/_ref)",
            "bit_exact": bad == 0, "receivers_differing": bad, "max_abs_diff": worst,
            "inputs": "identical: synth.lattice_noise (numpy) == synth.lattice_u8_torch (device), compared in this run"}


def chain_record(ctx, wname, w, steps, warmup, full_cpu=True, plugin=False):
    """Device-timed throughput, per-kernel roofline, host-path figures, in-run parity and the CPU
    baseline for one receiver-chain workload on this rank's GPU."""
    torch = ctx.torch
    from webradio_b200 import capi, shard
    args = ctx.args
    F, R, T = w["frames"], w["n_rx"], w["n_streams"]
    M1 = F // w["d1"]
    M2 = max(M1 // w["d2"], 1)
    per = R // T
    mine = shard.weak_scaling_shard(T, R, ctx.rank, ctx.world)     # this rank's own tuners (global ids)
    first_stream = mine.tuners[0]

    bank = capi.Bank(T, R, F, w["n1"], w["d1"], w["n2"], w["d2"], device=ctx.local)
    bank.set_variant(args.variant)
    t1, t2 = workload_taps(w)
    ifs, modes = synth.workload_ifs(w), synth.workload_modes(w)
    for r in range(R):
        bank.set_taps(r, 0, t1)
        bank.set_taps(r, 1, t2)
        bank.set_if(r, int(ifs[r]), w["fs"])
        bank.set_mode(r, int(modes[r]))
        bank.set_stream(r, r // per)

    # Synthetic tuner blocks on the RTL-SDR lattice, generated on the device by the same counter hash
    # as synth.lattice_noise (the CPU arm's generator); consecutive blocks of each stream, a rotating
    # set larger than L2 whatever --steps says, so that no step finds its input cached.
    u8 = args.input == "u8"
    fb = 2 if u8 else 8
    block_bytes = fb * F * T
    nbuf = max(2, -(-int(1.25 * L2_BYTES) // block_bytes))
    raw = [synth.lattice_u8_torch(F, mine.tuners, start=b * F, device="cuda") for b in range(nbuf)]
    inputs = raw if u8 else [synth.u8_to_f32_torch(x) for x in raw]
    # ... and the two generators agree: first block, first and last sampled stream, against numpy
    rx_ids = cpu_sample_receivers(w)
    for s in sorted({rx_ids[0] // per, rx_ids[-1] // per}):
        want = synth.lattice_noise(F, stream=first_stream + s)
        got = synth.u8_to_f32_torch(raw[0][s]).cpu().numpy().ravel()
        if not np.array_equal(got, want):
            raise SystemExit("bench.py: device and host input generators disagree")
    raw_keep = raw[:min(nbuf, 2)]      # the host path's blocks (both formats are derived from the bytes)
    audio = [torch.zeros(R, M2, device="cuda") for _ in range(min(nbuf, 4))]
    in_ptrs = [x.data_ptr() for x in inputs]
    out_ptrs = [y.data_ptr() for y in audio]
    stream = bank.stream()

    def run_steps(first, n):
        # n calls of wr_bank_process_device, looped on the C side of the ABI
        bank.run_device_steps(in_ptrs, F, F, out_ptrs, M2, first, n, u8=u8)

    # ---- block 0 from fresh state: the in-run parity check ----
    run_steps(0, 1)
    bank.sync()
    parity = parity_check(w, audio[0].cpu().numpy(), np.arange(R), first_stream, t1, t2) if ctx.rank == 0 else None

    clocks = ClockSampler(ctx.local)
    run_steps(1, warmup)
    bank.sync()

    # ---- timed region: exactly K steps, CUDA events on the launch stream ----
    ev0 = torch.cuda.Event(enable_timing=True)
    ev1 = torch.cuda.Event(enable_timing=True)
    ext = torch.cuda.ExternalStream(stream)
    ctx.barrier()
    clocks.start()
    launches0 = bank.launch_count()
    ev0.record(ext)
    run_steps(1 + warmup, steps)
    ev1.record(ext)
    ev1.synchronize()
    launches = bank.launch_count() - launches0
    ctx.barrier()
    # the job finishes when its slowest rank does: MAX over ranks of the device time
    ms = ctx.max_over_ranks([ev0.elapsed_time(ev1)])[0]
    value = shard.job_throughput(R * F * steps, ctx.world, ms) / 1e6
    # ---- per-kernel device time (CUDA events around each kernel, same inputs, K more steps), straight behind
    # the timed region so that it sees the same clocks (on a power-capped box the clocks sag under the
    # sustained load below: a pass taken after it once read a kernel LONGER than the step that contains it).
    # Three passes; the line reports the fastest pass's mean and lists all three.
    ksteps = min(max(steps, 20), 2000)
    kpasses = []
    for _ in range(3):
        bank.set_timing(True)
        run_steps(0, ksteps)
        chan_ms, audio_ms, nb = bank.kernel_times()
        bank.set_timing(False)
        kpasses.append((chan_ms / max(nb, 1), audio_ms / max(nb, 1)))
    chan_ms_avg, audio_ms_avg = min(kpasses)
    variant_used = bank.variant_in_use()
    # keep the load on long enough for the sampler to see it (a 20-step region of a small bank is 0.5 ms)
    t_end = time.perf_counter() + 0.4
    while time.perf_counter() < t_end:
        run_steps(0, max(steps, 50))
        bank.sync()
    clk = clocks.stop()

    # ---- host path: the C ABI with HOST buffers, H2D + kernels + D2H inside the timed region ----
    e2e = None
    if not args.no_e2e:
        e2e = host_path(ctx, bank, w, raw_keep, M2, steps)
    del inputs, raw, audio
    bank.close()
    torch.cuda.empty_cache()

    # ---- the DspBlock plug-in surface: the reference's Radio::run over the drop-in blocks ----
    if plugin and e2e is not None and T == 1:
        pl = plugin_leg(ctx, w, first_stream, t1, t2)
        if pl:
            e2e.update(pl)

    # ---- roofline of the dominant kernel ----
    peak, peak_src = hbm_peak()
    alg = algorithmic_bytes(w, fb)
    own = chan_kernel_bytes(w, variant_used, fb)
    achieved = alg / (chan_ms_avg * 1e-3) / 1e9 if chan_ms_avg > 0 else 0.0
    cap = stored_traffic(wname + ("_u8" if u8 else ""))
    names = {1: "chan_kernel_v1: fused NCO mix + channel FIR + demod", 2: "chan_kernel_v2: fused NCO mix + channel FIR",
             3: "chan_kernel_v3: fused NCO mix + channel FIR", 4: "chan_kernel_v4: fused NCO mix + streaming channel FIR"}
    roofline = {
        "bound": "hbm", "kernel": names.get(variant_used, str(variant_used)),
        "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "traffic": cap.get("chan_kernel_dram_bytes_per_launch"),
        "traffic_source": ("stored ncu --set full capture, not measured in this run: " + cap.get("source", TRAFFIC_FILE)) if cap else None,
        "peak_source": peak_src,
        "frac_of_nominal_8000_gbs": achieved / 8000.0,          # SURVEY.md 8d also asks for this one
        "algorithmic_bytes_per_launch": alg,
        "algorithmic_bytes_formula": "SURVEY.md 8d: 8*F*T + 4*R*F/(D1*D2)" if fb == 8 else "SURVEY.md 8d with 2-byte frames: 2*F*T + 4*R*F/(D1*D2)",
        "kernel_own_bytes_per_launch": own, "kernel_own_gbs": own / (chan_ms_avg * 1e-3) / 1e9 if chan_ms_avg > 0 else 0.0,
        "kernel_ms": chan_ms_avg, "audio_kernel_ms": audio_ms_avg,
        "kernel_ms_passes": [round(c, 6) for c, _ in kpasses],
        # serialised kernel times (CUDA events around each kernel); in the timed region the next block's
        # channel kernel starts under this block's demodulator kernel, so ms_per_step < their sum
        "kernel_share_of_step": chan_ms_avg / max(chan_ms_avg + audio_ms_avg, 1e-12),
        "whole_step_frac": alg * steps / (ms * 1e-3) / 1e9 / peak,
        "ncu_stored": {k: cap[k] for k in ("issue_active_pct", "l1tex_throughput_pct", "fma_pipe_cycles_active_pct",
                                            "warp_instructions_per_launch", "kernel", "source") if k in cap} or None,
        "receiver_frames_per_s": R * F / (chan_ms_avg * 1e-3) if chan_ms_avg > 0 else 0.0,
        "note": ("shared-tuner workload: every receiver re-uses the one tuner block from L2/shared memory, "
                 "so DRAM traffic is small by construction and the kernel is FP32-issue bound, not HBM bound "
                 "(SURVEY.md 7)") if T < R else "independent streams: each input byte is touched once",
    }

    cpu = None
    if ctx.rank == 0 and ctx.world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline_record(w, args.cpu_seconds, stock=full_cpu)

    rec = {
        "value": value, "unit": "MSamples/s", "steps": steps, "warmup": warmup, "ms_per_step": ms / steps,
        "config": bench_config(w, wname),
        "input": "interleaved float IQ" if not u8 else "raw RTL-SDR bytes (u8 IQ, converted in the channel kernel's load)",
        "l2": f"rotating set of {nbuf} distinct input blocks ({nbuf * block_bytes / 2**20:.0f} MiB > L2 126 MiB)",
        "clocks": clk, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu,
        "parity": parity,
        "tuner_msamples_per_s": ctx.world * T * F * steps / (ms * 1e-3) / 1e6,
        "kernel_variant": variant_used,
    }
    return rec


def host_path(ctx, bank, w, raw_blocks, M2, steps):
    """e2e.value: the synchronous wr_bank_process call on HOST buffers -- what DspBlock::process of the
    drop-in blocks makes once per tuner block (copy in, kernels, copy out inside the call).  Beside it:
    the same fed raw RTL-SDR bytes, and the pipelined submit/wait form of the C ABI (several blocks in
    flight; a host that owns its tuner loop can use it, the unchanged Radio::run cannot)."""
    torch = ctx.torch
    F, R, T = w["frames"], w["n_rx"], w["n_streams"]
    depth = bank.pipeline_depth()
    small = T == 1
    pin_u8 = [x.cpu().pin_memory() for x in raw_blocks]
    pin_f32 = [synth.u8_to_f32_torch(x).pin_memory() for x in pin_u8]
    pin_out = [torch.zeros(R, M2).pin_memory() for _ in range(depth + 1)]
    outp = [y.data_ptr() for y in pin_out]
    # host-timed: regions long enough that the host's timer and the link's wake-up do not show
    esteps = max(steps, 500) if small else max(min(steps, 10), 4)
    reps = 5 if small else 3

    def run(ptrs, n, pipelined, u8):
        bank.run_host_steps(ptrs, F, outp, M2, 0, n, pipelined=pipelined, u8=u8)

    out = {}
    for name, bufs, u8 in (("f32", pin_f32, False), ("u8", pin_u8, True)):
        ptrs = [x.data_ptr() for x in bufs]
        # warm-up: every pipeline slot used once (its HBM buffers are allocated on first use), the
        # link and the host clocks ramped (~0.2 s of traffic on small blocks)
        for _ in range(3 if small else 1):
            run(ptrs, max(esteps, 1000) if small else depth + 2, True, u8)
        res = {}
        for mode, pipelined, n in (("sync", False, max(esteps // 2, 4) if small else esteps), ("pipelined", True, esteps)):
            runs = []
            for _ in range(reps):
                ctx.barrier()
                t0 = time.perf_counter()
                run(ptrs, n, pipelined, u8)
                torch.cuda.synchronize()
                runs.append(time.perf_counter() - t0)
            runs = ctx.max_over_ranks(runs)
            res[mode] = ctx.world * R * F * n / statistics.median(runs) / 1e6
            res[mode + "_all"] = [ctx.world * R * F * n / x / 1e6 for x in runs]
            res[mode + "_steps"] = n
        out[name] = res
    fbytes = 8 * F * T
    return {
        "value": out["f32"]["sync"], "unit": "MSamples/s",
        "h2d_bytes_per_step": fbytes, "d2h_bytes_per_step": 4 * R * M2,
        "mode": "synchronous wr_bank_process on pinned host buffers, float IQ: the call DspBlock::process makes "
                "(one block per call; the copy in, the kernels and the copy out overlap inside it by sub-block)",
        "steps": out["f32"]["sync_steps"], "repeats": reps, "values": out["f32"]["sync_all"],
        "pipelined_value": out["f32"]["pipelined"], "pipelined_depth": depth,
        "pipelined_note": "wr_bank_submit / wr_bank_wait, several blocks in flight: not reachable from the unchanged Radio::run",
        "u8_value": out["u8"]["sync"], "u8_pipelined_value": out["u8"]["pipelined"],
        "u8_h2d_bytes_per_step": 2 * F * T,
        "u8_note": "the same calls fed raw RTL-SDR bytes, the reference tuner's native format (rtlsdrtuner.cxx:104-108)",
    }


def plugin_leg(ctx, w, first_stream, t1, t2):
    """The reference's UNMODIFIED src/radio.cxx (Radio::run, radio.cxx:56-59) over the drop-in
    DspBlock classes, SpectrumSink attached as FrontEnd does it (radio.cxx:126-128), pageable
    std::vector buffers: tests/harness/libwr_radio_dropin.so.  One front-end; every receiver's audio
    block of the first step is compared with the oracle before anything is timed."""
    so = os.path.join(ROOT, "tests", "harness", "libwr_radio_dropin.so")
    if not os.path.exists(so) or ctx.world > 1:
        return None
    from oracle import wro
    C = ctypes
    fp = C.POINTER(C.c_float)
    L = C.CDLL(so, mode=C.RTLD_LOCAL)
    L.wrr_create.restype = C.c_void_p
    L.wrr_create.argtypes = [C.c_uint, C.c_uint, C.c_uint]
    L.wrr_add_receiver_geo.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_uint, C.c_uint, C.c_uint, C.c_uint]
    L.wrr_set_taps.argtypes = [C.c_void_p, C.c_int, C.c_int, fp, C.c_uint]
    L.wrr_start.argtypes = [C.c_void_p]
    L.wrr_run_steps.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.c_uint, C.c_uint, C.c_uint]
    L.wrr_audio.restype = C.c_long
    L.wrr_audio.argtypes = [C.c_void_p, C.c_int, fp, C.c_long]
    L.wrr_destroy.argtypes = [C.c_void_p]
    L.wrr_ring_fill.argtypes = [C.c_void_p, fp]
    L.wrr_run_ring.argtypes = [C.c_void_p, C.c_uint]
    L.wrr_capture.argtypes = [C.c_int]
    F, R = w["frames"], w["n_rx"]
    ifs, modes = synth.workload_ifs(w), synth.workload_modes(w)
    rig = L.wrr_create(w["fs"], F, 512)
    try:
        for r in range(R):
            if L.wrr_add_receiver_geo(rig, int(ifs[r]), synth.MODE_NAMES[int(modes[r])].encode(),
                                      w["n1"], w["d1"], w["n2"], w["d2"]) < 0:
                return None
        if L.wrr_start(rig) != 0:
            return {"plugin_value": None, "plugin_note": "the drop-in pipeline did not start"}
        a1, a2 = np.ascontiguousarray(t1, np.float32), np.ascontiguousarray(t2, np.float32)
        for r in range(R):
            L.wrr_set_taps(rig, r, 0, a1.ctypes.data_as(fp), a1.size)
            L.wrr_set_taps(rig, r, 1, a2.ctypes.data_as(fp), a2.size)
        nblk = 8
        blocks = [synth.lattice_noise(F, stream=first_stream, start=b * F) for b in range(nblk)]
        ptrs = (C.c_void_p * nblk)(*[b.ctypes.data for b in blocks])
        L.wrr_run_steps(rig, ptrs, nblk, 0, 1)
        M2 = F // w["d1"] // w["d2"]
        bad = 0
        for r in range(R):
            got = np.empty(M2, np.float32)
            n = L.wrr_audio(rig, r, got.ctypes.data_as(fp), M2)
            want = wro.Rx(w["fs"], int(ifs[r]), t1, w["d1"], int(modes[r]), t2, w["d2"]).process(blocks[0])
            if n != M2 or not np.array_equal(got.view(np.uint32), want.view(np.uint32)):
                if np.max(np.abs(got - want)) > 3e-7:
                    raise SystemExit("bench.py: plug-in leg: audio differs from the oracle")
                bad += 1
        # timed as the reference's main loop runs (main.cxx:114-115): the tuner swaps ring buffers into
        # its output as RtlSdrTuner does (rtlsdrtuner.cxx:265-285, no copy), nobody listens to the audio
        # streams (AudioStreamManager returns at once, audiostream.cxx:65-73)
        for blk in blocks[:4]:       # N_BUFFERS of the reference's tuner (rtlsdrtuner.h)
            L.wrr_ring_fill(rig, blk.ctypes.data_as(fp))
        L.wrr_capture(0)
        L.wrr_run_ring(rig, 300)
        runs = []
        n = 300
        for _ in range(5):
            t0 = time.perf_counter()
            L.wrr_run_ring(rig, n)
            runs.append(time.perf_counter() - t0)
        L.wrr_capture(1)
        return {"plugin_value": R * F * n / statistics.median(runs) / 1e6,
                "plugin_values": [R * F * n / x / 1e6 for x in runs], "plugin_steps": n,
                "plugin_parity": {"receivers_checked": R, "bit_exact": bad == 0},
                "plugin_note": "Radio::run() of the reference's unmodified src/radio.cxx over the drop-in blocks: one "
                               "FrontEnd (a tuner that swaps ring buffers into its output as RtlSdrTuner does + "
                               "SpectrumSink, 512 points) and every receiver, std::vector buffers, synchronous "
                               "DspBlock::process per block"}
    finally:
        L.wrr_destroy(rig)


def gpu_arm(args, wname):
    ctx = Ctx(args)
    w = synth.WORKLOADS[wname]
    T = w["n_streams"]
    steps = args.steps if args.steps is not None else (2000 if T == 1 else 30)
    warmup = max(args.warmup if args.warmup is not None else 5, 3)
    subs = args.subs
    if subs is None:
        subs = "cfg2,cfg4,cfg5" if wname == "cfg3" else "none"
    subs = [s for s in subs.split(",") if s and s != "none" and s != wname]

    rec = chain_record(ctx, wname, w, steps, warmup, full_cpu=True, plugin=True)
    configs = {}
    for s in subs:
        if s == "cfg4":
            import bench_spectrum
            configs[s] = bench_spectrum.record(ctx, steps=None, warmup=3)
        else:
            ws = synth.WORKLOADS[s]
            ssteps = 400 if ws["n_streams"] == 1 else 20
            configs[s] = chain_record(ctx, s, ws, ssteps, 5, full_cpu=False, plugin=True)
            if s == "cfg5":
                configs[s]["share"] = (f"one GPU's share of BASELINE configs[4] (8192 receivers on 8 GPUs = 16 tuners x 64 "
                                       f"receivers per GPU); this run holds {ctx.world} such share(s) on {ctx.world} GPU(s)")
    if ctx.rank == 0:
        line = {
            "metric": METRIC, "value": rec["value"], "unit": "MSamples/s", "n_gpus": ctx.world, "steps": rec["steps"],
            "warmup": rec["warmup"], "ms_per_step": rec["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        }
        line.update({k: rec[k] for k in rec if k not in line})
        line["placement"] = ctx.placement
        if configs:
            line["configs"] = configs
        print(json.dumps(line), flush=True)
    ctx.close()


def claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries print there too (NCCL's version banner
    under torchrun does), so file descriptor 1 is pointed at stderr for the life of the process and
    the JSON line goes to the saved, real stdout -- after which fd 1 is restored."""
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    return real


def emit(real_fd, text):
    sys.stdout.flush()
    os.dup2(real_fd, 1)
    os.close(real_fd)
    sys.stdout.write(text)
    sys.stdout.flush()


class _Capture:
    """stdout of the arms while fd 1 is diverted: keeps what they print (the JSON line)."""

    def __init__(self):
        self.parts = []

    def write(self, x):
        self.parts.append(x)

    def flush(self):
        pass

    def isatty(self):
        return False

    def writable(self):
        return True

    def fileno(self):
        return 1   # diverted to stderr while the arms run

    encoding = "utf-8"


def main():
    args = parse_args()
    real, cap, py_stdout = claim_stdout(), _Capture(), sys.stdout
    sys.stdout = cap
    try:
        run(args)
    finally:
        sys.stdout = py_stdout
        emit(real, "".join(cap.parts))


def run(args):
    if args.workload == "cfg4":
        import bench_spectrum
        return bench_spectrum.main(args)
    if args.impl == "reference":
        reference_arm(args, synth.WORKLOADS[args.workload], args.workload)
    else:
        gpu_arm(args, args.workload)


if __name__ == "__main__":
    main()
