#!/usr/bin/env python
"""bench.py -- throughput of the per-receiver DSP hot path (NCO mix -> decimating FIR -> demod ->
audio FIR) on B200, next to the reference's own CPU path.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2] [--impl reference]

One "step" = one tuner block of the workload through every receiver of this rank's bank.
Prints ONE JSON line (rank 0).  Keys beyond the base contract:
  roofline     dominant kernel's algorithmic bytes / its CUDA-event time vs the measured HBM peak
  cpu_baseline the reference's CPU chain (oracle/_ref, else the C port) timed on the host cores
  e2e          same metric through the C ABI with HOST buffers (H2D + kernels + D2H every step)
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


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=None)
    p.add_argument("--warmup", type=int, default=None)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--workload", default="cfg2", choices=sorted(synth.WORKLOADS) + ["cfg4"])
    p.add_argument("--variant", type=int, default=0, help="kernel family: 0 auto, 1 v1, 2 v2, 3 v3")
    p.add_argument("--input", default="f32", choices=["f32", "u8"],
                   help="tuner block format: interleaved float IQ (the DspBlock convention) or raw RTL-SDR bytes "
                        "converted inside the channel kernel's load (SURVEY.md 8f-1)")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--cpu-seconds", type=float, default=12.0, help="target wall time of the CPU baseline sample")
    return p.parse_args()


# ------------------------------------------------------------------ workload ----

def workload_taps(w):
    """Tap values are an input of the FIR (the reference's own design collapses to all-zero for the
    narrow pass-bands of cfg2/3/5: maxbin = N*passband/Fs/2 = 0, lowpass.cxx:167).  cfg1 uses the
    reference design; the others a Hamming windowed-sinc of the same length."""
    from webradio_b200 import capi
    if w["n1"] == 64:
        t1 = capi.lowpass_design(64, w["pb1"], w["fs"])
    else:
        t1 = synth.windowed_sinc(w["n1"], w["pb1"] / w["fs"])
    t2 = capi.lowpass_design(w["n2"], w["pb2"], w["fs"] // w["d1"])
    if not np.any(t2):
        t2 = synth.windowed_sinc(w["n2"], w["pb2"] / (w["fs"] // w["d1"]))
    return t1, t2


def algorithmic_bytes(w, frame_bytes=8):
    """B_alg = 8*F*T + 4*R*F/(D1*D2) per block (T = unique input streams); 2 bytes per
    frame instead of 8 when the tuner block arrives as raw RTL-SDR bytes."""
    F, T, R = w["frames"], w["n_streams"], w["n_rx"]
    return frame_bytes * F * T + 4 * R * (F // w["d1"] // w["d2"])


def chan_kernel_bytes(w, variant, frame_bytes=8):
    """The dominant kernel alone: reads the tuner block(s) once and writes, per channel-rate
    sample per receiver, the demodulated float (v1: demod fused in) or the IQ pair (v2, v3)."""
    F, T, R = w["frames"], w["n_streams"], w["n_rx"]
    return frame_bytes * F * T + (8 if variant >= 2 else 4) * R * (F // w["d1"])


# ------------------------------------------------------------------ clocks ----

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
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100",
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


# ------------------------------------------------------------------ CPU reference arm ----

def _cpu_threads(n_rx):
    cores = os.cpu_count() or 1
    replicas = max(1, cores // n_rx)
    return min(cores, n_rx * replicas), replicas


def cpu_reference_run(w, steps, warmup, max_seconds=None, threads=None, flavour="ref"):
    """The reference's CPU chain on the host cores.  Each worker thread owns an independent graph
    (tuner source + its share of the receivers), exactly how the reference would be scaled out --
    its own Radio::run visits receivers sequentially on one thread (radio.cxx:56-59), which is
    what `threads=1` times (single thread next to all cores).
    Returns dict(value MS/s, seconds, steps, cores, kind, sample)."""
    import graphlib as G
    t1, t2 = workload_taps(w)
    ifs, modes = synth.workload_ifs(w), synth.workload_modes(w)
    F, R, T = w["frames"], w["n_rx"], w["n_streams"]
    nthreads, replicas = _cpu_threads(R) if threads is None else (threads, 1)
    total_rx = R * replicas
    iq = [synth.lattice_noise(F, stream=t) for t in range(min(T, 8))]
    use_ref = G.have(flavour)
    assign = [[] for _ in range(nthreads)]
    for i in range(total_rx):
        assign[i % nthreads].append(i % R)

    if use_ref:
        graphs = []
        for th in range(nthreads):
            # one graph per (thread, stream) so every receiver is fed its own tuner stream
            per_stream = {}
            for r in assign[th]:
                per_stream.setdefault(r % T % len(iq), []).append(r)
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
        from oracle import wro
        rxs = [[wro.Rx(w["fs"], int(ifs[r]), t1, w["d1"], int(modes[r]), t2, w["d2"]) for r in assign[th]]
               for th in range(nthreads)]

        def work(th, n):
            for _ in range(n):
                for j, rx in enumerate(rxs[th]):
                    rx.process(iq[assign[th][j] % T % len(iq)])

    def run_all(n):
        ths = [threading.Thread(target=work, args=(t, n)) for t in range(nthreads)]
        t0 = time.perf_counter()
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        return time.perf_counter() - t0

    run_all(max(1, warmup))
    if max_seconds is not None:
        probe = run_all(1)
        steps = max(1, min(steps, int(max_seconds / max(probe, 1e-6))))
    secs = run_all(steps)
    if use_ref:
        for gl in graphs:
            for g, _ in gl:
                g.close()
    frames = total_rx * F * steps
    return {
        "value": frames / secs / 1e6, "seconds": secs, "steps": steps, "cores": nthreads,
        "kind": "reference" if use_ref else "port",
        "sample": f"{steps} blocks of {F} frames x {total_rx} receivers "
                  f"({replicas} replica(s) of the workload) on {nthreads} threads, "
                  + (("oracle/_ref (unmodified reference, g++ -O0, the reference's stock flags)" if flavour == "ref_O0" else
                      "oracle/_ref (unmodified reference, g++ -O2 -ffp-contract=off)") if use_ref else "oracle port"),
        "host_cores": os.cpu_count(),
    }


def stock_flags_run(w, seconds):
    """The same chain built the way the reference's stock ./configure builds it (-O0, see
    oracle/Makefile), all host threads, a short sample: SURVEY.md 8d asks for it once.  The
    headline CPU figure stays the -O2 build, which is the faster one."""
    import graphlib as G
    if not G.have("ref_O0"):
        return {}
    r = cpu_reference_run(w, 1000, 1, max_seconds=seconds, flavour="ref_O0")
    return {"stock_O0_value": r["value"], "stock_O0_sample": r["sample"]}


def reference_arm(args, w, wname):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = args.steps if args.steps is not None else 20
    warmup = args.warmup if args.warmup is not None else 3
    r = cpu_reference_run(w, steps, warmup, max_seconds=120.0)
    r1 = cpu_reference_run(w, 3, 1, max_seconds=10.0, threads=1)
    stock = stock_flags_run(w, 5.0)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": "MSamples/s",
        "n_gpus": args.gpus, "steps": r["steps"], "warmup": warmup,
        "ms_per_step": 1e3 * r["seconds"] / r["steps"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": bench_config(w, wname, None),
        "cpu_baseline": {"value": r["value"], "unit": "MSamples/s", "cores": r["cores"], "kind": r["kind"],
                         "sample": r["sample"], "host_cores": r["host_cores"],
                         "single_thread_value": r1["value"], "single_thread_sample": r1["sample"], **stock},
        "e2e": {"value": r["value"], "unit": "MSamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def bench_config(w, wname, l2_note, input_format="f32"):
    c = {"workload": f"{wname}: {w['desc']}", "input": "interleaved float IQ" if input_format == "f32" else
         "raw RTL-SDR bytes (u8 IQ, converted in the channel kernel's load)",
         "sample_rate": w["fs"], "frames_per_step": w["frames"],
         "n_receivers": w["n_rx"], "n_streams": w["n_streams"], "channel_fir": [w["n1"], w["d1"]],
         "audio_fir": [w["n2"], w["d2"]], "modes": w["modes"], "parallelism": "receivers sharded by tuner, no collective"}
    if l2_note:
        c["l2"] = l2_note
    return c


# ------------------------------------------------------------------ GPU arm ----

def gpu_arm(args, w, wname):
    import torch
    import torch.distributed as dist

    from webradio_b200 import capi, shard

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    F, R, T = w["frames"], w["n_rx"], w["n_streams"]
    M1 = F // w["d1"]
    M2 = M1 // w["d2"]
    steps = args.steps if args.steps is not None else (2000 if T == 1 else 30)
    warmup = args.warmup if args.warmup is not None else 5
    warmup = max(warmup, 3)

    bank = capi.Bank(T, R, F, w["n1"], w["d1"], w["n2"], w["d2"], device=local)
    bank.set_variant(args.variant)
    t1, t2 = workload_taps(w)
    ifs, modes = synth.workload_ifs(w), synth.workload_modes(w)
    for r in range(R):
        bank.set_taps(r, 0, t1)
        bank.set_taps(r, 1, t2)
        bank.set_if(r, int(ifs[r]), w["fs"])
        bank.set_mode(r, int(modes[r]))
        bank.set_stream(r, r % T)

    # synthetic tuner blocks on the RTL-SDR sample lattice (b-128)/128, generated in HBM; a
    # rotating set larger than L2 so that no step finds its input cached
    u8 = args.input == "u8"
    frame_bytes = 2 if u8 else 8
    block_bytes = frame_bytes * F * T
    nbuf = max(2, -(-int(1.25 * L2_BYTES) // block_bytes))
    nbuf = min(nbuf, max(2, steps + warmup))
    # weak scaling: this rank owns its own copy of the workload's tuners and receivers
    mine = shard.weak_scaling_shard(T, R, rank, world)
    gen = torch.Generator(device="cuda")
    gen.manual_seed(0xB200 + mine.tuners[0])
    inputs = []
    for _ in range(nbuf):
        raw = torch.randint(0, 256, (T, F, 2), generator=gen, device="cuda", dtype=torch.int32)
        inputs.append(raw.to(torch.uint8).contiguous() if u8 else ((raw.float() - 128.0) / 128.0).contiguous())
        del raw
    audio = [torch.zeros(R, max(M2, 1), device="cuda") for _ in range(min(nbuf, 8))]
    l2_note = f"rotating set of {nbuf} distinct input blocks ({nbuf * block_bytes / 2**20:.0f} MiB > L2 126 MiB)" \
        if nbuf * block_bytes > L2_BYTES else \
        f"rotating set of {nbuf} input blocks ({nbuf * block_bytes / 2**20:.0f} MiB); run is shorter than the L2-sized set"
    stream = bank.stream()

    in_ptrs = [x.data_ptr() for x in inputs]
    out_ptrs = [y.data_ptr() for y in audio]

    def run_steps(first, n):
        # n calls of wr_bank_process_device, looped on the C side of the ABI
        bank.run_device_steps(in_ptrs, F, F, out_ptrs, max(M2, 1), first, n, u8=u8)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    clocks = ClockSampler(local)
    run_steps(0, warmup)
    bank.sync()

    # ---- timed region: exactly K steps, CUDA events on the launch stream ----
    ev0 = torch.cuda.Event(enable_timing=True)
    ev1 = torch.cuda.Event(enable_timing=True)
    ext = torch.cuda.ExternalStream(stream)
    barrier()
    clocks.start()
    launches0 = bank.launch_count()
    ev0.record(ext)
    run_steps(warmup, steps)
    ev1.record(ext)
    ev1.synchronize()
    launches = bank.launch_count() - launches0
    barrier()
    # the job finishes when its slowest rank does: MAX over ranks of the device time
    ms = shard.reduce_max_ms(ev0.elapsed_time(ev1), device="cuda")
    value = shard.job_throughput(R * F * steps, world, ms) / 1e6

    # ---- per-kernel device time (CUDA events around each kernel, same inputs, K more steps) ----
    bank.set_timing(True)
    ksteps = min(steps, 2000)
    run_steps(warmup + steps, ksteps)
    chan_ms, audio_ms, nb = bank.kernel_times()
    bank.set_timing(False)
    chan_ms_avg = chan_ms / max(nb, 1)
    audio_ms_avg = audio_ms / max(nb, 1)

    # ---- e2e: C ABI with HOST (pinned) buffers, H2D + kernels + D2H every step, pipelined ----
    depth = bank.pipeline_depth()
    nh = min(nbuf, 4 if T > 1 else 16)
    pin_in = [inputs[i].cpu().pin_memory() for i in range(nh)]
    pin_out = [torch.zeros(R, max(M2, 1)).pin_memory() for _ in range(depth + 1)]
    # e2e is host-timed: a region of a few blocks (a driver may ask for --steps 10) would measure the
    # host's timer and the link's wake-up, not the path.  Its regions are therefore at least 1000
    # blocks long for single-tuner workloads (20-30 ms) and at least 5 for the multi-stream ones
    # (a block there is milliseconds); `e2e.steps` says what was used.
    esteps = max(steps, 1000) if T == 1 else max(min(steps, 10), 5)

    pin_in_ptrs = [x.data_ptr() for x in pin_in]
    pin_out_ptrs = [y.data_ptr() for y in pin_out]

    def e2e_pipelined(n):
        bank.run_host_steps(pin_in_ptrs, F, pin_out_ptrs, max(M2, 1), 0, n, pipelined=True, u8=u8)

    def e2e_sync(n):
        bank.run_host_steps(pin_in_ptrs, F, pin_out_ptrs, max(M2, 1), 0, n, pipelined=False, u8=u8)

    # The clock sampler (an nvidia-smi loop) covered the device-timed region above; it is stopped
    # here because its driver queries stall the submitting thread of a host-driven loop.
    clk = clocks.stop()
    # Host-timed, so a hiccup of the host shows: `reps` timed regions of `esteps` steps each, the
    # MEDIAN is reported and all of them are listed.
    reps = 5 if T == 1 else 3
    # warm-up: the host path needs ~0.2 s of traffic before it is steady (measured: 144 k, 149 k,
    # 180 k, 222 k, 224 k MS/s over the first five regions of 2000 steps after a 3-step warm-up;
    # PCIe link and host clocks ramp)
    # ... and every slot of the pipeline must have been used once: a slot's HBM buffers are
    # allocated on first use (cudaMalloc synchronises), so fewer warm-up blocks than the depth put
    # allocations into the timed region (r01k: cfg3 fed bytes 6.1 k instead of ~25 k MS/s)
    for _ in range(4 if T == 1 else 1):
        e2e_pipelined(max(esteps, 2000) if T == 1 else depth + 2)
    e2e_runs, sync_runs = [], []
    ssteps = min(esteps, 200)
    for _ in range(reps):
        barrier()
        t0 = time.perf_counter()
        e2e_pipelined(esteps)
        torch.cuda.synchronize()
        e2e_runs.append(time.perf_counter() - t0)
    for _ in range(reps):
        barrier()
        t0 = time.perf_counter()
        e2e_sync(ssteps)
        torch.cuda.synchronize()
        sync_runs.append(time.perf_counter() - t0)
    if world > 1:
        tt = torch.tensor(e2e_runs + sync_runs, device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_runs, sync_runs = [float(x) for x in tt[:reps]], [float(x) for x in tt[reps:]]
    e2e_s = statistics.median(e2e_runs)
    e2e_sync_s = statistics.median(sync_runs)
    e2e_value = world * R * F * esteps / e2e_s / 1e6
    e2e_sync_value = world * R * F * ssteps / e2e_sync_s / 1e6
    e2e_all = [world * R * F * esteps / x / 1e6 for x in e2e_runs]

    # ---- roofline of the dominant kernel (fused mix + FIR + demod) ----
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak = float(json.load(open(peaks_path))["hbm_gbs"])
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"
    variant_used = bank.variant_in_use()
    kb = chan_kernel_bytes(w, variant_used, frame_bytes)
    achieved = kb / (chan_ms_avg * 1e-3) / 1e9 if chan_ms_avg > 0 else 0.0
    traffic, ncu = None, None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        rec = json.load(open(tpath)).get(wname + ("_u8" if u8 else ""), {})
        traffic = rec.get("chan_kernel_dram_bytes_per_launch")
        if rec:
            # what binds the kernel when it is not HBM (one ncu --set full capture, profiles/)
            ncu = {k: rec[k] for k in ("issue_active_pct", "l1tex_throughput_pct", "fma_pipe_cycles_active_pct",
                                       "warp_instructions_per_launch", "kernel", "source") if k in rec}
    roofline = {
        "bound": "hbm",
        "kernel": f"chan_kernel_v{variant_used}: fused NCO mix + channel FIR" if variant_used >= 2 else "chan_kernel_v1: fused NCO mix + channel FIR + demod",
        "achieved": achieved, "peak": peak,
        "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
        # SURVEY.md 8d also asks for the fraction of the chip's nominal ~8 TB/s
        "frac_of_nominal_8000_gbs": achieved / 8000.0,
        "algorithmic_bytes_per_launch": kb, "kernel_ms": chan_ms_avg, "audio_kernel_ms": audio_ms_avg,
        # serialised kernel times (CUDA events around each kernel); in the timed region above the
        # next block's channel kernel starts under this block's demodulator kernel (programmatic
        # dependent launch), so ms_per_step < kernel_ms + audio_kernel_ms
        "kernel_share_of_step": chan_ms_avg / max(chan_ms_avg + audio_ms_avg, 1e-12),
        "ncu": ncu,
        "receiver_frames_per_s": R * F / (chan_ms_avg * 1e-3) if chan_ms_avg > 0 else 0.0,
        "note": ("shared-tuner workload: every receiver re-uses the one tuner block from L2/shared memory, "
                 "so DRAM traffic is small by construction and the kernel is FP32-issue bound, not HBM bound "
                 "(SURVEY.md 7)") if T < R else "independent streams: each input byte is touched once",
    }

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = cpu_reference_run(w, 1000, 1, max_seconds=args.cpu_seconds)
        cpu = {"value": r["value"], "unit": "MSamples/s", "cores": r["cores"], "kind": r["kind"],
               "sample": r["sample"], "host_cores": r["host_cores"]}
        # the reference as shipped: ONE DSP thread visits every receiver (radio.cxx:56-59)
        r1 = cpu_reference_run(w, 1000, 1, max_seconds=min(4.0, args.cpu_seconds), threads=1)
        cpu["single_thread_value"] = r1["value"]
        cpu["single_thread_sample"] = r1["sample"]
        cpu.update(stock_flags_run(w, min(3.0, args.cpu_seconds)))

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "MSamples/s", "n_gpus": world, "steps": steps,
            "warmup": warmup, "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": bench_config(w, wname, l2_note, args.input),
            "clocks": clk,
            "e2e": {"value": e2e_value, "unit": "MSamples/s", "h2d_bytes_per_step": block_bytes,
                    "d2h_bytes_per_step": 4 * R * M2, "steps": esteps,
                    "mode": f"pipelined depth {depth}, copy-in hand-over by a counter in HBM the channel kernel waits on",
                    "sync_value": e2e_sync_value, "repeats": reps, "values": e2e_all},
            "gpu_launches": launches,
            "roofline": roofline,
            "cpu_baseline": cpu,
            "hbm_gbs_algorithmic_whole_step": algorithmic_bytes(w, frame_bytes) * steps / (ms * 1e-3) / 1e9,
            "tuner_msamples_per_s": world * T * F * steps / (ms * 1e-3) / 1e6,
            "kernel_variant": variant_used,
        }
        print(json.dumps(line), flush=True)
    bank.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


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
        from webradio_b200 import bench_spectrum
        return bench_spectrum.main(args)
    w = synth.WORKLOADS[args.workload]
    if args.impl == "reference":
        reference_arm(args, w, args.workload)
    else:
        gpu_arm(args, w, args.workload)


if __name__ == "__main__":
    main()
