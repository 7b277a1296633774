"""bench.py --workload cfg4: the SpectrumSink FFT behind the waterfall (K5 + K6).

BASELINE config 4: 8192-point FFT, 50% overlap (hop 4096), 256 receivers (streams) batched on one
B200.  One step = 256 streams x 528384 frames (= 4096 * 129, i.e. 128 FFT frames per stream)
through window + FFT + dB + fft-shift, 32768 transforms.  Algorithmic bytes per transform:
8 * hop read + 4 * N written = 65536 B.
"""
import json
import os
import time

import numpy as np

N, HOP, STREAMS, ROWS = 8192, 4096, 256, 128
FRAMES = HOP * (ROWS + 1)
METRIC = "input IQ MSamples/s through downconvert→FIR→demod; achieved HBM GB/s vs peak"
DESC = "cfg4: 8192-pt Spectrum FFT, 50% overlap, 256 receivers batched"


def config(l2_note=None):
    c = {"workload": DESC, "fft_size": N, "hop": HOP, "n_streams": STREAMS, "frames_per_step": FRAMES,
         "transforms_per_step": STREAMS * ROWS, "parallelism": "streams sharded by assignment, no collective"}
    if l2_note:
        c["l2"] = l2_note
    return c


def cpu_reference(max_seconds):
    """Reference SpectrumSink semantics on the host cores via the oracle port (window, transform,
    dB as spectrumsink.cxx; FFTW itself is not installed -- the float64 stand-in transform is
    stated in oracle/shim/fftw3.h), one stream per thread."""
    import threading

    from oracle import wro
    from webradio_b200 import synth
    cores = os.cpu_count() or 1
    nthreads = min(cores, STREAMS)
    sps = [wro.Spectrum(N, HOP) for _ in range(nthreads)]
    chunk = HOP * 9  # 8 transforms per call
    x = synth.lattice_noise(chunk, stream=1)

    def work(i, reps):
        for _ in range(reps):
            sps[i].process(x, rows=True)

    def run(reps):
        ths = [threading.Thread(target=work, args=(i, reps)) for i in range(nthreads)]
        t0 = time.perf_counter()
        [t.start() for t in ths]
        [t.join() for t in ths]
        return time.perf_counter() - t0

    run(1)
    probe = run(1)
    reps = max(1, int(max_seconds / max(probe, 1e-6)))
    secs = run(reps)
    frames = nthreads * reps * chunk
    return {"value": frames / secs / 1e6, "unit": "MSamples/s", "cores": nthreads, "kind": "port",
            "sample": f"{nthreads} streams x {reps} x {chunk} frames (8 transforms per call) through the oracle port's "
                      f"SpectrumSink (float64 stand-in for FFTW3f)", "host_cores": cores, "seconds": secs, "reps": reps}


def main(args):
    import bench
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        if rank != 0:
            return
        r = cpu_reference(60.0)
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": r["value"], "unit": "MSamples/s", "n_gpus": args.gpus,
            "steps": r["reps"], "warmup": 1, "ms_per_step": 1e3 * r["seconds"] / r["reps"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config(),
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample", "host_cores")},
            "e2e": {"value": r["value"], "unit": "MSamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}), flush=True)
        return

    import torch
    import torch.distributed as dist

    from webradio_b200 import capi, shard
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    steps = args.steps if args.steps is not None else 10
    warmup = max(args.warmup if args.warmup is not None else 3, 3)

    gen = torch.Generator(device="cuda")
    gen.manual_seed(0xB204 + rank)
    # two distinct input batches (1.03 GiB each, far larger than L2)
    inputs = []
    for _ in range(2):
        u8 = torch.randint(0, 256, (STREAMS, FRAMES, 2), generator=gen, device="cuda", dtype=torch.uint8)
        inputs.append(((u8.float() - 128.0) / 128.0).contiguous())
        del u8
    rows = torch.empty(STREAMS, ROWS + 1, N, device="cuda")
    stream = torch.cuda.Stream()
    sps = []

    def fresh():
        # a fresh handle per step: no carry-over, exactly ROWS transforms per stream
        return capi.Spectrum(N, HOP, STREAMS, max_frames=FRAMES, device=local)

    def step(i, sp):
        n = sp.process_device(inputs[i % 2].data_ptr(), FRAMES, FRAMES, rows.data_ptr(), (ROWS + 1) * N, stream.cuda_stream)
        assert n == ROWS, n

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    handles = [fresh() for _ in range(warmup + steps)]
    for i in range(warmup):
        step(i, handles[i])
    barrier()
    clocks = bench.ClockSampler(local)
    clocks.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = sum(h.launch_count() for h in handles)
    ev0.record(stream)
    for i in range(steps):
        step(warmup + i, handles[warmup + i])
    ev1.record(stream)
    ev1.synchronize()
    launches = sum(h.launch_count() for h in handles) - launches0
    barrier()
    ms = shard.reduce_max_ms(ev0.elapsed_time(ev1), device="cuda")
    value = shard.job_throughput(STREAMS * FRAMES * steps, world, ms) / 1e6
    del handles

    # e2e: host buffers through wr_spectrum_process (H2D, transforms, D2H of the dB rows)
    esteps = min(steps, 3)
    h_in = inputs[0].cpu().pin_memory()
    h_rows = torch.empty(STREAMS, ROWS + 1, N).pin_memory()
    hs = [fresh() for _ in range(esteps + 1)]
    hs[0].L.wr_spectrum_process(hs[0].h, h_in.data_ptr(), FRAMES, h_rows.data_ptr(), (ROWS + 1) * N)
    barrier()
    t0 = time.perf_counter()
    for i in range(esteps):
        n = hs[i + 1].L.wr_spectrum_process(hs[i + 1].h, h_in.data_ptr(), FRAMES, h_rows.data_ptr(), (ROWS + 1) * N)
        assert n == ROWS
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    clk = clocks.stop()
    e2e_value = world * STREAMS * FRAMES * esteps / e2e_s / 1e6

    peaks = os.path.join(bench.ROOT, "MEASURED_PEAKS.json")
    peak = float(json.load(open(peaks))["hbm_gbs"]) if os.path.exists(peaks) else bench.HBM_FALLBACK_GBS
    alg = (8 * HOP + 4 * N) * STREAMS * ROWS
    kernel_ms = ms / steps  # the step IS the kernel (plus a tiny carry kernel)
    achieved = alg / (kernel_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(bench.ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get("cfg4", {}).get("spectrum_kernel_dram_bytes_per_launch")
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = cpu_reference(args.cpu_seconds)
        cpu = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample", "host_cores")}
    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": "MSamples/s", "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": config("two alternating 1.03 GiB input batches (>> L2 126 MiB)"),
            "clocks": clk,
            "e2e": {"value": e2e_value, "unit": "MSamples/s", "h2d_bytes_per_step": 8 * STREAMS * FRAMES,
                    "d2h_bytes_per_step": 4 * STREAMS * ROWS * N, "steps": esteps, "mode": "synchronous"},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": "window + Stockham FFT + dB + fft-shift", "achieved": achieved,
                         "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "frac_of_nominal_8000_gbs": achieved / 8000.0, "traffic": traffic,
                         "algorithmic_bytes_per_launch": alg, "kernel_ms": kernel_ms,
                         "transforms_per_s": STREAMS * ROWS * steps / (ms * 1e-3)},
            "cpu_baseline": cpu}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
