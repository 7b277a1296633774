"""bench.py --workload cfg4 (and the `configs.cfg4` record of the default line): the SpectrumSink FFT
behind the waterfall (K5 + K6).

BASELINE config 4: 8192-point FFT, 50% overlap (hop 4096), 256 receivers (streams) batched on one
B200.  One step = 256 streams x 528384 frames (= 4096 * 129, i.e. 128 FFT frames per stream)
through window + FFT + dB + fft-shift, 32768 transforms.  Algorithmic bytes per transform:
8 * hop read + 4 * N written = 65536 B.
Lives next to bench.py (not in the package): its CPU leg and its in-run check use oracle/.
"""
import json
import os
import time

import numpy as np

from webradio_b200 import synth

N, HOP, STREAMS, ROWS = 8192, 4096, 256, 128
FRAMES = HOP * (ROWS + 1)
METRIC = "input IQ MSamples/s through downconvert→FIR→demod; achieved HBM GB/s vs peak"
DESC = "cfg4: 8192-pt Spectrum FFT, 50% overlap, 256 receivers batched"
MIN_CPU_SECONDS = 2.0


def config():
    return {"workload": DESC, "fft_size": N, "hop": HOP, "n_streams": STREAMS, "frames_per_step": FRAMES,
            "transforms_per_step": STREAMS * ROWS, "parallelism": "streams sharded by assignment, no collective"}


def cpu_reference(max_seconds, first_stream=0):
    """Reference SpectrumSink semantics on the host cores via the oracle port (window, transform,
    dB as spectrumsink.cxx; FFTW itself is not installed -- the float64 stand-in transform is
    stated in oracle/shim/fftw3.h), one stream per thread, the GPU arm's synth blocks."""
    import threading

    from oracle import wro
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    nthreads = min(cores, STREAMS)
    sps = [wro.Spectrum(N, HOP) for _ in range(nthreads)]
    chunk = HOP * 9  # 8 transforms per call
    xs = [synth.lattice_noise(chunk, stream=first_stream + i) for i in range(nthreads)]

    def work(i, reps):
        for _ in range(reps):
            sps[i].process(xs[i], rows=True)

    def run(reps):
        ths = [threading.Thread(target=work, args=(i, reps)) for i in range(nthreads)]
        t0 = time.perf_counter()
        [t.start() for t in ths]
        [t.join() for t in ths]
        return time.perf_counter() - t0

    run(1)
    probe = run(1)
    reps = max(1, int(np.ceil(max(max_seconds, MIN_CPU_SECONDS) / max(probe, 1e-6))))
    secs = run(reps)
    frames = nthreads * reps * chunk
    return {"value": frames / secs / 1e6, "unit": "MSamples/s", "cores": nthreads, "kind": "port",
            "sample": f"{nthreads} streams x {reps} x {chunk} frames (8 transforms per call) in {secs:.1f} s through the "
                      f"oracle port's SpectrumSink (float64 stand-in for FFTW3f), synth.lattice_noise blocks",
            "host_cores": os.cpu_count(), "seconds": secs, "reps": reps}


def parity_check(rows_dev, first_stream):
    """Rows of this run against the oracle on the same synth block (north_star tolerance:
    |mag - mag_ref| <= 1e-5 * max|mag_ref| per transform)."""
    from oracle import wro
    worst = 0.0
    checked = 0
    for s in sorted({int(v) for v in np.linspace(0, STREAMS - 1, 8)}):
        x = synth.lattice_noise(FRAMES, stream=first_stream + s)        # every transform of eight streams
        want = wro.Spectrum(N, HOP).process(x, rows=True).astype(np.float64)
        got = rows_dev[s, :want.shape[0]].cpu().numpy().astype(np.float64)
        mg, mw = 10 ** (got / 20), 10 ** (want / 20)
        rel = np.max(np.abs(mg - mw), axis=1) / np.max(mw, axis=1)
        worst = max(worst, float(rel.max()))
        checked += want.shape[0]
    if worst > 1e-5:
        raise SystemExit(f"bench.py cfg4: spectrum rows differ from the oracle by {worst:g} of the frame peak (> 1e-5)")
    return {"transforms_checked": checked, "max_error_over_frame_peak": worst, "tolerance": 1e-5,
            "oracle": "oracle/libwr_oracle.so SpectrumSink (float64 transform)", "inputs": "identical synth.lattice_noise blocks"}


def record(ctx, steps=None, warmup=3):
    import bench
    torch = ctx.torch
    from webradio_b200 import capi, shard
    args = ctx.args
    steps = steps if steps is not None else 10
    warmup = max(warmup, 3)
    first_stream = ctx.rank * STREAMS
    streams = list(range(first_stream, first_stream + STREAMS))
    # two distinct input batches (1.03 GiB each, far larger than L2): consecutive blocks of each stream
    inputs = [synth.u8_to_f32_torch(synth.lattice_u8_torch(FRAMES, streams, start=b * FRAMES, device="cuda")) for b in range(2)]
    rows = torch.empty(STREAMS, ROWS + 1, N, device="cuda")
    stream = torch.cuda.Stream()

    def fresh():
        # a fresh handle per step: no carry-over, exactly ROWS transforms per stream
        return capi.Spectrum(N, HOP, STREAMS, max_frames=FRAMES, device=ctx.local)

    def step(i, sp, check=True):
        n = sp.process_device(inputs[i % 2].data_ptr(), FRAMES, FRAMES, rows.data_ptr(), (ROWS + 1) * N, stream.cuda_stream)
        assert n == ROWS or not check, n

    h0 = fresh()
    step(0, h0)
    stream.synchronize()
    parity = parity_check(rows, first_stream) if ctx.rank == 0 else None
    handles = [fresh() for _ in range(warmup + steps)]
    for i in range(warmup):
        step(i, handles[i])
    ctx.barrier()
    clocks = bench.ClockSampler(ctx.local)
    clocks.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = sum(h.launch_count() for h in handles)
    ev0.record(stream)
    for i in range(steps):
        step(warmup + i, handles[warmup + i])
    ev1.record(stream)
    ev1.synchronize()
    launches = sum(h.launch_count() for h in handles) - launches0
    ctx.barrier()
    ms = ctx.max_over_ranks([ev0.elapsed_time(ev1)])[0]
    value = shard.job_throughput(STREAMS * FRAMES * steps, ctx.world, ms) / 1e6
    t_end = time.perf_counter() + 0.3
    while time.perf_counter() < t_end:
        step(0, h0, check=False)    # (a reused handle carries half a frame over: 129 rows)
        stream.synchronize()
    clk = clocks.stop()
    del handles

    e2e = None
    if not args.no_e2e:
        # host buffers through wr_spectrum_process (H2D, transforms, D2H of the dB rows)
        esteps = 3
        h_in = inputs[0].cpu().pin_memory()
        h_rows = torch.empty(STREAMS, ROWS + 1, N).pin_memory()
        # ONE handle, warmed by a first call: a handle allocates its device blocks (2 GiB here) on first use, which
        # is not part of a step (a fresh handle per step had put the allocation inside the timed region: 1.2-3.3 k
        # MS/s from run to run).  A reused handle carries half a frame over, so a step yields ROWS or ROWS + 1 rows.
        hs = [fresh()]
        hs[0].L.wr_spectrum_process(hs[0].h, h_in.data_ptr(), FRAMES, h_rows.data_ptr(), (ROWS + 1) * N)
        ctx.barrier()
        t0 = time.perf_counter()
        for i in range(esteps):
            n = hs[0].L.wr_spectrum_process(hs[0].h, h_in.data_ptr(), FRAMES, h_rows.data_ptr(), (ROWS + 1) * N)
            assert n in (ROWS, ROWS + 1)
        torch.cuda.synchronize()
        e2e_s = ctx.max_over_ranks([time.perf_counter() - t0])[0]
        e2e = {"value": ctx.world * STREAMS * FRAMES * esteps / e2e_s / 1e6, "unit": "MSamples/s",
               "h2d_bytes_per_step": 8 * STREAMS * FRAMES, "d2h_bytes_per_step": 4 * STREAMS * ROWS * N, "steps": esteps,
               "mode": "synchronous wr_spectrum_process on pinned host buffers, every dB row copied back"}
        del hs, h_in, h_rows
    del inputs, rows
    torch.cuda.empty_cache()

    peak, peak_src = bench.hbm_peak()
    alg = (8 * HOP + 4 * N) * STREAMS * ROWS
    kernel_ms = ms / steps  # the step IS the kernel (plus a tiny carry kernel)
    achieved = alg / (kernel_ms * 1e-3) / 1e9
    cap = bench.stored_traffic("cfg4")
    cpu = None
    if ctx.rank == 0 and ctx.world == 1 and not args.no_cpu_baseline:
        r = cpu_reference(args.cpu_seconds)
        cpu = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample", "host_cores", "seconds")}
    return {
        "value": value, "unit": "MSamples/s", "steps": steps, "warmup": warmup, "ms_per_step": ms / steps,
        "config": config(), "l2": "two alternating 1.03 GiB input batches (>> L2 126 MiB)", "clocks": clk,
        "e2e": e2e, "gpu_launches": launches,
        "roofline": {"bound": "hbm", "kernel": "spectrum kernel: window + FFT + dB + fft-shift", "achieved": achieved,
                     "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
                     "frac_of_nominal_8000_gbs": achieved / 8000.0,
                     "traffic": cap.get("spectrum_kernel_dram_bytes_per_launch"),
                     "traffic_source": ("stored ncu --set full capture, not measured in this run: " + cap.get("source", "profiles/traffic.json")) if cap else None,
                     "algorithmic_bytes_per_launch": alg, "algorithmic_bytes_formula": "SURVEY.md 8d: (8*hop + 4*N) per transform",
                     "kernel_ms": kernel_ms, "transforms_per_s": STREAMS * ROWS * steps / (ms * 1e-3)},
        "cpu_baseline": cpu, "parity": parity}


def main(args):
    import bench
    rank = int(os.environ.get("RANK", "0"))
    if args.impl == "reference":
        if rank != 0:
            return
        r = cpu_reference(max(args.cpu_seconds, 10.0))
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": r["value"], "unit": "MSamples/s", "n_gpus": args.gpus,
            "steps": r["reps"], "warmup": 1, "ms_per_step": 1e3 * r["seconds"] / r["reps"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config(),
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample", "host_cores")},
            "e2e": {"value": r["value"], "unit": "MSamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}), flush=True)
        return
    ctx = bench.Ctx(args)
    rec = record(ctx, steps=args.steps, warmup=args.warmup if args.warmup is not None else 3)
    if ctx.rank == 0:
        line = {"metric": METRIC, "value": rec["value"], "unit": "MSamples/s", "n_gpus": ctx.world, "steps": rec["steps"],
                "warmup": rec["warmup"], "ms_per_step": rec["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic"}
        line.update({k: rec[k] for k in rec if k not in line})
        print(json.dumps(line), flush=True)
    ctx.close()
