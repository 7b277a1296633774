"""The WR_TRACE / WR_TRACE_CTA summarisers (scripts/trace_summary.py, scripts/cta_summary.py) on
synthetic trace files with a known rhythm: host logic, no GPU."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(script, *args):
    return subprocess.check_output([sys.executable, os.path.join(ROOT, "scripts", script), *map(str, args)], text=True)


def test_trace_summary_reports_the_block_rhythm(tmp_path):
    p = tmp_path / "trace.csv"
    with open(p, "w") as f:
        f.write("# bank R=64 T=1 F=102400 in=1 out=0 depth=6\n")
        f.write("seq,host_submit,host_submitted,host_waited,chan_start,chan_input,chan_end,demod_start,demod_end\n")
        t = 1_000_000
        # 60 device-resident blocks, 20 us apart: channel kernel 14 us, demodulator starts 1 us later, runs 6 us
        for i in range(60):
            f.write(f"{i + 1},0,0,0,{t},{t + 100},{t + 14000},{t + 15000},{t + 21000}\n")
            t += 20000
        t += 5_000_000
        # 60 host-path blocks, 30 us apart, the input arrives 4 us after the kernel started
        for i in range(60):
            h = 2 * t
            f.write(f"{i + 61},{h},{h + 8000},{h + 150000},{t},{t + 4000},{t + 18000},{t + 19000},{t + 25000}\n")
            t += 30000
    out = run("trace_summary.py", p, 10, 50)
    assert "device-resident" in out and "host path" in out
    dev = next(l for l in out.splitlines() if "device-resident" in l)
    host = next(l for l in out.splitlines() if "host path" in l)
    assert "device step 20.00 us" in dev and "chan 14.00 us (input wait 0.10)" in dev and "demod 6.00 us" in dev
    assert "chan(n+1) starts 1.00 us before demod(n) ends" in dev
    assert "device step 30.00 us" in host and "(input wait 4.00)" in host
    assert "submit 8.00 us/call" in out and "submit period 60.00 us" in out


def test_cta_summary_reports_start_and_life_times(tmp_path):
    p = tmp_path / "cta.csv"
    with open(p, "w") as f:
        f.write("# bank R=64 T=1 F=102400\nblock,kernel,cta,start,end\n")
        for blk in range(3):
            base = 1_000_000 + blk * 20000
            for c in range(8):
                f.write(f"{300 + blk},chan,{c},{base + 100 * c},{base + 14000 + 100 * c}\n")
            for c in range(16):
                f.write(f"{300 + blk},demod,{c},{base + 15000},{base + 18000 + 10 * c}\n")
    out = run("cta_summary.py", p)
    assert "block 301: 8 channel CTAs, 16 demodulator CTAs" in out
    start = next(l for l in out.splitlines() if "channel CTAs (mixers):" in l)
    assert "min   0.0" in start and "max   0.7" in start
