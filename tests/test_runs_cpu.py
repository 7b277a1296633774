"""How the streaming channel kernel (webradio_b200/csrc/wr_kernels_v4.cuh, v4_cut) cuts a bank's
channel-rate outputs into runs -- host logic, no device needed.  Every output of every receiver
belongs to exactly one run, a run is long enough for the kernel's prologue, 32 consecutive runs
touch at most two receivers, and one round of runs fills the grid where the bank allows it."""
import ctypes as C

import pytest

from webradio_b200 import capi

KSKIP = {(255, 50): 6, (127, 50): 3, (127, 40): 4}


def cut(ntaps, decim, R, M1, sms=148):
    L = capi.lib()
    o = [C.c_uint(0) for _ in range(6)]
    ok = L.wr_plan_runs(ntaps, decim, R, M1, sms, *[C.byref(x) for x in o])
    return ok, tuple(x.value for x in o)


CASES = [(255, 50, 1024, 2048), (255, 50, 1024, 2047), (255, 50, 148, 2048), (255, 50, 37, 2048), (255, 50, 5, 2048),
         (255, 50, 2000, 2048), (255, 50, 1024, 512), (255, 50, 5000, 400), (127, 50, 1024, 2048), (127, 40, 777, 2560),
         (127, 50, 3, 8192), (255, 50, 1184, 2048), (255, 50, 1185, 2048), (255, 50, 64, 8192), (255, 50, 40000, 2048)]


@pytest.mark.parametrize("ntaps,decim,R,M1", CASES)
def test_runs_cover_every_output_once(ntaps, decim, R, M1):
    ok, (nr, k, rem, rounds, grid, warps) = cut(ntaps, decim, R, M1)
    assert ok == 1
    assert nr >= 32                                   # a warp's 32 consecutive runs: at most two receivers
    assert rem < nr and rem * (k + 1) + (nr - rem) * k == M1
    assert k >= KSKIP[(ntaps, decim)] + 1             # the prologue's outputs lie inside a receiver's first run
    assert 1 <= grid <= 148 and warps == 8
    assert rounds * grid * warps * 32 >= R * nr       # every run has a lane
    assert (rounds - 1) * grid * warps * 32 < R * nr  # and no round is idle


def test_cfg3_fills_the_grid_in_one_round():
    """BASELINE cfg3: 1024 receivers on 148 x 8 warps.  Whole receivers per warp leave a quarter of the
    schedulers with one warp instead of two (70 periods per warp); 37 runs of 55/56 outputs per
    receiver fill every lane of the grid exactly (61 periods)."""
    ok, (nr, k, rem, rounds, grid, warps) = cut(255, 50, 1024, 2048)
    assert (ok, nr, k, rem, rounds, grid) == (1, 37, 55, 13, 1, 148)
    assert 1024 * nr == grid * warps * 32


def test_unserved():
    assert cut(64, 10, 1024, 10240)[0] == 0           # no v4 instantiation
    assert cut(255, 50, 1024, 100)[0] == 0            # block too short for 32 runs per receiver
