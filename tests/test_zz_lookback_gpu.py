"""The FM look-back sample's way in and out of a receiver bank (wr_rx_get_lookback /
wr_rx_set_lookback) and what it is for -- a receiver that is detached from a live front-end and
attached again (reference radio.cxx:109-117) -- on the CUDA library.  The bodies live in
tests/lookback_cases.py and are dry-run on the CPU stand-in by tests/test_blocks_mock_cpu.py.
(Named to sort last: these entry points were added after the last GPU visit of round 1.)"""
import pytest

import graphlib as G
import lookback_cases as LC
from helpers import fm_exact

pytestmark = pytest.mark.gpu


def test_lookback_carry_over_between_banks(wro):
    from webradio_b200 import capi
    t1 = capi.lowpass_design(LC.N1, 80000, LC.FS)
    t2 = capi.lowpass_design(LC.N2, 8000, LC.FS // LC.D1)
    LC.carry_over(lambda: capi.Bank(1, 1, LC.F, LC.N1, LC.D1, LC.N2, LC.D2), wro, t1, t2,
                  capi.phase_step(LC.IF_HZ, LC.FS), exact_fm=fm_exact())


@pytest.mark.skipif(not (G.have("blocks") and G.have("ref")), reason="harness libraries not built")
@pytest.mark.parametrize("capture", [0x8, 0xF], ids=["fused", "strict"])
def test_hot_detach_and_reattach_on_the_gpu_blocks(capture):
    if not hasattr(G.load("blocks"), "wrh_graph_detach") or not hasattr(G.load("ref"), "wrh_graph_detach"):
        pytest.skip("harness libraries predate wrh_graph_detach")
    LC.hot_reattach("blocks", capture)


@pytest.mark.skipif(not (G.have("blocks") and G.have("ref")), reason="harness libraries not built")
@pytest.mark.parametrize("seed", range(12))
def test_random_scenario_on_the_gpu_blocks(seed):
    """tests/fuzz_cases.py on the CUDA library: seeded sequences of setters, detach / attach and
    restarts over fused, strict and mixed chains, every tapped stage bit for bit against the
    reference blocks (the same seeds run on the CPU stand-in in tests/test_blocks_fuzz_cpu.py)."""
    if not hasattr(G.load("blocks"), "wrh_graph_detach") or not hasattr(G.load("ref"), "wrh_graph_restart"):
        pytest.skip("harness libraries predate wrh_graph_detach / wrh_graph_restart")
    if not fm_exact():
        pytest.skip("FM is bit-exact only against glibc's atan2f; the tolerance path is covered by test_blocks_gpu.py")
    from fuzz_cases import scenario
    scenario(seed, which="blocks")
