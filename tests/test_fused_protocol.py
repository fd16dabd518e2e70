"""CPU model of the fused kernel's mbarrier / ring protocol (tools/fused_protocol_sim.py): the warp roles of
rusty_sr_b200/csrc/conv_fused.cuh, with the kernel's own index and parity formulas, under a randomised
scheduler — every read must see the data the algorithm expects and the run must not deadlock."""
import os
import random
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import fused_protocol_sim as F  # noqa: E402


@pytest.mark.parametrize("items", [[1], [2], [3], [1, 1], [5, 1, 4], [23], [110, 7]])
def test_protocol_items(items):
    for seed in range(6):
        F.Sim(items, seed=seed, queue_depth=(1, 2, 8)[seed % 3]).run()


def test_protocol_random():
    rng = random.Random(7)
    for k in range(60):
        items = [rng.choice([1, 1, 2, 3, 4, 5, 6, 7, 11]) for _ in range(rng.randint(1, 3))]
        F.Sim(items, seed=1000 + k).run()


def test_model_detects_a_broken_ring():
    """Sanity of the model itself: an l2 ring of 2 slots cannot hold the 3-row window."""
    old = F.NR2
    F.NR2 = 2
    try:
        with pytest.raises((F.Deadlock, AssertionError)):
            F.Sim([7], seed=0).run()
    finally:
        F.NR2 = old


# ---- CTA-pair variant (the library's default plan): two CTAs in lock step, one MMA issuer, multicast commits, remote arrivals
@pytest.mark.parametrize("items", [[1], [2], [3], [1, 1], [5, 1, 4], [23], [60, 7]])
def test_pair_protocol_items(items):
    for seed in range(6):
        F.PairSim(items, seed=seed, queue_depth=(1, 2, 8)[seed % 3]).run()


def test_pair_protocol_random():
    rng = random.Random(11)
    for k in range(80):
        items = [rng.choice([1, 1, 2, 3, 4, 5, 6, 7, 11]) for _ in range(rng.randint(1, 3))]
        F.PairSim(items, seed=2000 + k, queue_depth=rng.choice([1, 2, 3, 8])).run()


@pytest.mark.parametrize("tps,nb,nr1", [(1, 9, 4), (1, 6, 5), (3, 2, 4), (9, 2, 4), (3, 3, 4), (3, 4, 3)])
def test_pair_protocol_other_ring_geometries(tps, nb, nr1):
    """The build-time variants of the pair kernel (RSR_FUSED_PAIR_TPS / _NB / _NR1) run the same protocol."""
    old = (F.PAIR_TPS, F.PAIR_NB, F.PAIR_NR1)
    F.PAIR_TPS, F.PAIR_NB, F.PAIR_NR1 = tps, nb, nr1
    try:
        for seed in range(8):
            F.PairSim([5, 2, 9], seed=seed, queue_depth=(1, 8)[seed % 2]).run()
    finally:
        F.PAIR_TPS, F.PAIR_NB, F.PAIR_NR1 = old


def test_pair_model_detects_a_broken_ring():
    old = F.NR2
    F.NR2 = 2
    try:
        with pytest.raises((F.Deadlock, AssertionError)):
            F.PairSim([7], seed=0).run()
    finally:
        F.NR2 = old
