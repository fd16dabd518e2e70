"""World-size-2 (and 3) gloo tests of the multi-GPU host logic on CPU: band planning, halo
slicing, scatter/gather.  The per-band compute is the CPU oracle standing in for
Upscaler.forward_band — the thing under test here is the sharding, which is identical on GPU."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, noise_image
from rusty_sr_b200.sharding import HALO, band_plan, batch_plan, upscale_frame_banded


def test_band_plan_covers_every_row_once():
    for H in (1, 5, 13, 540, 1080, 4320):
        for world in (1, 2, 3, 8):
            plan = band_plan(H, world)
            assert plan[0].y0 == 0 and plan[-1].y1 == H
            for a, b in zip(plan, plan[1:]):
                assert a.y1 == b.y0
            for b in plan:
                assert b.top == min(HALO, b.y0) and b.bottom == min(HALO, H - b.y1)
                assert b.rows_in.start >= 0 and b.rows_in.stop <= H
    # 8K frame on 8 GPUs: 540 rows + 7-row halos: 2.6 % redundant rows
    plan = band_plan(4320, 8)
    assert all(b.y1 - b.y0 == 540 for b in plan)
    assert sum(b.top + b.bottom for b in plan) == 14 * 7


def test_batch_plan():
    assert batch_plan(32, 8)[3] == [3, 11, 19, 27]
    assert sorted(sum(batch_plan(10, 4), [])) == list(range(10))
    assert batch_plan(2, 4)[3] == []


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, H, W, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle as O
        import rusty_sr_b200 as R
        O.set_threads(2)
        params = R.load_builtin_params("anime")
        frame = noise_image(H, W, 99) if rank == 0 else None

        def forward_band(rows, top, bottom):
            full = O.forward(params, rows, 3, np.float32)
            return full[top * 3: full.shape[0] - bottom * 3]

        out = upscale_frame_banded(forward_band, frame, H, W, 3)
        if rank == 0:
            ref = O.forward(params, frame, 3, np.float32)
            q.put((out.shape == ref.shape, float(np.abs(out - ref).max())))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,H,W", [(2, 40, 36), (3, 31, 20)])
def test_banded_frame_over_gloo_matches_unsharded(world, H, W):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, H, W, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
        assert p.exitcode == 0
    same_shape, err = q.get(timeout=10)
    assert same_shape
    # per-band zero padding at image edges + 7-row halos elsewhere => identical up to fp32 summation order
    assert err < 1e-5
