"""One large frame sharded as horizontal bands across the GPUs of one box (BASELINE config 4).
Launch: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
            --master-port 29533 tools/run_banded.py [W H] [--check]
Scheme: overlap-recompute — rank r receives its rows + min(7, available) halo rows of
the INPUT over NCCL, runs the unmodified pipeline on that sub-image, and sends the centre rows of the
output back to rank 0 over NCCL.  Everything stays in device memory.  Prints one JSON line."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rusty_sr_b200 as R  # noqa: E402
from rusty_sr_b200.sharding import band_plan  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")]
W, H = (int(args[0]), int(args[1])) if len(args) >= 2 else (7680, 4320)
check = "--check" in sys.argv
F = 3
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)

up = R.Upscaler(parameters="imagenet", device=local)
# one explicit stream for NCCL p2p, tensor ops and the forward pass (the legacy default stream has
# handle 0, which rsr_set_stream would read as "use the context's own stream")
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
up.set_stream(stream.cuda_stream)
plan = band_plan(H, world)
me = plan[rank]
rows_in = me.y1 - me.y0 + me.top + me.bottom

frame = None
if rank == 0:
    rng = np.random.default_rng(4321)   # SURVEY.md 8(d) config 4 seed
    frame = torch.from_numpy(rng.integers(0, 256, size=(H, W, 3), dtype=np.uint8)).to(dev).float().div_(255.0)
band_in = frame[me.rows_in] if rank == 0 else torch.empty((rows_in, W, 3), dtype=torch.float32, device=dev)
band_out = torch.empty((rows_in * F, W * F, 3), dtype=torch.float32, device=dev)
full_out = torch.empty((H * F, W * F, 3), dtype=torch.float32, device=dev) if rank == 0 else None


def run_once():
    # scatter input bands (+halo) from rank 0
    if world > 1:
        if rank == 0:
            reqs = [dist.isend(frame[b.rows_in].contiguous(), dst=b.rank) for b in plan[1:] if b.y1 > b.y0]
            for q in reqs:
                q.wait()
        elif rows_in > 0:
            dist.recv(band_in, src=0)
    # forward on the band
    if me.y1 > me.y0:
        up.forward_device(band_in.data_ptr(), 1, W, rows_in, d_out_f32=band_out.data_ptr())
    centre = band_out[me.top * F:(rows_in - me.bottom) * F]
    # gather output rows on rank 0
    if rank == 0:
        full_out[me.y0 * F:me.y1 * F].copy_(centre)
        for b in plan[1:]:
            if b.y1 > b.y0:
                dist.recv(full_out[b.y0 * F:b.y1 * F], src=b.rank)
    elif me.y1 > me.y0:
        dist.send(centre.contiguous(), dst=0)


def timed(fn, n):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(n):
        fn()
    e1.record(stream)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    t = torch.tensor([e0.elapsed_time(e1) / n], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


run_once()  # warm-up (allocates the workspace)
run_once()
ms_total = timed(run_once, 3)


def compute_only():
    if me.y1 > me.y0:
        up.forward_device(band_in.data_ptr(), 1, W, rows_in, d_out_f32=band_out.data_ptr())


ms_compute = timed(compute_only, 3)
result = {"workload": f"{W}x{H} -> {W*F}x{H*F} (factor 3), {world} band(s) of ~{(H + world - 1) // world} rows + 7-row halos",
          "n_gpus": world, "ms_per_frame": ms_total, "ms_compute_only": ms_compute,
          "output_mpix_per_s": W * H * F * F / 1e6 / (ms_total * 1e-3),
          "output_mpix_per_s_compute_only": W * H * F * F / 1e6 / (ms_compute * 1e-3),
          "collective": "NCCL send/recv of input bands (+halo) and output rows; no inter-layer traffic"}
if check and rank == 0:
    # the banded result must equal the unsharded frame computed on one GPU (bit-identical)
    ref = torch.empty_like(full_out)
    up.forward_device(frame.data_ptr(), 1, W, H, d_out_f32=ref.data_ptr())
    up.synchronize()
    result["max_abs_diff_vs_unsharded"] = float((ref - full_out).abs().max().item())
if rank == 0:
    print(json.dumps(result), flush=True)
up.close()
if world > 1:
    dist.destroy_process_group()
