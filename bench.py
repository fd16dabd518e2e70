#!/usr/bin/env python
"""bench.py — throughput of the rusty_sr forward path on B200 (one JSON line on stdout).

    python bench.py --gpus N --steps K --warmup W            # our sm_100a path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port)

Metric (BASELINE.json): output Mpix/s of the super-resolution forward pass.  Headline workload (BASELINE config 2) at
every N: one 1920x1080 RGB frame per GPU per step -> 5760x3240 (factor 3: the reference is compiled with FACTOR = 3 and
ships factor-3 weights only, main.rs:31), bundled `imagenet` weights, uniform-random u8 input
(numpy default_rng(1234 + rank)).  Image batches shard across GPUs with no data-path collective, so scaling is "weak":
every rank upscales its own frame.

  value   : frames resident in HBM (f32 [H][W][3] in, f32 [3H][3W][3] out: the graph.forward contract, main.rs:171),
            timed with CUDA events on the launching stream, max over ranks.  `sustained` repeats the step for >= 2 s.
  e2e     : the same metric through the reference-facing host calls with pinned HOST buffers, H2D and D2H inside the
            timed region: rsr_submit_u8_rgb / rsr_wait (u8 RGB in -> u8 RGB out = img_to_data -> forward -> data_to_img,
            main.rs:170-175; `.to_rgba()` is a host step in the reference); the RGBA and synchronous forms beside it.
  roofline: the dominant launch (conv1|2|3, 59 % of the FLOPs), algorithmic FLOPs / its CUDA-event duration, against the
            measured dense-bf16 peak of MEASURED_PEAKS.json; `per_launch_roofline` has every launch.
  config3 : BASELINE config 3 — 32 x 256x256 u8 patches (seed 1234 + i), image i -> rank i mod N, no collective.
  config4 : BASELINE config 4 — ONE 7680x4320 u8 frame (seed 4321) as N horizontal bands; every rank owns its band of
            the input in HBM, the 7 halo rows per side are exchanged with NCCL send/recv inside the timed region, the
            output stays band-sharded.  Speed-up is against the whole frame on rank 0 alone in the same run; the result
            is checked (checksums) to be identical to the unsharded frame.
  config5 : BASELINE config 5 — per-layer 3x3 conv sweep (N = 1 only).
  cpu_baseline: the CPU oracle (port of the reference path; the Rust reference cannot be built here) on the host
            cores, on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W_LR, H_LR, FACTOR = 1920, 1080, 3
OUT_MPIX = W_LR * H_LR * FACTOR * FACTOR / 1e6          # 18.6624 Mpix per frame
FLOPS_PER_LR_PX = 260352                                # SURVEY.md 8(a), whole net
FLOPS_CONV123_PER_LR_PX = 3 * 51200                     # conv1 + conv2 + conv3 (network.rs:60-62)
# Per launch: (useful FLOP per LR pixel, algorithmic HBM bytes per LR pixel) — DESIGN.md §5 table; every tensor read once + written
# once in its stored layout, padding lanes not counted as useful FLOPs.  The fused middle launch moves 640 B/px (l1, P2, P3 in;
# l3, E out) where conv5|6|7 + conv8|9 move 768 + 640.
LAUNCH_WORK = {
    "conv0": (4800, 12 + 128 + 16),      # image in; f and the float4 copy of x (bilinear skip) out
    "conv1|2|3": (153600, 128 + 384),
    "conv5|6|7": (18432 * 2 + 15552, 384 + 384),
    "conv8|9": (18432 + 15552, 384 + 256),
    "conv5|6|7+conv8|9 (fused)": (18432 * 3 + 15552 * 2, 384 + 256),
    "conv10+expand+skip": (15552, 256 + 16 + 108),
}
# The reference's per-layer roofline (fp32 activations read once + written once per reference conv):
# per LR pixel (FLOP, bytes) of the ten convolutions — the yardstick of BASELINE.json's ">= 70 % of the per-layer conv roofline".
REF_LAYERS = [(4800, 140)] + [(51200, 256)] * 3 + [(18432, 256)] * 3 + [(15552, 236)] * 3
WORKLOAD = f"{W_LR}x{H_LR} RGB -> {W_LR*FACTOR}x{H_LR*FACTOR} (factor 3, bundled imagenet weights), 1 frame/GPU/step"
N_ROTATE = 6   # distinct input frames cycled through so the input is never L2 resident
HALO = 7       # receptive field of sr_net in rows (2 + 2 + 1 + 1 + 1; network.rs:33,60-72)


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            d = json.load(fh)
        return {"bf16_tflops": float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1400.0))),
                "bf16_tflops_burst": float(d.get("bf16_tflops", 1590.0)),
                "hbm_gbs": float(d.get("hbm_gbs", 6650.0)), "source": "measured (MEASURED_PEAKS.json)"}
    return {"bf16_tflops": 1400.0, "bf16_tflops_burst": 1590.0, "hbm_gbs": 6650.0,
            "source": "fallback (B200_PROFILING.md)"}


def per_launch_roofline(ktimes, npx, peaks, mma_per_chunk=3):
    """Every launch against min(tensor, HBM): roofline_ms = max(useful FLOPs / measured sustained dense-bf16 peak, algorithmic
    bytes / measured copy bandwidth), frac = roofline_ms / measured ms.  *_executed: the same with the tensor-pipe work the
    operand split actually executes (`mma_per_chunk` MMAs per 16-channel chunk instead of 1) against the measured burst peak —
    the ceiling of the formulation."""
    per_launch = {}
    for k, ms in ktimes.items():
        if k not in LAUNCH_WORK or not ms or ms != ms:
            continue
        fl, by = LAUNCH_WORK[k]
        t_tensor = fl * npx / (peaks["bf16_tflops"] * 1e12) * 1e3
        t_hbm = by * npx / (peaks["hbm_gbs"] * 1e9) * 1e3
        # conv0 always runs the three-product bf16 split on its 16-wide K chunks
        t_x = (3 if k == "conv0" else mma_per_chunk) * fl * npx / (peaks["bf16_tflops_burst"] * 1e12) * 1e3
        per_launch[k] = {"ms": ms, "useful_tflops": fl * npx / (ms * 1e-3) / 1e12, "hbm_gbs": by * npx / (ms * 1e-3) / 1e9,
                         "bound": "tensor" if t_tensor > t_hbm else "hbm", "roofline_ms": max(t_tensor, t_hbm),
                         "frac": max(t_tensor, t_hbm) / ms,
                         "roofline_ms_executed": max(t_x, t_hbm), "frac_executed": max(t_x, t_hbm) / ms}
    return per_launch


def reference_layer_roofline_ms(npx, peaks, burst=False):
    """Sum over the reference's ten convolutions of max(FLOPs / dense-bf16 peak, fp32 bytes / copy bandwidth)."""
    tf = peaks["bf16_tflops_burst" if burst else "bf16_tflops"] * 1e12
    return sum(max(fl * npx / tf, by * npx / (peaks["hbm_gbs"] * 1e9)) for fl, by in REF_LAYERS) * 1e3


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed regions."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.marks = {}
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def mark(self, name):
        """Call at the start and at the end of a region: two mark('x') calls bracket region 'x'."""
        self.marks.setdefault(name, []).append(time.perf_counter())

    @staticmethod
    def _stats(rows):
        sm, smax, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for _, r in rows:
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[0])); smax.append(float(r[1])); power.append(float(r[2]))
            except ValueError:
                continue
            for name, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return sm, smax, power, reasons

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = self._stats(self.rows)
        thr = 0.6 * max(power) if power else 0.0
        busy = [c for c, p in zip(sm, power) if p >= thr] or sm
        out = {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": max(smax) if smax else None,
               "power_w_max": max(power) if power else None, "samples": len(sm), "samples_under_load": len(busy),
               "sampled_over": "whole bench run, 50 ms period; sm_mhz = median over the samples at >= 60 % of the peak power",
               "reasons": sorted(reasons)}
        for name, ts in self.marks.items():
            if len(ts) >= 2:
                s2, _, p2, r2 = self._stats([r for r in self.rows if ts[0] <= r[0] <= ts[1]])
                if s2:
                    out[name] = {"sm_mhz": statistics.median(s2), "sm_mhz_min": min(s2), "power_w_median": statistics.median(p2),
                                 "samples": len(s2), "reasons": sorted(r2)}
        return out


def make_inputs(rank: int):
    rng = np.random.default_rng(1234 + rank)
    return rng.integers(0, 256, size=(N_ROTATE, H_LR, W_LR, 3), dtype=np.uint8)


# ----------------------------------------------------------------------------- reference arm
def cpu_sample(rows: int, steps: int, warmup: int):
    """Times the CPU oracle (img_to_data -> forward fp32 -> data_to_img) on a rows x 1920 slab."""
    from oracle import oracle as O
    import rusty_sr_b200 as R
    params = R.load_builtin_params("imagenet")
    # use every host core this process may run on, even when the launcher (torchrun) exported OMP_NUM_THREADS=1
    try:
        O.set_threads(len(os.sched_getaffinity(0)))
    except AttributeError:
        O.set_threads(os.cpu_count() or 1)
    img = make_inputs(0)[0][:rows]
    for _ in range(warmup):
        O.data_to_img(O.forward(params, O.img_to_data(img[:8]), 3, np.float32))
    times = []
    for _ in range(steps):
        t = time.perf_counter()
        O.data_to_img(O.forward(params, O.img_to_data(img), 3, np.float32))
        times.append(time.perf_counter() - t)
    mpix = rows * W_LR * FACTOR * FACTOR / 1e6
    return mpix, times, O.max_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0  # the CPU arm runs on rank 0 only
    rows = 540  # half of a 1080p frame per step: enough rows to occupy 64-128 host threads, ~2 s per step
    mpix, times, threads = cpu_sample(rows, args.steps, min(args.warmup, 3))
    total = sum(times)
    value = mpix * len(times) / total
    sample = f"{W_LR}x{rows} slab of the 1080p frame per step (+7-row receptive field not needed: own zero padding)"
    line = {
        "impl": "reference", "metric": "output Mpix/s (3x upscale, sr_net forward)", "value": value, "unit": "Mpix/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": WORKLOAD, "factor": FACTOR, "weights": "imagenet.rsr"},
        "cpu_baseline": {"value": value, "unit": "Mpix/s", "cores": threads, "kind": "port", "sample": sample,
                         "note": "CPU restatement of rusty_sr/alumina path (Rust toolchain + alumina crate unavailable); "
                                 "OpenMP over all host cores, the real reference is single-threaded"},
        "e2e": {"value": value, "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


# ----------------------------------------------------------------------------------- our arm
class Dist:
    """The few torch.distributed calls the bench needs (plumbing only)."""

    def __init__(self, torch, dist, world, rank, dev):
        self.torch, self.dist, self.world, self.rank, self.dev = torch, dist, world, rank, dev

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def max(self, v: float) -> float:
        t = self.torch.tensor([v], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def gather_i64(self, vals):
        """[len(vals)] int64 per rank -> [world][len(vals)] on every rank."""
        t = self.torch.tensor(list(vals), dtype=self.torch.int64, device=self.dev)
        if self.world == 1:
            return [t.tolist()]
        out = [self.torch.empty_like(t) for _ in range(self.world)]
        self.dist.all_gather(out, t)
        return [o.tolist() for o in out]


def timed(torch, stream, D, fn, steps):
    """steps x fn(i) between barriers, CUDA events on `stream`; returns the max over ranks of ms per step."""
    D.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(steps):
        fn(i)
    e1.record(stream)
    D.barrier()
    return D.max(e0.elapsed_time(e1)) / steps


def timed_local(torch, stream, dev, fn, steps):
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(steps):
        fn(i)
    e1.record(stream)
    torch.cuda.synchronize(dev)
    return e0.elapsed_time(e1) / steps


def checksum(torch, t):
    """Two 63-bit checksums of a u8 tensor: sum of bytes, and sum of byte x (position hash)."""
    v = t.reshape(-1).to(torch.int64)
    idx = torch.arange(v.numel(), device=v.device, dtype=torch.int64)
    return int(v.sum().item()), int((v * ((idx * 40503) % 1000003 + 1)).sum().item())


def bench_config3(torch, R, up, stream, D, args):
    """BASELINE config 3: 32 x 256x256 u8 RGB patches (default_rng(1234 + i)); image i -> rank i mod N; weights replicated,
    no collective.  u8 RGB in HBM -> u8 RGB in HBM.  Baseline of the speed-up: all 32 on rank 0 alone, same run."""
    NIMG, S = 32, 256
    world, rank, dev = D.world, D.rank, D.dev
    mine = list(range(rank, NIMG, world))
    imgs = np.stack([np.random.default_rng(1234 + i).integers(0, 256, size=(S, S, 3), dtype=np.uint8) for i in range(NIMG)])
    d_all = torch.from_numpy(imgs).to(dev)
    d_in = d_all[mine].contiguous() if mine else None
    d_out = torch.empty((max(len(mine), 1), S * FACTOR, S * FACTOR, 3), dtype=torch.uint8, device=dev)

    def step(_):
        if mine:
            up.forward_device_u8_rgb(d_in.data_ptr(), len(mine), S, S, 3, d_out.data_ptr())

    steps = max(args.steps, 10)
    ms1, ref_sums = None, None
    with torch.cuda.stream(stream):
        for i in range(3):
            step(i)
        ms = timed(torch, stream, D, step, steps)
        plan = up.plan()
        sums = checksum(torch, d_out[:len(mine)]) if mine else (0, 0)
        if rank == 0:   # single-GPU time of the same run and the reference result: the whole batch on rank 0
            d_out1 = torch.empty((NIMG, S * FACTOR, S * FACTOR, 3), dtype=torch.uint8, device=dev)
            one = lambda _: up.forward_device_u8_rgb(d_all.data_ptr(), NIMG, S, S, 3, d_out1.data_ptr())
            for i in range(3):
                one(i)
            ms1 = timed_local(torch, stream, dev, one, steps)
            ref_sums = [checksum(torch, d_out1[list(range(r, NIMG, world))]) if r < NIMG else (0, 0) for r in range(world)]
        all_sums = D.gather_i64(sums)
    D.barrier()
    if rank != 0:
        return None
    mpix = NIMG * S * S * FACTOR * FACTOR / 1e6
    return {"workload": f"{NIMG} x {S}x{S} u8 RGB patches -> {S*FACTOR}x{S*FACTOR} u8 RGB, seeds 1234+i, image i -> rank i mod {world}, resident in HBM",
            "ms_per_batch": ms, "value": mpix / (ms * 1e-3), "unit": "Mpix/s", "images_per_rank": [len(range(r, NIMG, world)) for r in range(world)],
            "single_gpu_ms_per_batch": ms1, "single_gpu_value": mpix / (ms1 * 1e-3), "speedup_vs_single_gpu": ms1 / ms,
            "identical_to_single_gpu": [list(a) for a in all_sums] == [list(b) for b in ref_sums], "launch_plan": plan,
            "collective": "none (independent images)"}


def bench_config4(torch, dist, R, up, stream, D, args):
    """BASELINE config 4: one 7680x4320 u8 RGB frame (default_rng(4321)) -> 23040x12960, as `world` horizontal bands.
    Every rank owns the rows of its band in HBM; per step it exchanges the 7 halo rows per side with its neighbours
    (NCCL send/recv, inside the timed region) and upscales band + halo (overlap-recompute, rsr_upscale_band_device_u8);
    the u8 RGB output stays band-sharded on the ranks."""
    import ctypes
    W, H = 7680, 4320
    world, rank, dev = D.world, D.rank, D.dev
    frame = np.random.default_rng(4321).integers(0, 256, size=(H, W, 3), dtype=np.uint8)   # every rank derives the same frame
    y0, y1 = rank * H // world, (rank + 1) * H // world
    top, bot = min(HALO, y0), min(HALO, H - y1)
    rows = y1 - y0
    band = torch.zeros((top + rows + bot, W, 3), dtype=torch.uint8, device=dev)
    band[top:top + rows].copy_(torch.from_numpy(frame[y0:y1]))          # own rows only; the halo rows arrive over NCCL
    out = torch.empty((rows * FACTOR, W * FACTOR, 3), dtype=torch.uint8, device=dev)

    def exchange():
        if world == 1:
            return
        ops = []
        if rank > 0:          # own rows [y0, y0+7) -> bottom halo of rank-1; rows [y0-7, y0) <- rank-1
            ops.append(dist.P2POp(dist.isend, band[top:top + HALO], rank - 1))
            ops.append(dist.P2POp(dist.irecv, band[:top], rank - 1))
        if rank < world - 1:
            ops.append(dist.P2POp(dist.isend, band[top + rows - HALO:top + rows], rank + 1))
            ops.append(dist.P2POp(dist.irecv, band[top + rows:], rank + 1))
        for w in dist.batch_isend_irecv(ops):
            w.wait()

    fwd = lambda _=0: up.forward_band_device(band.data_ptr(), W, top + rows + bot, top, bot, out.data_ptr(), u8_channels=3, out_channels=3)

    def step(_):
        exchange()
        fwd()

    steps = max(3, min(args.steps, 10))
    ms1, ref_sums, e2e = None, None, None
    with torch.cuda.stream(stream):
        for i in range(2):
            step(i)
        ms = timed(torch, stream, D, step, steps)
        ms_compute = timed(torch, stream, D, fwd, steps)
        sums = checksum(torch, out)
        all_sums = D.gather_i64(sums)
        if rank == 0:
            # the whole frame on this GPU alone: the N = 1 time of the same run and the reference result
            d_frame = torch.from_numpy(frame).to(dev)
            d_full = torch.empty((H * FACTOR, W * FACTOR, 3), dtype=torch.uint8, device=dev)
            one = lambda _: up.forward_device_u8_rgb(d_frame.data_ptr(), 1, W, H, 3, d_full.data_ptr())
            one(0)
            ms1 = timed_local(torch, stream, dev, one, steps)
            ref_sums = [checksum(torch, d_full[(r * H // world) * FACTOR:((r + 1) * H // world) * FACTOR]) for r in range(world)]
            del d_full, d_frame
            torch.cuda.empty_cache()
    D.barrier()
    if rank == 0:
        # end to end from ONE host process (what the reference CLI is): rsr_multi_upscale_frame_u8 over all N GPUs,
        # pinned host u8 RGB in, RGBA u8 out in pinned host memory (one host thread per device inside the call)
        try:
            pin_in = torch.from_numpy(frame).pin_memory()
            pin_out = torch.empty((H * FACTOR, W * FACTOR, 4), dtype=torch.uint8).pin_memory()
            L = R._lib.load()
            with R.MultiUpscaler(parameters="imagenet", devices=list(range(world))) as m:
                def call():
                    R._lib.check(L.rsr_multi_upscale_frame_u8(m._h, ctypes.c_void_p(pin_in.numpy().ctypes.data), W, H, 3,
                                                              ctypes.c_void_p(pin_out.numpy().ctypes.data)))
                call()
                ts = []
                for _ in range(3):
                    t = time.perf_counter()
                    call()
                    ts.append(time.perf_counter() - t)
                e2e = {"call": f"rsr_multi_upscale_frame_u8 over {m.size()} device(s) from one process, pinned host u8 RGB -> pinned host u8 RGBA",
                       "ms_per_frame": 1e3 * min(ts), "value": W * H * 9 / 1e6 / min(ts), "unit": "Mpix/s",
                       "h2d_bytes": W * H * 3, "d2h_bytes": W * H * 9 * 4,
                       "limiter": "device->host copies of 1.19 GB RGBA into one host buffer (PCIe links / host DRAM), not the kernels"}
            del pin_in, pin_out
        except Exception as ex:  # noqa
            e2e = {"error": str(ex)[:300]}
    D.barrier()
    if rank != 0:
        return None
    mpix = W * H * FACTOR * FACTOR / 1e6
    return {"workload": f"one {W}x{H} u8 RGB frame (seed 4321) -> {W*FACTOR}x{H*FACTOR} u8 RGB; {world} band(s) of ~{H // world} rows + {HALO}-row halos, "
                        "input and output band-sharded in HBM",
            "ms_per_frame": ms, "value": mpix / (ms * 1e-3), "unit": "Mpix/s",
            "ms_per_frame_compute_only": ms_compute,
            "halo_exchange": ("NCCL send/recv of 7 input rows per side (161 KB) per step, inside the timed region" if world > 1 else "none (one band)"),
            "single_gpu_ms_per_frame": ms1, "single_gpu_value": mpix / (ms1 * 1e-3),
            "speedup_vs_single_gpu": ms1 / ms, "speedup_compute_only": ms1 / ms_compute,
            "redundant_rows_frac": (2 * HALO * (world - 1)) / H,
            "identical_to_unsharded_frame": [list(a) for a in all_sums] == [list(b) for b in ref_sums],
            "e2e_one_process": e2e}


def bench_config5(R, peaks):
    """BASELINE config 5: 3x3 stride-1 zero-padded conv, C_in = C_out = C, N(0,1)/sqrt(9C) data (seed 99); kernel time only."""
    rows = []
    rng = np.random.default_rng(99)
    for C, S in ((32, 1024), (32, 2048), (64, 1024), (64, 2048), (128, 1024), (256, 1024)):
        x = rng.standard_normal((S, S, C), dtype=np.float32)
        w = rng.standard_normal((C, 3, 3, C), dtype=np.float32) / np.float32(np.sqrt(9 * C))
        _, ms = R.conv2d_3x3(x, w, iters=5)
        fl = 2 * 9 * C * C * S * S
        by = 4 * 2 * C * S * S + 4 * 9 * C * C
        roof = max(fl / (peaks["bf16_tflops"] * 1e12), by / (peaks["hbm_gbs"] * 1e9)) * 1e3
        rows.append({"C": C, "HW": S, "ms": ms, "tflops_useful": fl / (ms * 1e-3) / 1e12, "gbs_fp32_algorithmic": by / (ms * 1e-3) / 1e9,
                     "roofline_ms": roof, "frac": roof / ms})
    return {"workload": "3x3 conv, stride 1, zero pad, Cin = Cout = C, fp32 in/out (rsr_conv2d_3x3_f32: fp16 + mxfp8 operands; output blocks of <= 64 channels with row-folded MMAs, wider outputs as blocks of 128 on CTA pairs)", "rows": rows}


def run_ours(args):
    import ctypes
    import torch
    import torch.distributed as dist
    import rusty_sr_b200 as R

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        emit({"error": "no CUDA device: rusty_sr_b200 has no CPU fallback"})
        return 2
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    D = Dist(torch, dist, world, rank, dev)
    peaks = load_peaks()

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    mark = (lambda name: sampler.mark(name)) if sampler else (lambda name: None)
    plan_flags = {"default": 0, "fused": R.RSR_FLAG_FUSED_MID, "split": R.RSR_FLAG_SPLIT_MID}[args.plan]
    if args.prec == "bf16x3":
        plan_flags |= R.RSR_FLAG_BF16X3
    up = R.Upscaler(parameters="imagenet", device=local, flags=plan_flags)
    stream = torch.cuda.Stream(device=dev)
    up.set_stream(stream.cuda_stream)

    imgs_u8 = make_inputs(rank)                                          # [6][1080][1920][3] u8
    x_host = imgs_u8.astype(np.float32) / 255.0                          # img_to_data (main.rs:170)
    d_in = torch.from_numpy(x_host).to(dev)                              # f32 frames resident in HBM
    d_out = torch.empty((H_LR * FACTOR, W_LR * FACTOR, 3), dtype=torch.float32, device=dev)
    frame_elems = H_LR * W_LR * 3
    launches = [0]

    def step(i):
        ptr = d_in.data_ptr() + (i % N_ROTATE) * frame_elems * 4
        up.forward_device(ptr, 1, W_LR, H_LR, d_out_f32=d_out.data_ptr())

    def counted_step(i):
        step(i)
        launches[0] += up.last_launch_count()

    # ---- device-resident throughput ("value"): K steps (burst regime), then the same step for >= 2 s (sustained) ----
    with torch.cuda.stream(stream):
        for i in range(args.warmup):
            step(i)
        mark("timed_region")
        ms_step = timed(torch, stream, D, counted_step, args.steps)
        mark("timed_region")
        # per-kernel durations (CUDA events the library records on the same stream around every launch), first in the same
        # (burst) regime as `value`, then again right after the sustained loop (power-capped clocks)
        def kernel_breakdown():
            runs = []
            up.set_kernel_timing(True)      # events between the launches: serialises them, so only for this breakdown
            for i in range(min(args.steps, 10)):
                step(i)
                runs.append(up.last_kernel_times_ms())
            up.set_kernel_timing(False)
            kt = {k: statistics.mean(r[k] for r in runs) for k in runs[0]}
            if up.last_launch_count() == 4:      # conv5|6|7 and conv8|9 run as one launch (conv_fused.cuh)
                kt["conv5|6|7+conv8|9 (fused)"] = kt.pop("conv5|6|7") + kt.pop("conv8|9")
            return kt

        ktimes = kernel_breakdown()
        plan = up.plan()
        n_sus = max(args.steps, int(args.sustain_s * 1e3 / ms_step)) if args.sustain_s > 0 else 0
        ms_sus, ktimes_sus = None, None
        if n_sus:
            mark("sustained")
            ms_sus = timed(torch, stream, D, step, n_sus)
            mark("sustained")
            ktimes_sus = kernel_breakdown()
    value = world * OUT_MPIX / (ms_step * 1e-3)

    # ---- end to end through the host calls, pinned buffers ---------------------------------------
    # The call a user makes for a stream of frames: rsr_submit_u8[_rgb] / rsr_wait (3 frames in flight, so the H2D of
    # frame k+1 and the D2H of frame k-1 overlap the forward pass of frame k).  Every step moves its input from pinned host
    # memory to the device and its result back to pinned host memory.
    DEPTH = 3
    pin_in = torch.from_numpy(imgs_u8).pin_memory()
    in_np = pin_in.numpy()

    def e2e_measure(out_ch):
        pin_out = torch.empty((DEPTH, H_LR * FACTOR, W_LR * FACTOR, out_ch), dtype=torch.uint8).pin_memory()
        out_np = pin_out.numpy()

        def run(n):
            tickets = []
            for i in range(n):
                if i >= DEPTH:
                    up.wait(tickets[i - DEPTH])      # the output buffer about to be reused has landed
                tickets.append(up.submit_u8(in_np[i % N_ROTATE], out_np[i % DEPTH]))
            up.wait_all()

        run(max(args.warmup, DEPTH))
        D.barrier()
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e2.record(stream)
        t_wall = time.perf_counter()
        run(args.steps)              # returns after the last D2H copy has completed
        t_wall = time.perf_counter() - t_wall
        e3.record(stream)
        D.barrier()
        # device-side bracket (events on the compute stream) and host wall clock agree because the region ends with a
        # host-synchronous wait on the last D2H; report the larger of the two
        return D.max(max(e2.elapsed_time(e3), t_wall * 1e3)) / args.steps, pin_out

    e2e_rgb_ms, _ = e2e_measure(3)
    e2e_rgba_ms, pin_rgba = e2e_measure(4)
    # the plain synchronous call, one frame at a time, for reference
    L = R._lib.load()
    out0 = pin_rgba.numpy()[0]
    t_sync = time.perf_counter()
    for i in range(min(args.steps, 10)):
        R._lib.check(L.rsr_upscale_u8(up._h, ctypes.c_void_p(in_np[i % N_ROTATE].ctypes.data), W_LR, H_LR, 3,
                                      ctypes.c_void_p(out0.ctypes.data)))
    sync_ms_step = (time.perf_counter() - t_sync) * 1e3 / min(args.steps, 10)
    # host <-> device copy bandwidth of this box, every rank copying at once (names the e2e limiter)
    d_tmp = torch.empty((H_LR * FACTOR, W_LR * FACTOR, 4), dtype=torch.uint8, device=dev)
    pin_tmp = pin_rgba[0]
    D.barrier()
    t = time.perf_counter()
    for _ in range(5):
        pin_tmp.copy_(d_tmp, non_blocking=True)
    torch.cuda.synchronize(dev)
    d2h_gbs = 5 * d_tmp.numel() / (time.perf_counter() - t) / 1e9
    D.barrier()
    t = time.perf_counter()
    for _ in range(5):
        d_tmp.copy_(pin_tmp, non_blocking=True)
    torch.cuda.synchronize(dev)
    h2d_gbs = 5 * d_tmp.numel() / (time.perf_counter() - t) / 1e9
    d2h_all = D.gather_i64([int(d2h_gbs * 1000)])
    del d_tmp, pin_tmp, pin_rgba

    # ---- BASELINE configs 3 and 4 (every rank takes part) and 5 (single GPU) ------------------------
    cfg3 = cfg4 = cfg5 = None
    if not args.skip_configs:
        mark("config3_4")
        cfg3 = bench_config3(torch, R, up, stream, D, args)
        cfg4 = bench_config4(torch, dist, R, up, stream, D, args)
        mark("config3_4")
        if world == 1 and rank == 0:
            cfg5 = bench_config5(R, peaks)
    clocks = sampler.stop() if sampler else None

    if rank == 0:
        npx = W_LR * H_LR
        k_ms = ktimes.get("conv1|2|3", float("nan"))
        achieved_tf = FLOPS_CONV123_PER_LR_PX * npx / (k_ms * 1e-3) / 1e12
        mma_per_chunk = 3 if args.prec == "bf16x3" else 2
        traffic = None
        prof = os.path.join(ROOT, "profiles", "ncu_summary.json")
        if os.path.exists(prof):
            try:
                traffic = json.load(open(prof)).get("conv1|2|3", {}).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        cpu = None
        if world == 1:
            rows = 540
            cpu_mpix, cpu_times, cpu_threads = cpu_sample(rows, 2, 1)
            cpu = {"value": cpu_mpix * len(cpu_times) / sum(cpu_times), "unit": "Mpix/s", "cores": cpu_threads,
                   "kind": "port",
                   "sample": f"2 x ({W_LR}x{rows} slab of the same frame), img_to_data+forward(fp32)+data_to_img",
                   "note": "CPU restatement of rusty_sr/alumina path (Rust toolchain unavailable)"}
        total_k = sum(ktimes.values())
        per_launch = per_launch_roofline(ktimes, npx, peaks, mma_per_chunk)
        ref_roof = reference_layer_roofline_ms(npx, peaks)
        line = {
            "metric": "output Mpix/s (3x upscale, sr_net forward)", "value": value, "unit": "Mpix/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": ("bf16x3 split operands" if args.prec == "bf16x3" else "fp16 + 2 x fp8(e4m3) split operands")
                     + ", f32 accumulate (fp32-equivalent within the 1e-3 contract)",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "factor": FACTOR, "weights": "imagenet.rsr",
                       "parallelism": f"image-batch sharding x{world}, no collective",
                       "launch_plan": plan,
                       "l2": f"working set ~1.3 GB/step >> 126 MB L2; input rotates over {N_ROTATE} distinct frames (149 MB)",
                       "io": "value: f32 in/out resident in HBM; e2e: u8 RGB host -> u8 RGB host (pinned)"},
            "sustained": None if ms_sus is None else {
                "ms_per_step": ms_sus, "value": world * OUT_MPIX / (ms_sus * 1e-3), "steps": n_sus,
                "note": f"the same step repeated for >= {args.sustain_s:g} s (power-capped steady state); `value` is the K-step burst figure"},
            "roofline": {"kernel": "conv1|2|3 (" + ("conv_pair_kernel, CTA pairs cta_group::2, M=256" if "conv1|2|3" in plan.split("CTA pairs")[-1]
                                                    else "conv_strip_kernel, single CTA") + ", N=96, R=2)", "bound": "tensor",
                         "achieved": achieved_tf, "peak": peaks["bf16_tflops_burst"], "unit": "TFLOP/s",
                         "frac": achieved_tf / peaks["bf16_tflops_burst"], "traffic": traffic,
                         "peak_source": peaks["source"] + ": burst dense-bf16 figure (the kernel is timed inside a region of tens of ms); sustained: "
                                        f"{peaks['bf16_tflops']:.1f}",
                         "frac_of_sustained_peak": achieved_tf / peaks["bf16_tflops"],
                         "note": "achieved = useful (fp32-equivalent) FLOPs 2*25*32*96*P / CUDA-event time; the tensor pipe executes "
                                 f"{mma_per_chunk}x that (operand split), so the ceiling of this formulation is 1/{mma_per_chunk}",
                         "executed_frac": mma_per_chunk * achieved_tf / peaks["bf16_tflops_burst"]},
            "kernel_ms": ktimes,
            "kernel_ms_after_sustained_loop": ktimes_sus,
            "per_launch_roofline": per_launch,
            "reference_layer_roofline": {"ms": ref_roof, "frac": ref_roof / ms_step,
                                         "frac_sustained": None if ms_sus is None else ref_roof / ms_sus,
                                         "note": "sum over the reference's 10 convs of max(FLOPs / sustained dense-bf16 peak, fp32 activation bytes / copy "
                                                 "bandwidth) (SURVEY.md 8(d)) over the measured frame time; BASELINE target >= 0.70"},
            "whole_net_tflops_useful": FLOPS_PER_LR_PX * npx / (total_k * 1e-3) / 1e12 if total_k else None,
            "cpu_baseline": cpu,
            "e2e": {"value": world * OUT_MPIX / (e2e_rgb_ms * 1e-3), "unit": "Mpix/s", "ms_per_step": e2e_rgb_ms,
                    "h2d_bytes_per_step": int(H_LR * W_LR * 3), "d2h_bytes_per_step": int(H_LR * W_LR * 9 * 3),
                    "call": "rsr_submit_u8_rgb / rsr_wait (host u8 RGB -> host u8 RGB = data_to_img, pinned, 3 frames in flight)",
                    "frac_of_value": ms_step / e2e_rgb_ms,
                    "frac_of_sustained_value": None if ms_sus is None else ms_sus / e2e_rgb_ms,
                    "rgba": {"value": world * OUT_MPIX / (e2e_rgba_ms * 1e-3), "ms_per_step": e2e_rgba_ms, "d2h_bytes_per_step": int(H_LR * W_LR * 9 * 4),
                             "call": "rsr_submit_u8 / rsr_wait (RGBA out: data_to_img(..).to_rgba())"},
                    "unpipelined_ms_per_step": sync_ms_step,
                    "unpipelined_call": "rsr_upscale_u8, one synchronous call per frame",
                    "copy_probe": {"d2h_gbs_this_rank": d2h_gbs, "h2d_gbs_this_rank": h2d_gbs,
                                   "d2h_gbs_aggregate": sum(v[0] for v in d2h_all) / 1000.0,
                                   "d2h_floor_ms_per_step": (H_LR * W_LR * 9 * 3 / (d2h_gbs * 1e9) * 1e3) if d2h_gbs else None,
                                   "note": "pinned 74.6 MB buffers, every rank copying at once; d2h_floor = d2h_bytes_per_step / this rank's "
                                           "bandwidth: the result's way over PCIe bounds the e2e step once the forward pass is as fast as the copy"}},
            "config3": cfg3, "config4": cfg4, "config5": cfg5,
            "gpu_launches": launches[0],
            "clocks": clocks,
        }
        emit(line)
    up.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


_OUT = None


def emit(obj) -> None:
    """The one JSON line goes to the process's ORIGINAL stdout; everything else any library prints to fd 1
    (NCCL's version banner, torchrun chatter) has been diverted to stderr by main()."""
    (_OUT or sys.stdout).write(json.dumps(obj) + "\n")
    (_OUT or sys.stdout).flush()


def main():
    global _OUT
    sys.stdout.flush()
    _OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)          # from here on fd 1 is stderr: only emit() reaches the real stdout
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--plan", default="default", choices=["default", "fused", "split"],
                    help="launch plan of the middle levels (include/rsr_b200.h RSR_FLAG_*_MID); default = library default")
    ap.add_argument("--prec", default="default", choices=["default", "bf16x3"],
                    help="operand format of the 32-channel convolutions: default = fp16 + 2 x fp8 (2 MMAs per K chunk), bf16x3 = RSR_FLAG_BF16X3")
    ap.add_argument("--sustain-s", type=float, default=2.0, help="length of the sustained-regime loop in seconds (0 = skip)")
    ap.add_argument("--skip-configs", action="store_true", help="headline workload only (no BASELINE configs 3 / 4 / 5)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
