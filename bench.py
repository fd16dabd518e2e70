#!/usr/bin/env python
"""bench.py — throughput of the rusty_sr forward path on B200 (one JSON line on stdout).

    python bench.py --gpus N --steps K --warmup W            # our sm_100a path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port)

Metric (BASELINE.json): output Mpix/s of the super-resolution forward pass.  Workload at every N:
one 1920x1080 RGB frame per GPU per step -> 5760x3240 (factor 3: the reference is compiled with
FACTOR = 3 and ships factor-3 weights only, main.rs:31), bundled `imagenet` weights,
uniform-random u8 input (numpy default_rng(1234 + rank)).  Image batches shard across GPUs with no
data-path collective, so scaling is "weak": every rank upscales its own frame.

  value   : frames resident in HBM (f32 [H][W][3] in, f32 [3H][3W][3] out: the graph.forward
            contract, main.rs:171), timed with CUDA events on the launching stream, max over ranks.
  e2e     : the same metric through the reference-facing host call `rsr_upscale_u8` (what the
            reference's upscale() does around forward: img_to_data -> forward -> data_to_img,
            main.rs:170-175) with pinned HOST buffers; H2D and D2H inside the timed region.
  roofline: the dominant launch (conv1|2|3, 59 % of the FLOPs), algorithmic FLOPs / its CUDA-event
            duration, against the measured dense-bf16 peak of MEASURED_PEAKS.json.
  cpu_baseline: the CPU oracle (port of the reference path; the Rust reference cannot be built
            here) on the host cores, on a bounded sample.
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
    "pack_input": (0, 12 + 80),
    "conv0": (4800, 64 + 128),
    "conv1|2|3": (153600, 128 + 384),
    "conv5|6|7": (18432 * 2 + 15552, 384 + 384),
    "conv8|9": (18432 + 15552, 384 + 256),
    "conv5|6|7+conv8|9 (fused)": (18432 * 3 + 15552 * 2, 384 + 256),
    "conv10+expand+skip": (15552, 256 + 16 + 108),
}
WORKLOAD = f"{W_LR}x{H_LR} RGB -> {W_LR*FACTOR}x{H_LR*FACTOR} (factor 3, bundled imagenet weights), 1 frame/GPU/step"
N_ROTATE = 6   # distinct input frames cycled through so the input is never L2 resident


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            d = json.load(fh)
        return {"bf16_tflops": float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1400.0))),
                "bf16_tflops_burst": float(d.get("bf16_tflops", 1590.0)),
                "hbm_gbs": float(d.get("hbm_gbs", 6650.0)), "source": "measured (MEASURED_PEAKS.json, sustained)"}
    return {"bf16_tflops": 1400.0, "bf16_tflops_burst": 1590.0, "hbm_gbs": 6650.0,
            "source": "fallback (B200_PROFILING.md)"}


def per_launch_roofline(ktimes, npx, peaks):
    """Every launch against min(tensor, HBM): roofline_ms = max(useful FLOPs / measured sustained dense-bf16 peak, algorithmic
    bytes / measured copy bandwidth), frac = roofline_ms / measured ms.  *_bf16x3: the same with the FLOPs the BF16x3 split
    actually executes (3 x useful) against the measured burst peak — the ceiling of this formulation."""
    per_launch = {}
    for k, ms in ktimes.items():
        if k not in LAUNCH_WORK or not ms or ms != ms:
            continue
        fl, by = LAUNCH_WORK[k]
        t_tensor = fl * npx / (peaks["bf16_tflops"] * 1e12) * 1e3
        t_hbm = by * npx / (peaks["hbm_gbs"] * 1e9) * 1e3
        t_x3 = 3 * fl * npx / (peaks["bf16_tflops_burst"] * 1e12) * 1e3   # the three split products, at the burst peak
        per_launch[k] = {"ms": ms, "useful_tflops": fl * npx / (ms * 1e-3) / 1e12, "hbm_gbs": by * npx / (ms * 1e-3) / 1e9,
                         "bound": "tensor" if t_tensor > t_hbm else "hbm", "roofline_ms": max(t_tensor, t_hbm),
                         "frac": max(t_tensor, t_hbm) / ms,
                         "roofline_ms_bf16x3": max(t_x3, t_hbm), "frac_bf16x3": max(t_x3, t_hbm) / ms}
    return per_launch


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[0])); smax.append(float(r[1])); power.append(float(r[2]))
            except ValueError:
                continue
            for name, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        thr = 0.6 * max(power) if power else 0.0
        busy = [c for c, p in zip(sm, power) if p >= thr] or sm
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "samples_under_load": len(busy),
                "sampled_over": "whole bench run (warm-up, timed device loop, per-kernel loop, e2e loop), 100 ms period",
                "reasons": sorted(reasons)}


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
def run_ours(args):
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
    peaks = load_peaks()

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
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

    def step(i):
        ptr = d_in.data_ptr() + (i % N_ROTATE) * frame_elems * 4
        up.forward_device(ptr, 1, W_LR, H_LR, d_out_f32=d_out.data_ptr())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- device-resident throughput ("value") -------------------------------------------------
    with torch.cuda.stream(stream):
        for i in range(args.warmup):
            step(i)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ktimes = {}
        launches = 0
        barrier()
        e0.record(stream)
        for i in range(args.steps):
            step(i)
            launches += up.last_launch_count()
        e1.record(stream)
        barrier()
        ms_total = e0.elapsed_time(e1)
        # per-kernel durations (CUDA events the library records on the same stream around every launch)
        kt_runs = []
        for i in range(min(args.steps, 10)):
            step(i)
            kt_runs.append(up.last_kernel_times_ms())
        for k in kt_runs[0]:
            ktimes[k] = statistics.mean(r[k] for r in kt_runs)
        fused = up.last_launch_count() == 5      # conv5|6|7 and conv8|9 run as one launch (conv_fused.cuh)
        plan = up.plan()
        if fused:
            ktimes["conv5|6|7+conv8|9 (fused)"] = ktimes.pop("conv5|6|7") + ktimes.pop("conv8|9")

    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = world * OUT_MPIX / (ms_step * 1e-3)

    # ---- end to end through the host call, pinned buffers ---------------------------------------
    # The call a user makes for a stream of frames: rsr_submit_u8 / rsr_wait (3 frames in flight, so the
    # H2D of frame k+1 and the D2H of frame k-1 overlap the forward pass of frame k).  Every step moves
    # its input from pinned host memory to the device and its result back to pinned host memory.
    DEPTH = 3
    pin_in = torch.from_numpy(imgs_u8).pin_memory()
    pin_out = torch.empty((DEPTH, H_LR * FACTOR, W_LR * FACTOR, 4), dtype=torch.uint8).pin_memory()
    in_np, out_np = pin_in.numpy(), pin_out.numpy()

    def e2e_run(n):
        tickets = []
        for i in range(n):
            if i >= DEPTH:
                up.wait(tickets[i - DEPTH])      # the output buffer about to be reused has landed
            tickets.append(up.submit_u8(in_np[i % N_ROTATE], out_np[i % DEPTH]))
        up.wait_all()

    e2e_run(max(args.warmup, DEPTH))
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record(stream)
    t_wall = time.perf_counter()
    e2e_run(args.steps)              # returns after the last D2H copy has completed
    t_wall = time.perf_counter() - t_wall
    e3.record(stream)
    barrier()
    # device-side bracket (events on the compute stream) and host wall clock agree because the region
    # ends with a host-synchronous wait on the last D2H; report the larger of the two
    t2 = torch.tensor([max(e2.elapsed_time(e3), t_wall * 1e3)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    e2e_ms_step = float(t2.item()) / args.steps
    e2e_value = world * OUT_MPIX / (e2e_ms_step * 1e-3)
    # the plain synchronous call, one frame at a time, for reference
    import ctypes
    L = R._lib.load()
    t_sync = time.perf_counter()
    for i in range(min(args.steps, 10)):
        R._lib.check(L.rsr_upscale_u8(up._h, ctypes.c_void_p(in_np[i % N_ROTATE].ctypes.data), W_LR, H_LR, 3,
                                      ctypes.c_void_p(out_np[0].ctypes.data)))
    sync_ms_step = (time.perf_counter() - t_sync) * 1e3 / min(args.steps, 10)
    clocks = sampler.stop() if sampler else None

    if rank == 0:
        npx = W_LR * H_LR
        k_ms = ktimes.get("conv1|2|3", float("nan"))
        achieved_tf = FLOPS_CONV123_PER_LR_PX * npx / (k_ms * 1e-3) / 1e12
        traffic = None
        prof = os.path.join(ROOT, "profiles", "ncu_summary.json")
        if os.path.exists(prof):
            try:
                traffic = json.load(open(prof)).get("conv1|2|3", {}).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        cpu_mpix, cpu_times, cpu_threads = (None, None, None)
        cpu = None
        if world == 1:
            rows = 540
            cpu_mpix, cpu_times, cpu_threads = cpu_sample(rows, 2, 1)
            cpu = {"value": cpu_mpix * len(cpu_times) / sum(cpu_times), "unit": "Mpix/s", "cores": cpu_threads,
                   "kind": "port",
                   "sample": f"2 x ({W_LR}x{rows} slab of the same frame), img_to_data+forward(fp32)+data_to_img",
                   "note": "CPU restatement of rusty_sr/alumina path (Rust toolchain unavailable)"}
        total_k = sum(ktimes.values())
        per_launch = per_launch_roofline(ktimes, npx, peaks)
        line = {
            "metric": "output Mpix/s (3x upscale, sr_net forward)", "value": value, "unit": "Mpix/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16x3 split operands, f32 accumulate (fp32-equivalent)", "data": "synthetic",
            "config": {"workload": WORKLOAD, "factor": FACTOR, "weights": "imagenet.rsr",
                       "parallelism": f"image-batch sharding x{world}, no collective",
                       "launch_plan": plan,
                       "l2": f"working set ~1.3 GB/step >> 126 MB L2; input rotates over {N_ROTATE} distinct frames (149 MB)",
                       "io": "value: f32 in/out resident in HBM; e2e: u8 RGB host -> u8 RGBA host (pinned)"},
            "roofline": {"kernel": "conv1|2|3 (" + ("conv_pair_kernel, CTA pairs cta_group::2, M=256" if "conv1|2|3" in plan.split("CTA pairs")[-1]
                                                    else "conv_strip_kernel, single CTA") + ", N=96, R=2)", "bound": "tensor",
                         "achieved": achieved_tf, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                         "frac": achieved_tf / peaks["bf16_tflops"], "traffic": traffic,
                         "peak_source": peaks["source"],
                         "note": "achieved = useful (fp32-equivalent) FLOPs 2*25*32*96*P / CUDA-event time; the "
                                 "tensor pipe executes 3x that (BF16x3 split), so the ceiling of this formulation is 1/3",
                         "executed_frac": 3 * achieved_tf / peaks["bf16_tflops"]},
            "kernel_ms": ktimes,
            "per_launch_roofline": per_launch,
            "whole_net_tflops_useful": FLOPS_PER_LR_PX * npx / (total_k * 1e-3) / 1e12 if total_k else None,
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": "Mpix/s", "ms_per_step": e2e_ms_step,
                    "h2d_bytes_per_step": int(H_LR * W_LR * 3), "d2h_bytes_per_step": int(H_LR * W_LR * 9 * 4),
                    "call": "rsr_submit_u8 / rsr_wait (host u8 RGB -> host u8 RGBA, pinned, 3 frames in flight)",
                    "unpipelined_ms_per_step": sync_ms_step,
                    "unpipelined_call": "rsr_upscale_u8, one synchronous call per frame"},
            "gpu_launches": launches,
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
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
