#!/usr/bin/env python
"""CPU numerics probe for the operand splits the tensor-core path can use (no GPU needed).

The forward pass (reference src/network.rs:27-72) is evaluated in fp64 with every convolution operand rounded the way a
candidate tcgen05 formulation would store it, and compared with the exact fp64 result:

  bf16x3      a_hi*w_hi + a_mid*w_hi + a_hi*w_mid, all bf16 (round 1 default; 3 kind::f16 MMAs per 16-channel chunk)
  f16+e4m3x2  a_h*w_h in fp16 (11-bit significands) + 2^-11 * (e4m3(2^11 a_l) * e4m3(w) + e4m3(a) * e4m3(2^11 w_l)):
              one kind::f16 MMA (K = 16) + ONE fp8 MMA (K = 32: both cross terms concatenated along K) per chunk = cost 2.0
  f16+e5m2x2  the same with e5m2 cross operands
  f16+alo     fp16 main product + only the a_l * w cross term (cost 1.5)
  f16         single fp16 product (cost 1.0)

conv0 always keeps bf16x3 on the x-im2col'd input (HBM-bound launch, K = 75).  Accumulation is exact here (fp64); the tensor
core's fp32 accumulation adds ~1e-6.  Usage: python tools/precision_probe.py [--size 256] [--json out.json]
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def rnd(t, dtype):
    return t.to(torch.float32).to(dtype).to(torch.float64)


def split_terms(a, w, scheme):
    """Returns a list of (a_operand, w_operand, scale) whose scaled products sum to the emulated a*w."""
    if scheme == "exact":
        return [(a, w, 1.0)]
    if scheme == "bf16x3":
        ah, wh = rnd(a, torch.bfloat16), rnd(w, torch.bfloat16)
        am, wm = rnd(a - ah, torch.bfloat16), rnd(w - wh, torch.bfloat16)
        return [(ah, wh, 1.0), (am, wh, 1.0), (ah, wm, 1.0)]
    if scheme == "f16":
        return [(rnd(a, torch.float16), rnd(w, torch.float16), 1.0)]
    if scheme.startswith("f16+"):
        f8 = torch.float8_e5m2 if "e5m2" in scheme else torch.float8_e4m3fn
        S = 2.0 ** 11
        ah, wh = rnd(a, torch.float16), rnd(w, torch.float16)
        al, wl = a - ah, w - wh
        terms = [(ah, wh, 1.0), (rnd(al * S, f8), rnd(w, f8), 1.0 / S)]
        if not scheme.endswith("alo"):
            terms.append((rnd(a, f8), rnd(wl * S, f8), 1.0 / S))
        return terms
    raise ValueError(scheme)


def forward(params, x_hwc3, scheme, factor=3):
    import rusty_sr_b200 as R
    p = {}
    for s in R.param_layout():
        n = int(np.prod(s.shape))
        p[s.name] = torch.from_numpy(params[s.offset:s.offset + n].astype(np.float64).reshape(s.shape))

    def conv(t, name, sch=None):
        w = p[name].permute(0, 3, 1, 2)
        out = 0
        for a_, w_, sc in split_terms(t, w, sch or scheme):
            out = out + sc * F.conv2d(a_, w_, padding=w.shape[2] // 2)
        return out

    # activations are stored in fp32 by the epilogue before being split: round them
    f32 = lambda t: t.to(torch.float32).to(torch.float64)
    belu = lambda t, name: f32(p[name].view(1, -1, 1, 1) * t + torch.sqrt(t * t + 1.0) - 1.0)
    bias = lambda name: p[name].view(1, -1, 1, 1)
    x = torch.from_numpy(x_hwc3.astype(np.float64)).permute(2, 0, 1)[None]
    c0 = "exact" if scheme == "exact" else "bf16x3"
    f = belu(conv(x, "conv0", c0) + bias("f_bias"), "f_activ")
    l1 = belu(conv(f, "conv1") + bias("l1_bias"), "l1_activ")
    l2 = belu(conv(f, "conv2") + conv(l1, "conv5") + bias("l2_bias"), "l2_activ")
    l3 = belu(conv(f, "conv3") + conv(l1, "conv6") + conv(l2, "conv8") + bias("l3_bias"), "l3_activ")
    e = conv(l1, "conv7") + conv(l2, "conv9") + conv(l3, "conv10") + bias("expand_bias")
    n, _, H, W = e.shape
    e = e.view(n, factor * factor, 3, H, W).permute(0, 2, 1, 3, 4).reshape(n, 3 * factor * factor, H, W)
    out = F.pixel_shuffle(e, factor) + F.interpolate(x, scale_factor=factor, mode="bilinear", align_corners=False)
    acts = {"f": f, "l1": l1, "l2": l2, "l3": l3}
    return out[0].permute(1, 2, 0).numpy(), {k: float(v.abs().max()) for k, v in acts.items()}


def srgb_to_linear(v):
    v = np.clip(v, 0.0, 1.0)
    return np.where(v <= 0.04045, v / 12.92, ((v + 0.055) / 1.055) ** 2.4)


def main():
    from conftest import load_png, noise_image
    from oracle import oracle as O
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--json", default=None)
    ap.add_argument("--schemes", default="bf16x3,f16+e4m3x2,f16+e5m2x2,f16+alo,f16")
    args = ap.parse_args()
    torch.set_num_threads(os.cpu_count() or 4)
    weights = {}
    for k in ("imagenet", "imagenetlinear", "anime"):
        with open(os.path.join(ROOT, "rusty_sr_b200", "res", k + ".rsr"), "rb") as fh:
            weights[k] = O.decode_rsr(fh.read())
    inputs = {
        f"noise{args.size}": noise_image(args.size, args.size, 1234),
        "cartoon_lr": O.img_to_data(load_png("cartoon_lr")),
        "butterfly_lr": O.img_to_data(load_png("butterfly_lr")),
        "const0.5": np.full((32, 32, 3), 0.5, np.float32),
        "out_of_range": (noise_image(48, 48, 7) * 3.0 - 1.0).astype(np.float32),
    }
    rows = []
    for wname, prm in weights.items():
        for iname, x in inputs.items():
            ref, amax = forward(prm, x, "exact")
            for sch in args.schemes.split(","):
                got, _ = forward(prm, x, sch)
                raw = float(np.abs(got - ref).max())
                lin = float(np.abs(srgb_to_linear(got) - srgb_to_linear(ref)).max())
                rows.append({"weights": wname, "input": iname, "scheme": sch, "max_abs_srgb": raw, "max_abs_linear": lin,
                             "act_max": amax})
                print(f"{wname:15s} {iname:13s} {sch:11s} sRGB {raw:.2e}  linear {lin:.2e}   act max {max(amax.values()):.1f}", flush=True)
    worst = {}
    for r in rows:
        w = worst.setdefault(r["scheme"], {"max_abs_srgb": 0.0, "max_abs_linear": 0.0})
        w["max_abs_srgb"] = max(w["max_abs_srgb"], r["max_abs_srgb"])
        w["max_abs_linear"] = max(w["max_abs_linear"], r["max_abs_linear"])
    print(json.dumps(worst, indent=1))
    if args.json:
        with open(args.json, "w") as fh:
            json.dump({"rows": rows, "worst": worst}, fh, indent=1)


if __name__ == "__main__":
    main()
