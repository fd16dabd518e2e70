#!/bin/bash
# compute-sanitizer passes (memcheck, synccheck, racecheck) over small forward passes on every shipped launch plan:
# default (conv1|2|3 and the fused middle launch on CTA pairs, fp16 + mxfp8 operands), split plan, bf16x3 operands,
# factor 4, plus the generic conv.  Run under gpurun; logs land in gpurun_out/<SAN_TAG>_sanitizer_<tool>.log.
mkdir -p gpurun_out
cat > /tmp/san_case.py <<'PY'
import sys, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import rusty_sr_b200 as R
from conftest import synthetic_params
rng = np.random.default_rng(0)
img = rng.integers(0, 256, size=(37, 150, 3), dtype=np.uint8)
x = img.astype(np.float32) / 255.0
for name, flags in (("default", 0), ("split", R.RSR_FLAG_SPLIT_MID), ("bf16x3", R.RSR_FLAG_BF16X3), ("nopairs", R.RSR_FLAG_NO_PAIRS | R.RSR_FLAG_FUSED_MID)):
    with R.Upscaler(parameters="anime", flags=flags) as up:
        a = up.forward_u8(img); b = up.forward(x); c = up.forward_u8(img, rgb=True)
        d = up.forward_u8(np.concatenate([img, np.full(img.shape[:2] + (1,), 255, np.uint8)], axis=2))   # RGBA in
        assert np.array_equal(a, d)
        print("ok", name, up.plan(), a.shape, b.shape, c.shape, flush=True)
p4 = synthetic_params(R.load_builtin_params("imagenet"), 4)
with R.Upscaler(p4, factor=4) as up:
    print("ok factor 4", up.forward(x[:20, :140]).shape, flush=True)
xx = rng.standard_normal((9, 140, 64), dtype=np.float32); w = rng.standard_normal((64, 3, 3, 64), dtype=np.float32) * 0.05
o, ms = R.conv2d_3x3(xx, w)
print("ok conv2d", o.shape)
PY
for tool in ${SAN_TOOLS:-memcheck synccheck racecheck}; do
  timeout ${SAN_TIMEOUT:-600} compute-sanitizer --tool $tool --error-exitcode 7 python /tmp/san_case.py > gpurun_out/${SAN_TAG:-r03}_sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|^ok |Error|error|hazard" gpurun_out/${SAN_TAG:-r03}_sanitizer_$tool.log | head -12
done
