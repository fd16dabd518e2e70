#!/bin/bash
# compute-sanitizer passes over one small forward pass + the generic conv (run under gpurun).
mkdir -p gpurun_out
cat > /tmp/san_case.py <<'PY'
import sys, numpy as np
sys.path.insert(0, ".")
import rusty_sr_b200 as R
rng = np.random.default_rng(0)
img = rng.integers(0, 256, size=(37, 150, 3), dtype=np.uint8)
with R.Upscaler(parameters="anime") as up:
    a = up.forward_u8(img)
    b = up.forward(img.astype(np.float32) / 255.0)
x = rng.standard_normal((9, 140, 64), dtype=np.float32); w = rng.standard_normal((64, 3, 3, 64), dtype=np.float32) * 0.05
o, ms = R.conv2d_3x3(x, w)
print("ok", a.shape, b.shape, o.shape)
PY
for tool in memcheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 7 python /tmp/san_case.py > gpurun_out/sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|ok \(|Error|error" gpurun_out/sanitizer_$tool.log | head -8
done
