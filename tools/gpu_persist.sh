#!/bin/bash
# Development loop for the persistent mode: parity summary + timings (full frame and a 1/8 tile), both modes.
# Every step runs under a short timeout: a hung kernel must never reach gpurun's own limit.
mkdir -p gpurun_out
export YGL_WATCHDOG_S=5
summ() {
python - <<PY
import json
d = json.load(open("gpurun_out/gpu_check.json"))
bad = 0
for k, v in d.items():
    if k.startswith("rays_"):
        bad += v["mismatch"] + v.get("instance_mismatch", 0) + v["any_mismatch"]
    else:
        bad += sum(r["vs_ref"]["frac_exact"] != 1.0 for r in v)
print("PARITY(bit-exact vs unmodified reference)", "OK" if bad == 0 else f"FAIL({bad})", {k: [r["vs_ref"]["frac_exact"] for r in v] for k, v in d.items() if k.startswith("render_")})
PY
}
for MODE in persistent wavefront; do
  echo "== $MODE"
  YGL_MODE=$MODE timeout -s KILL 150 python tools/gpu_check.py > gpurun_out/check_$MODE.log 2>&1 && summ || { echo "gpu_check FAILED/TIMEOUT"; tail -5 gpurun_out/check_$MODE.log; }
  YGL_MODE=$MODE timeout -s KILL 60 python tools/gpu_perf.py c3 1920 8 2 2>&1 | tail -2
  YGL_MODE=$MODE TILE=0,8 timeout -s KILL 60 python tools/gpu_perf.py c3 1920 32 2 2>&1 | tail -2
  YGL_MODE=$MODE timeout -s KILL 60 python tools/gpu_perf.py c1 256 16 2 2>&1 | tail -1
done
