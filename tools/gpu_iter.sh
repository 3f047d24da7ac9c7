#!/bin/bash
# Development loop on the GPU box: parity summary, C3/C1 timing, optional ncu capture of one kernel.
# Every step runs under a short timeout: a hung kernel must never reach gpurun's own limit.
# usage: tools/gpu_iter.sh [tag] [kernel-regex]
TAG=${1:-x}; KRE=${2:-}
mkdir -p gpurun_out
timeout 150 python tools/gpu_check.py > gpurun_out/check_$TAG.log 2>&1 || { echo "gpu_check FAILED/TIMEOUT"; tail -3 gpurun_out/check_$TAG.log; exit 1; }
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
timeout 100 python tools/gpu_perf.py c3 1920 8 2 2>&1 | tail -2
timeout 60 python tools/gpu_perf.py c1 256 16 2 2>&1 | tail -1
if [ -n "$KRE" ]; then
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"$KRE" -s 2 -c 1 -o gpurun_out/prof_$TAG python tools/gpu_perf.py c3 1920 2 0 > gpurun_out/ncu_$TAG.log 2>&1
  tail -1 gpurun_out/ncu_$TAG.log
fi
