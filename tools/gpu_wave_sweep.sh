#!/bin/bash
# Development tool: suspension-knob sweep of the wavefront mode on a 1/8 tile of C3 (and the full frame).
# each argument: SUSPEND_BELOW:MIN_ROUNDS[:REFILL]
export YGL_MODE=wavefront
for CFG in "$@"; do
  IFS=: read B R F <<< "$CFG"
  export YGL_SUSPEND=$B YGL_SUSPEND_ROUNDS=$R
  if [ -n "$F" ]; then export YGL_REFILL=$F; else unset YGL_REFILL; fi
  A=$(TILE=0,8 timeout -s KILL 40 python tools/gpu_perf.py c3 1920 32 1 2>&1 | tail -1 | awk '{print $4, $5, $6, $(NF-3), $(NF-2), $(NF-1), $NF}')
  C=$(timeout -s KILL 40 python tools/gpu_perf.py c3 1920 8 1 2>&1 | tail -1 | awk '{print $4, $5}')
  echo "below $B rounds $R refill ${F:-8}: tile8 $A | full $C"
done
