#!/bin/bash
# Development tool: layout sweep of the persistent mode on C3 (full frame and a 1/8 tile).
# each argument: EXT_SMS:LPDF_WARPS[:ROLES]
export YGL_WATCHDOG_S=5 YGL_MODE=persistent
for CFG in "$@"; do
  IFS=: read E L R <<< "$CFG"
  echo "== ext_sms $E lpdf_warps $L roles ${R:-default}"
  export YGL_PERSIST_EXT_SMS=$E YGL_PERSIST_LPDF_WARPS=$L
  if [ -n "$R" ]; then export YGL_PERSIST_ROLES=$R; else unset YGL_PERSIST_ROLES; fi
  timeout -s KILL 40 python tools/gpu_perf.py c3 1920 8 1 2>&1 | tail -1
  TILE=0,8 timeout -s KILL 40 python tools/gpu_perf.py c3 1920 32 1 2>&1 | tail -1
done
