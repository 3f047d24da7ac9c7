#!/bin/bash
# Development tool: A/B timing of builds of the library (default / YGL_OUTLINE / experimental) and knobs.
# each line: LIB MODE [ENV=VALUE ...]
export YGL_WATCHDOG_S=5
run() {
  local LIB=$1 MODE=$2; shift 2
  A=$(env "$@" YGL_B200_LIB=$PWD/yocto-gl_b200/$LIB/libygl_b200.so YGL_MODE=$MODE timeout -s KILL 40 python tools/gpu_perf.py c3 1920 8 1 2>&1 | tail -1 | awk '{print $4, $5, $6, $(NF-3), $(NF-2)}')
  B=$(env "$@" YGL_B200_LIB=$PWD/yocto-gl_b200/$LIB/libygl_b200.so YGL_MODE=$MODE TILE=0,8 timeout -s KILL 40 python tools/gpu_perf.py c3 1920 32 1 2>&1 | tail -1 | awk '{print $4, $5, $6, $(NF-3), $(NF-2)}')
  C=$(env "$@" YGL_B200_LIB=$PWD/yocto-gl_b200/$LIB/libygl_b200.so YGL_MODE=$MODE timeout -s KILL 40 python tools/gpu_perf.py c1 256 16 1 2>&1 | tail -1 | awk '{print $4, $5, $6}')
  echo "$LIB $MODE $*: c3 $A | tile8 $B | c1 $C"
}
run lib wavefront
run lib wavefront YGL_LONE_STEPS=100
run lib wavefront YGL_LONE_STEPS=200
run lib wavefront YGL_LONE_STEPS=400
run lib wavefront YGL_LONE_STEPS=800
run lib wavefront YGL_LONE=24 YGL_LONE_STEPS=200
