#!/bin/bash
# Development tool (GPU box): time C3 with prebuilt library variants yocto-gl_b200/lib/variant_*.so
for lib in "$@"; do
  echo "$lib: $(YGL_B200_LIB=$PWD/yocto-gl_b200/lib/$lib PROFILE=1 python tools/gpu_perf.py c3 1920 32 1 | tail -1 | grep -o "Msamples/s\|[0-9.]* Msamples\|extend_ms.: [0-9.]*\|loop_ms.: [0-9.]*" | tr '\n' ' ')"
done
