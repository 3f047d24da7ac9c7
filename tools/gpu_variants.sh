#!/bin/bash
# Development tool (GPU box): rebuild the library with different -D flags and time C3.
for mb in "$@"; do
  touch yocto-gl_b200/csrc/ygl_kernels.cu
  make -C yocto-gl_b200/csrc EXTRA="-DYGL_EXT_MINBLOCKS=$mb" 2>&1 | grep -E "error" 
  regs=$(grep -A2 "k_extendILb0" yocto-gl_b200/lib/obj/ygl_kernels.ptxas.log | grep -o "Used [0-9]* registers")
  echo "minblocks=$mb $regs $(PROFILE=1 python tools/gpu_perf.py c3 1920 32 1 | tail -1 | grep -o "extend_ms.: [0-9.]*")"
done
