#!/bin/bash
# Development tool: one gpurun call = several bounded steps, every one under its own timeout so that a hung kernel can
# never reach gpurun's limit. Results land in gpurun_out/<tag>_*.  usage: tools/gpu_session.sh <tag> <step> [<step> ...]
# steps: tests | tests_fast | bench | ab | launches | ncu | scale2
TAG=$1; shift
mkdir -p gpurun_out
PERF="python tools/gpu_perf.py"
for STEP in "$@"; do
  echo "=== $STEP"
  case $STEP in
    tests)
      timeout -s KILL 1500 python -m pytest tests -m gpu -q --maxfail=12 --durations=12 -s > gpurun_out/${TAG}_tests.log 2>&1; echo "pytest rc=$?"; grep -a "bvh build\|passed\|failed\|^FAILED\|^ERROR" gpurun_out/${TAG}_tests.log | tail -n 40 | cut -c1-220 ;;
    tests_new)   # the parity core on a variant build (YGL_B200_LIB is read by the python binding)
      YGL_B200_LIB=$PWD/yocto-gl_b200/${NEWLIB:-lib_new}/libygl_b200.so timeout -s KILL 900 python -m pytest tests -m gpu -q --maxfail=5 -k "${NEWTESTS:-render_matches_oracle or full_size or full_state or options or tiles_equal or counters or deep or intersect}" > gpurun_out/${TAG}_tests_new.log 2>&1; echo "pytest(new) rc=$?"; tail -n 4 gpurun_out/${TAG}_tests_new.log | cut -c1-200 ;;
    tests_fast)
      timeout -s KILL 900 python -m pytest tests -m gpu -q --maxfail=8 -k "not full_size and not libm and not counters" > gpurun_out/${TAG}_tests.log 2>&1; echo "pytest rc=$?"; tail -n 25 gpurun_out/${TAG}_tests.log | cut -c1-220 ;;
    bench)
      timeout -s KILL 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; tail -c 2500 gpurun_out/${TAG}_bench.json ;;
    ab)
      for LIB in lib lib_cls2; do
        [ -f yocto-gl_b200/$LIB/libygl_b200.so ] || continue
        for OPTS in "" "bin=0"; do
          echo "--- $LIB $OPTS"
          YGL_B200_LIB=$PWD/yocto-gl_b200/$LIB/libygl_b200.so timeout -s KILL 120 $PERF c3 1920 16 2 profile=1 $OPTS 2>&1 | tail -n 2 | cut -c1-260
          YGL_B200_LIB=$PWD/yocto-gl_b200/$LIB/libygl_b200.so timeout -s KILL 120 $PERF c3 1920 32 2 tile=0,8 profile=1 $OPTS 2>&1 | tail -n 1 | cut -c1-260
        done
      done
      timeout -s KILL 120 $PERF c1 256 16 2 2>&1 | tail -n 1 | cut -c1-200
      timeout -s KILL 120 $PERF c2 1280 16 2 2>&1 | tail -n 1 | cut -c1-200
      timeout -s KILL 120 $PERF c5 1920 8 2 2>&1 | tail -n 1 | cut -c1-200 ;;
    san)     # memcheck over the three stack variants of the extend kernel and the shading kernels, small sizes
      timeout -s KILL 500 compute-sanitizer --tool memcheck --print-limit 3 python -m pytest tests/test_gpu_parity.py -q -x -k "${SANTESTS:-(render_matches_oracle and (cornell or instanced4 or hair)) or deep_tree or refitted_bvh and instanced4}" > gpurun_out/${TAG}_san.log 2>&1; echo "sanitizer rc=$?"; grep -a "ERROR SUMMARY\|passed\|failed" gpurun_out/${TAG}_san.log | tail -n 4
      timeout -s KILL 300 compute-sanitizer --tool memcheck --print-limit 3 $PERF c3 256 2 0 > gpurun_out/${TAG}_san_c3.log 2>&1; echo "sanitizer c3 rc=$?"; grep -a "ERROR SUMMARY\|Msamples" gpurun_out/${TAG}_san_c3.log | tail -n 3 | cut -c1-200 ;;
    ab2)
      IFS=';' read -ra CFGS <<< "${AB_CFGS:-lib;lib_new}"
      for CFG in "${CFGS[@]}"; do
        set -- $CFG; LIB=$1; shift
        [ -f yocto-gl_b200/$LIB/libygl_b200.so ] || continue
        echo "--- $LIB $*"
        YGL_B200_LIB=$PWD/yocto-gl_b200/$LIB/libygl_b200.so timeout -s KILL 120 $PERF c3 1920 16 2 profile=1 "$@" 2>&1 | tail -n 1 | cut -c1-260
        YGL_B200_LIB=$PWD/yocto-gl_b200/$LIB/libygl_b200.so timeout -s KILL 120 $PERF c3 1920 32 2 tile=0,8 profile=1 "$@" 2>&1 | tail -n 1 | cut -c1-260
        YGL_B200_LIB=$PWD/yocto-gl_b200/$LIB/libygl_b200.so timeout -s KILL 120 $PERF c2 1280 16 2 "$@" 2>&1 | tail -n 1 | cut -c1-200
        YGL_B200_LIB=$PWD/yocto-gl_b200/$LIB/libygl_b200.so timeout -s KILL 120 $PERF c5 1920 8 2 "$@" 2>&1 | tail -n 1 | cut -c1-200
      done ;;
    scale)   # usage: gpurun --gpus N -- tools/gpu_session.sh TAG scale   (N from nvidia-smi)
      N=$(nvidia-smi -L | wc -l)
      for G in ${SCALE_N:-1 2 4 8}; do
        [ $G -le $N ] || continue
        if [ $G -eq 1 ]; then CMD="python bench.py --no-cpu-baseline"; else CMD="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port $((29500 + G)) bench.py --gpus $G"; fi
        timeout -s KILL 420 $CMD --steps 8 --warmup 3 > gpurun_out/${TAG}_scale_n$G.json 2> gpurun_out/${TAG}_scale_n$G.err; echo "n=$G rc=$?"
        python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_scale_n$G.json").read().strip().splitlines()[-1])
    print("  value %.1f e2e %.1f ms/step %.1f frac %.3f share %.2f launches %d" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["kernel_share_of_step"], d["gpu_launches"]))
except Exception as e:
    print("  no line:", e)
PY
      done ;;
    final)   # the round's measurement set: bench line, other configs, launch list, ncu captures (tools/make_profiles.py)
      timeout -s KILL 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; tail -c 600 gpurun_out/${TAG}_bench.json
      for W in c1 c2 c5; do timeout -s KILL 300 python bench.py --workload $W --no-cpu-baseline > gpurun_out/${TAG}_bench_$W.json 2> gpurun_out/${TAG}_bench_$W.err; echo "bench $W rc=$?"; done
      timeout -s KILL 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_bench_ncu.log 2>&1; echo "launches rc=$?"
      timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:'k_extend|k_shade|k_lightpdf|k_finish' -s 60 -c 8 -o gpurun_out/${TAG}_kernels $PERF c3 1920 4 0 > gpurun_out/${TAG}_ncu.log 2>&1; echo "ncu rc=$?"
      timeout -s KILL 300 ncu --set full --clock-control none -k regex:'k_extend' -c 1 -o gpurun_out/${TAG}_extend_first $PERF c3 1920 1 0 > gpurun_out/${TAG}_ncu_first.log 2>&1; echo "ncu first rc=$?" 
      # gpurun brings back at most 64 MiB: compress the captures, and drop the big one rather than lose everything
      gzip -f gpurun_out/${TAG}_*.ncu-rep; du -sm gpurun_out | cut -f1
      if [ "$(du -sm gpurun_out | cut -f1)" -gt 58 ]; then rm -f gpurun_out/${TAG}_kernels.ncu-rep.gz; echo "dropped the kernels capture (too large)"; fi ;;
    launches)
      timeout -s KILL 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_bench_ncu.log 2>&1; echo "launches rc=$?"
      python tools/ncu_summary.py --launches gpurun_out/${TAG}_launches.csv | head -n 20 ;;
    sweep)   # SWEEP="c3 1920 32 2 tile=0,8 profile=1 refill=4;..." : one gpu_perf.py line per entry
      IFS=';' read -ra RUNS <<< "$SWEEP"
      for RUN in "${RUNS[@]}"; do
        timeout -s KILL 150 $PERF $RUN 2>&1 | tail -n 1 | cut -c1-250
      done ;;
    ncu_ext)   # two mid-frame launches of the extend kernel only, with source correlation
      timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:'k_extend' -s 3 -c 2 -o gpurun_out/${TAG}_extend $PERF c3 1920 2 0 > gpurun_out/${TAG}_ncu.log 2>&1; echo "ncu rc=$?"
      gzip -f gpurun_out/${TAG}_extend.ncu-rep ;;
    ncu)
      timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:'k_extend|k_shade|k_lightpdf|k_finish' -s 40 -c 8 -o gpurun_out/${TAG}_kernels $PERF c3 1920 2 0 > gpurun_out/${TAG}_ncu.log 2>&1; echo "ncu rc=$?"; tail -n 2 gpurun_out/${TAG}_ncu.log | cut -c1-200 ;;
  esac
done
