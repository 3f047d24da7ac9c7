#!/bin/bash
# Round-end measurement set on one GPU: bench line, ncu launch list of the same command, one full capture of
# the dominant kernels. Every step under a timeout. Budget note: the launch-list pass replays every kernel under ncu;
# keep it short (--steps 1 --warmup 1, no CPU baseline / e2e legs, first 200 launches) - the first version of this
# script took 11 GPU-minutes.
mkdir -p gpurun_out
timeout -s KILL 400 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -c 1800 gpurun_out/bench_final.json
timeout -s KILL 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_ncu.log 2>&1
tail -2 gpurun_out/bench_ncu.log | cut -c1-200
timeout -s KILL 200 ncu --set full --clock-control none --import-source on -k regex:'k_extend|k_shade|k_lightpdf' -s 9 -c 3 -o gpurun_out/r01_final_kernels python tools/gpu_perf.py c3 1920 2 0 > gpurun_out/ncu_final.log 2>&1
tail -2 gpurun_out/ncu_final.log | cut -c1-200
