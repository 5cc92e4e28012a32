#!/bin/bash
# Round-2 evidence bundle (run on the GPU box through gpurun): bench lines first (no profiler), then the ncu launch list.
set -x
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 3 2> gpurun_out/r02_bench.err | grep "^{" > gpurun_out/r02_bench.json
timeout 600 python bench.py --steps 200 --warmup 3 --no-extras 2>/dev/null | grep "^{" > gpurun_out/r02_bench_200steps.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | grep "^{" > gpurun_out/r02_bench_reference.json
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off \
    --csv --log-file gpurun_out/r02_launches.csv python tools/prof_decode.py > gpurun_out/r02_ncu_launches.log 2>&1
tail -2 gpurun_out/r02_ncu_launches.log
wc -l gpurun_out/r02_launches.csv
