#!/bin/bash
# Run on the GPU box (gpurun): launch list of one bench run + full ncu captures of every hot kernel.
# Each ncu command follows a plain run of the same program that exited 0 (numbers under ncu are never bench values).
set -e
TAG=${1:-r01b}
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_step_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu1.log 2>&1
python tools/run_pre.py > /dev/null
ncu --set full --clock-control none --import-source on \
    -k regex:"loss_march|bracket_sample|resize_march|percentile_from_brackets|normalize_stats|metrics_sample|depth_extract|median_scale|metrics_sum|metrics_finalize|loss_finalize|step_epilogue" \
    -s 12 -c 12 -o gpurun_out/${TAG}_hot_kernels -f python tools/run_pre.py loss > gpurun_out/${TAG}_ncu2.log 2>&1
tail -2 gpurun_out/${TAG}_ncu2.log
