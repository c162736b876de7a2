#!/bin/bash
# Run on the GPU box (gpurun): launch list of one bench run + ONE full ncu capture of the dominant kernel only.
set -e
TAG=${1:-r01f}
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_step_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu1.log 2>&1
python tools/run_pre.py loss > /dev/null
ncu --set full --clock-control none --import-source on -k regex:"loss_march" -s 1 -c 1 \
    -o gpurun_out/${TAG}_loss_march -f python tools/run_pre.py loss > gpurun_out/${TAG}_ncu2.log 2>&1
tail -2 gpurun_out/${TAG}_ncu2.log
