#!/bin/bash
# ncu session on the GPU box (run under gpurun): launch list of one bench command + full captures of the two hot kernels.
# usage: bash tools/profile_gpu.sh [tag] [config]
TAG=${1:-r1}
CFG=${2:-c2}
mkdir -p gpurun_out
CMD="python bench.py --config $CFG --steps 2 --warmup 3 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:dn_light_kernel -s 4 -c 2 -f -o gpurun_out/${TAG}_light $CMD > gpurun_out/${TAG}_light.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:dn_draw_kernel -s 4 -c 2 -f -o gpurun_out/${TAG}_draw $CMD > gpurun_out/${TAG}_draw.log 2>&1
ls -la gpurun_out/
