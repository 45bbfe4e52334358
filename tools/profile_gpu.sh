#!/bin/bash
# ncu session on the GPU box (run under gpurun): launch list of one bench command + full captures of the hot kernels.
# usage: bash tools/profile_gpu.sh [tag]
TAG=${1:-r1}
mkdir -p gpurun_out
CMD="python bench.py --config c2 --steps 2 --warmup 3 --no-cpu-baseline --sampler-ms 0"
# every launch of the default bench command with its device time (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_launches.log 2>&1
# full captures: the warp-per-request lighting kernel, the draw kernel and the commit kernel on config 2
ncu --set full --clock-control none --import-source on -k regex:dn_light_kernel -s 4 -c 1 -f -o gpurun_out/${TAG}_light_warp $CMD --light-kernel warp > gpurun_out/${TAG}_light_warp.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:dn_draw_kernel -s 4 -c 1 -f -o gpurun_out/${TAG}_draw $CMD > gpurun_out/${TAG}_draw.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:dn_commit -s 4 -c 1 -f -o gpurun_out/${TAG}_commit $CMD > gpurun_out/${TAG}_commit.log 2>&1
# the persistent and the wavefront kernels on the sparse map (one steady-state dispatch / one full pass)
if [ "$2" == "sparse" ]; then
ncu --set full --clock-control none --import-source on -k regex:dn_light_flat -s 4 -c 1 -f -o gpurun_out/${TAG}_light_flat_c3s python bench.py --config c3s --steps 1 --warmup 3 --no-cpu-baseline --sampler-ms 0 --light-kernel flat > gpurun_out/${TAG}_light_flat_c3s.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:dn_wave -s 680 -c 2 -f -o gpurun_out/${TAG}_wave_c3s python tools/light_sweep.py c3s 6 wave > gpurun_out/${TAG}_wave_c3s.log 2>&1
fi
ls -la gpurun_out/ | grep ${TAG}
