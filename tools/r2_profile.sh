#!/bin/bash
# round 2 ncu set (one B200, --clock-control none): launch list of the default bench command + one full capture per kernel of the path
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
BENCH2="python bench.py --config c2 --steps 2 --warmup 3 --no-cpu-baseline --no-c3 --sampler-ms 0"
# every launch of the default bench command (c2 block + c3 block) with its device time: cold-cache and serialised, compare SHARES
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2_launches_c2.csv $BENCH2 > gpurun_out/r2_launches_c2.log 2>&1
# config 2: lighting (warp per request), draw, commit, compaction (count / write / scan)
$NCU -k regex:dn_light_kernel -s 12 -c 1 -f -o gpurun_out/r2_light_warp_c2 $BENCH2 --light-kernel warp > gpurun_out/r2_p1.log 2>&1
$NCU -k regex:dn_draw_kernel -s 12 -c 1 -f -o gpurun_out/r2_draw_c2 $BENCH2 --light-kernel warp > gpurun_out/r2_p2.log 2>&1
$NCU -k regex:dn_commit_kernel -s 12 -c 1 -f -o gpurun_out/r2_commit_c2 $BENCH2 --light-kernel warp > gpurun_out/r2_p3.log 2>&1
$NCU -k regex:dn_compact_kernel -s 24 -c 2 -f -o gpurun_out/r2_compact_c2 $BENCH2 --light-kernel warp > gpurun_out/r2_p4.log 2>&1
$NCU -k regex:dn_scan_blocks -s 12 -c 1 -f -o gpurun_out/r2_scan_c2 $BENCH2 --light-kernel warp > gpurun_out/r2_p5.log 2>&1
$NCU -k regex:dn_merge_visible -s 12 -c 1 -f -o gpurun_out/r2_merge_c2 $BENCH2 --light-kernel warp > gpurun_out/r2_p6.log 2>&1
# the initial upload of the terrain map: scatter kernel (first launch = a full 16384-chunk batch) and the opaque-flag refresh
$NCU -k regex:dn_scatter_chunks -s 0 -c 1 -f -o gpurun_out/r2_scatter_c2 $BENCH2 --light-kernel warp > gpurun_out/r2_p7.log 2>&1
$NCU -k regex:dn_refresh_opaque -s 0 -c 1 -f -o gpurun_out/r2_refresh_c2 $BENCH2 --light-kernel warp > gpurun_out/r2_p8.log 2>&1
# sparse map (config 3 at 1/8 volume): persistent kernel (one whole dispatch), warp kernel, wavefront pair (a full pass)
$NCU -k regex:dn_light_flat -s 2 -c 1 -f -o gpurun_out/r2_light_flat_c3s python tools/light_sweep.py c3s 1 flat > gpurun_out/r2_p9.log 2>&1
$NCU -k regex:dn_light_kernel -s 2 -c 1 -f -o gpurun_out/r2_light_warp_c3s python tools/light_sweep.py c3s 1 warp > gpurun_out/r2_p10.log 2>&1
$NCU -k regex:dn_wave_step -s 120 -c 1 -f -o gpurun_out/r2_wave_step_c3s python tools/light_sweep.py c3s 1 wave > gpurun_out/r2_p11.log 2>&1
$NCU -k regex:dn_wave_serve -s 120 -c 1 -f -o gpurun_out/r2_wave_serve_c3s python tools/light_sweep.py c3s 1 wave > gpurun_out/r2_p12.log 2>&1
# dense map (config 5 at 1/64 volume): warp kernel
$NCU -k regex:dn_light_kernel -s 2 -c 1 -f -o gpurun_out/r2_light_warp_c5s python tools/light_sweep.py c5s 1 warp > gpurun_out/r2_p13.log 2>&1
# picking and the peer-memory kernels (3 replicas on one GPU / pushes after the wavefront kernels), from the parity suite
$NCU -k regex:dn_pick -c 1 -f -o gpurun_out/r2_pick python -m pytest tests/test_parity_gpu.py -q -x -k "picking_against_reference_goldens and warp" > gpurun_out/r2_p14.log 2>&1
$NCU -k "regex:dn_peer|dn_push_staging|dn_merge_visible_peers" -c 8 -f -o gpurun_out/r2_peer python -m pytest tests/test_parity_gpu.py -q -x -k "peer_sharded_equals_unsharded and wave" > gpurun_out/r2_p15.log 2>&1
ls -la gpurun_out | grep "r2_.*ncu-rep"
