#!/bin/bash
# round 2, GPU session 2: diagnostics of the sparse-map parity failure, the rest of the parity suite, full-pass captures of both step kernels
mkdir -p gpurun_out
timeout 600 python tools/r2_sessions/r2_diag_sparse.py sparse 20,20,20 > gpurun_out/g2_diag_sparse.log 2>&1
timeout 600 python tools/r2_sessions/r2_diag_sparse.py dense 8,8,8 > gpurun_out/g2_diag_dense.log 2>&1
tail -c 3000 gpurun_out/g2_diag_sparse.log
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_parity_gpu.py::test_sparse_balls_against_oracle 2>&1 | tail -25 > gpurun_out/g2_pytest.log
tail -5 gpurun_out/g2_pytest.log
# full passes (pass 6 of the third dispatch of the sweep: 2 x 64 + 5 launches skipped)
DN_B200_WAVE_STEP=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:dn_wave_step2 -s 133 -c 1 -f -o gpurun_out/g2_step2_c3s python tools/light_sweep.py c3s 3 wave > gpurun_out/g2_ncu2.log 2>&1
DN_B200_WAVE_STEP=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:dn_wave_step_kernel -s 133 -c 1 -f -o gpurun_out/g2_step1_c3s python tools/light_sweep.py c3s 3 wave > gpurun_out/g2_ncu1.log 2>&1
DN_B200_WAVE_STEP=2 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:dn_wave -s 256 -c 140 --csv --log-file gpurun_out/g2_wave2_launches.csv python tools/light_sweep.py c3s 3 wave > /dev/null 2>&1
ls -la gpurun_out | grep g2_
