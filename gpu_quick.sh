#!/bin/bash
# quick GPU session: tests, then whatever else is passed
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; grep -m1 "model name" /proc/cpuinfo >> gpurun_out/gpu.txt; free -g | head -2 >> gpurun_out/gpu.txt
ldconfig -p | grep -E 'libEGL|libOSMesa|libGLX_nvidia|libEGL_nvidia|libGL\.so' >> gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 | tee gpurun_out/pytest_gpu.log
