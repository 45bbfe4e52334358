#!/bin/bash
# session 30: compute-sanitizer memcheck and racecheck of the FINAL build: every lighting kernel incl. auto-mode split dispatches, edits, picking,
# checkpoint, and the peer-store path (three replicas on one GPU)
mkdir -p gpurun_out
SEL="mixed_materials or lighting_split_and_edits or empty_map or picking_equals or checkpoint or peer_sharded_equals_unsharded or bulk_edit_stream"
for tool in memcheck racecheck; do
  timeout 1700 compute-sanitizer --tool $tool --log-file gpurun_out/g30_sanitizer_$tool.log python -m pytest tests/test_parity_gpu.py -q -x -k "$SEL" 2>&1 | tail -2
  tail -3 gpurun_out/g30_sanitizer_$tool.log
done
