#!/bin/bash
# compute-sanitizer passes over the small-shape GPU tests (memcheck, racecheck, synccheck).
mkdir -p gpurun_out
SEL='test_random_vs_oracle or test_golden or test_band_entry_points or test_special_values or test_empty_inputs'
for tool in memcheck racecheck synccheck; do
  echo "== $tool (KernelWeighting / Scatter2Gather parity tests)"
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$SEL" 2>&1 | tail -6
done 2>&1 | tee gpurun_out/sanitizer_kw.txt
for tool in memcheck racecheck; do
  echo "== $tool (fused splat / conv1x1 / U-net glue tests)"
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 7 python -m pytest tests/test_modules.py tests/test_conv1x1.py -m gpu -q -x -k "fused or chain or upsample or unet_fast" 2>&1 | tail -6
done 2>&1 | tee gpurun_out/sanitizer_fused.txt
