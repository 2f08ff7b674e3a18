#!/bin/bash
# Run on the GPU box: compute-sanitizer memcheck + racecheck over the kernel parity tests of the
# tcgen05 / TMA / bulk-copy kernels and one small whole update.
tag=${1:-r03}
out=gpurun_out
mkdir -p $out
SEL='test_conv_forward or test_conv_backward or test_gather or test_curl or test_gemm_layouts or test_conv_wgrad_staging or test_conv_wgrad_variants'
for tool in memcheck racecheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 66 --print-limit 20 \
      python -m pytest tests/test_kernels_gpu.py -x -q -k "$SEL" > $out/${tag}_sanitizer_${tool}_kernels.txt 2>&1
  echo "exit $?" >> $out/${tag}_sanitizer_${tool}_kernels.txt
  CURLA_GRAPH=0 timeout 900 compute-sanitizer --tool $tool --error-exitcode 66 --print-limit 20 \
      python -m pytest tests/test_update_parity_gpu.py -x -q -k "test_fused_update_equals_phased and crop90x160" > $out/${tag}_sanitizer_${tool}_update.txt 2>&1
  echo "exit $?" >> $out/${tag}_sanitizer_${tool}_update.txt
done
tail -n 6 $out/${tag}_sanitizer_*.txt
