#!/bin/bash
# Run on the GPU box (gpurun -- 'bash profiles/tools/gpu_profile.sh TAG'): the ncu launch list of the bench
# command and one `--set full` capture per hot kernel; everything lands in gpurun_out/ and is summarised
# here afterwards with profiles/tools/summarize.py.
tag=${1:-r03}
out=gpurun_out
mkdir -p $out
BENCH="python bench.py --steps 2 --warmup 3 --prof-steps 0 --no-cpu-baseline"
# launch list: skip the warm-up launches roughly (agent construction + 3 warm-ups), keep two updates
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $out/${tag}_launches.csv $BENCH > $out/${tag}_launches.log 2>&1
cap() {  # name regex count
  ncu --set full --clock-control none --import-source on -k regex:$2 -s 6 -c $3 -o $out/${tag}_$1 -f $BENCH > $out/${tag}_ncu_$1.log 2>&1
}
cap conv96 'k_conv_tc96' 7
cap conv1 'k_conv_tc<' 2
cap wgrad 'k_conv_wgrad_tc' 4
cap gather 'k_gather_s2d_u8' 1
cap gemmtc 'k_gemm_tc' 6
cap curl 'k_curl_tc' 2
ls -la $out | grep $tag
