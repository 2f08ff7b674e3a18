#!/bin/bash
# Run on the GPU box (gpurun -- 'bash profiles/tools/gpu_profile.sh TAG'): the ncu launch list of the bench
# command and one `--set full` capture per hot kernel; everything lands in gpurun_out/ and is summarised
# here afterwards with profiles/tools/summarize.py.  CURLA_GRAPH=0: the same kernels launched eagerly (ncu
# then names every launch; the graph replays the identical sequence).
tag=${1:-r03}
out=gpurun_out
mkdir -p $out
export CURLA_GRAPH=0
BENCH="python bench.py --steps 2 --warmup 3 --prof-steps 0 --no-cpu-baseline"
# launch list: skip the agent construction + set-up + warm-up launches roughly, keep about two updates
ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file $out/${tag}_launches.csv $BENCH > $out/${tag}_launches.log 2>&1
cap() {  # name regex skip count
  ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 -o $out/${tag}_$1 -f $BENCH > $out/${tag}_ncu_$1.log 2>&1
  ncu -i $out/${tag}_$1.ncu-rep --page raw --csv > $out/${tag}_$1_raw.csv 2>/dev/null
}
cap conv 'k_conv_tc' 40 8
cap wgrad 'k_conv_wgrad_tc' 16 4
cap gather 'k_gather_s2d_u8' 8 1
cap gemmtc 'k_gemm_tc' 60 8
cap curl 'k_curl_tc' 8 2
ls -la $out | grep $tag
