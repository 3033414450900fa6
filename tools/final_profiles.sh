#!/bin/bash
# Round-end measurement set: bench (both arms), ncu launch list with DRAM bytes, three full captures.
# Profiling runs use `bench.py --quick` (warm-up + the device-resident timed loop only).
set -x
tag=${1:-r01d}
python bench.py --steps 30 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${tag}_reference.json 2>> gpurun_out/bench_$tag.err
tools/ncu_list.sh $tag ""
python tools/summarize_launches.py gpurun_out/launches_$tag.csv gpurun_out/traffic_$tag.json > gpurun_out/launches_$tag.txt
B="python bench.py --steps 1 --warmup 3 --quick"
# conv_tc launches per decode: 50 (positions 30 / 48 of the 4th decode = stage-1 k=11 c2 + residual / fused MRF of stage 3);
# conv_pair launches per decode: 10 (position 4 = C=32 k=3 pair)
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 180 -c 1 -f -o gpurun_out/prof_${tag}_stage1_k11 $B > gpurun_out/ncu_full1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 198 -c 1 -f -o gpurun_out/prof_${tag}_mrf3_folded $B > gpurun_out/ncu_full2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_pair_kernel -s 34 -c 1 -f -o gpurun_out/prof_${tag}_pair_c32_k3 $B > gpurun_out/ncu_full3.log 2>&1
ls -la gpurun_out/*.ncu-rep
