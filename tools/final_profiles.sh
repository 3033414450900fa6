#!/bin/bash
# Round-end measurement set: bench (both arms), ncu launch list with DRAM bytes and tensor-pipe activity, full captures
# of the dominant launch types and of the cuBLAS GEMM behind MEASURED_PEAKS.json (calibration of the tensor-pipe metric).
# Profiling runs use `bench.py --quick` (warm-up + the device-resident timed loop only).
set -x
tag=${1:-r02}
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${tag}_reference.json 2>> gpurun_out/bench_$tag.err
tools/ncu_list.sh $tag ""
python tools/summarize_launches.py gpurun_out/launches_$tag.csv gpurun_out/traffic_$tag.json > gpurun_out/launches_$tag.txt
B="python bench.py --steps 1 --warmup 3 --quick"
# per decode: 12 conv_tc launches (position 9 = folded fused-MRF launch of stage 2, HBM-bound), 16 conv_tc2 (11 = stage-0
# k=11 c2 + residual, paired tiles), 1 conv_mrf128 (last pairs + MRF of stage 1), 7 conv_mrfp (6 = last pairs + MRF of
# stage 3), 8 conv_pairf (0 = C=128 k=3 pair, 4 = C=128 k=11 pair); the 4th decode is the timed one
ncu --set full --clock-control none --import-source on -k regex:conv_mrf128_kernel -s 3 -c 1 -f -o gpurun_out/prof_${tag}_stage1_tail $B > gpurun_out/ncu_full1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 45 -c 1 -f -o gpurun_out/prof_${tag}_stage2_mrf $B > gpurun_out/ncu_full7.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_tc2_kernel -s 59 -c 1 -f -o gpurun_out/prof_${tag}_stage0_k11_cta2 $B > gpurun_out/ncu_full5.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_mrfp_kernel -s 27 -c 1 -f -o gpurun_out/prof_${tag}_mrfp3 $B > gpurun_out/ncu_full2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_pairf_kernel -s 24 -c 1 -f -o gpurun_out/prof_${tag}_pairf128 $B > gpurun_out/ncu_full3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_pairf_kernel -s 28 -c 1 -f -o gpurun_out/prof_${tag}_pairf128_k11 $B > gpurun_out/ncu_full6.log 2>&1
python tools/cublas_peak.py > gpurun_out/cublas_peak_$tag.txt 2>&1
ncu --set full --clock-control none -k regex:"gemm|nvjet|cutlass|sm100" -s 8 -c 1 -f -o gpurun_out/prof_${tag}_cublas_8192 python tools/cublas_peak.py > gpurun_out/ncu_full4.log 2>&1
ls -la gpurun_out/*.ncu-rep
