#!/bin/bash
# Round-end measurement set: bench (both arms), ncu launch list with DRAM bytes, three full captures.
set -x
python bench.py --steps 30 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_final_reference.json 2>> gpurun_out/bench_final.err
tools/ncu_list.sh final ""
B="python bench.py --steps 1 --warmup 3 --quick"
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 180 -c 1 -f -o gpurun_out/prof_r01b_stage1_k11 $B > gpurun_out/ncu_full1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 198 -c 1 -f -o gpurun_out/prof_r01b_mrf3_folded $B > gpurun_out/ncu_full2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_pairf_kernel -s 3 -c 1 -f -o gpurun_out/prof_r01b_pairf_c32_k11 $B > gpurun_out/ncu_full3.log 2>&1
ls -la gpurun_out/*.ncu-rep
