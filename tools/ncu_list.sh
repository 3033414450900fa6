#!/bin/bash
# usage: tools/ncu_list.sh <tag> [VITSDEC_OPTS]   -> gpurun_out/launches_<tag>.csv
# Device time, DRAM bytes and tensor-pipe activity of every launch of the decode (load-time packing kernels are skipped;
# two decodes are captured, tools/summarize_launches.py reports the second).
tag=$1; opts=$2
VITSDEC_OPTS=$opts ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed \
  --clock-control none -k regex:"conv_tc_kernel|conv_tc2_kernel|conv_pair|conv_mrfp_kernel|conv_mrf128_kernel|pack_z_kernel|cond_kernel" -c 130 --csv \
  --log-file gpurun_out/launches_$tag.csv python bench.py --steps 1 --warmup 3 --quick > gpurun_out/ncu_$tag.log 2>&1
