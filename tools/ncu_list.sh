#!/bin/bash
# usage: tools/ncu_list.sh <tag> [VITSDEC_OPTS]   -> gpurun_out/launches_<tag>.csv (one warm-up + one timed decode step)
tag=$1; opts=$2
VITSDEC_OPTS=$opts ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none -c 2000 --csv \
  --log-file gpurun_out/launches_$tag.csv python bench.py --steps 1 --warmup 3 --quick > gpurun_out/ncu_$tag.log 2>&1
