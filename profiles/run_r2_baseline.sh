#!/bin/bash
# round-2 baseline: FP64 micro-benchmark + ncu --set full captures of K1, K2, K3 (+ median), K5 of the CURRENT build
cd /root/repo
O=gpurun_out/r2base; mkdir -p $O
nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o profiles/dfma_bench profiles/dfma_bench.cu && ./profiles/dfma_bench | tee $O/dfma.txt
M='gpu__time_duration.sum,launch__grid_size,launch__block_size,launch__registers_per_thread,launch__shared_mem_per_block_dynamic,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,sm__icc_request_hit_rate.pct,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,smsp__pcsamp_warps_issue_stalled'
cap() {  # name, kernel regex, skip, count, command...
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  "$@" > $O/$name.plain.log 2>&1 || { echo "$name plain run failed"; tail -5 $O/$name.plain.log; return; }
  ncu --set full --clock-control none --import-source on -k "regex:$rx" --launch-skip $skip -c $cnt -f -o $O/$name "$@" > $O/$name.ncu.log 2>&1
  ncu -i $O/$name.ncu-rep --page raw --csv > $O/$name.raw.csv 2>/dev/null
  ncu -i $O/$name.ncu-rep --page source --csv --print-source sass > $O/$name.sass.csv 2>/dev/null
  ls -la $O/$name.ncu-rep
  rm -f $O/$name.ncu-rep
}
NO_TICKS=1 NC=100 cap k3 'k3_windows|k3_median' 4 2 python profiles/k3_phase_timing.py
NC=64 cap k2 'k2_pass' 4 4 python profiles/k2_run.py
NC=100 cap k1 'k1_' 4 4 python profiles/k1_run.py
cap k5 'k5_recon' 1 1 python profiles/k5_run.py
NO_TICKS=1 NC=200 python profiles/k3_phase_timing.py 2>&1 | tail -4 | tee $O/k3_times.txt
