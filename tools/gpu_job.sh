#!/bin/bash
# Batch of GPU work for one gpurun call (edited between calls; the script travels with the repo snapshot).
mkdir -p gpurun_out
echo "=== full gpu tests"; timeout 1800 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
echo "=== ncu launch list (one bench step, B=128)"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 420 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; wc -l gpurun_out/launches_r2.csv
echo "=== ncu full (2048-frame chunk), exported as csv on the box"
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,smsp__inst_executed.sum,sm__inst_executed.avg.per_cycle_active"
timeout 900 ncu --metrics $M --clock-control none -k regex:"firstconv_tc|conv3x3_zp|attention_kernel|maxpool3s2_kernel|affine_norm_zp|gemm_tc" -s 60 -c 60 --csv --log-file gpurun_out/kernels_r2.csv python bench.py --batch 16 --steps 1 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; wc -l gpurun_out/kernels_r2.csv
echo "=== bench.py (all extras)"; timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2c.json 2> gpurun_out/bench_r2c.err; tail -c 3000 gpurun_out/bench_r2c.json; tail -3 gpurun_out/bench_r2c.err
echo "=== bc ops breakdown"; timeout 600 python tools/bc_bench.py --ops 2>&1 | tail -40
du -sh gpurun_out
