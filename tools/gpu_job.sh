#!/bin/bash
# One gpurun call = one batch of GPU work (this is the script the builder edits between calls; it travels with the repo snapshot).
mkdir -p gpurun_out
echo "=== ncu wgrad (F=512, both kernels)"
F=512 timeout 900 ncu --metrics gpu__time_duration.sum,lts__t_bytes.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__m_xbar2l1tex_read_bytes.sum --clock-control none -k regex:"wgrad_tc_kernel|gemm_tc_kernel" --launch-skip 4 -c 24 --csv --log-file gpurun_out/wgrad_ncu.csv python tools/wgrad_bench.py > gpurun_out/wgrad_ncu.log 2>&1
tail -3 gpurun_out/wgrad_ncu.log; wc -l gpurun_out/wgrad_ncu.csv
