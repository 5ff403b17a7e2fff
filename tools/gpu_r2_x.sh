#!/bin/bash
# is the A multicast active? L2 read sectors of the K-heavy projections with and without it (debug library, ncu);
# P5_GEMM_PREFER=0 makes 8 the REGULAR cluster dimension of those launches
mkdir -p gpurun_out
for mc in 1 0; do
P5_GEMM_PREFER=0 P5_GEMM_MULTICAST=$mc timeout 600 ncu --metrics gpu__time_duration.sum,lts__t_sectors_op_read.sum,dram__bytes_read.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:gemm_tcgen05 -s 9 -c 4 --csv --log-file gpurun_out/ncu_mc_$mc.csv \
    python tools/profile_target.py --debug > gpurun_out/ncu_mc_$mc.log 2>&1
tail -1 gpurun_out/ncu_mc_$mc.log | cut -c1-100
done
