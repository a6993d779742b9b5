#!/bin/bash
# Round-2 profiling pass (run under gpurun, one GPU): launch list of the bench step + `ncu --set full` of the
# top kernels.  Every ncu run follows a plain run of the same command that exited 0 (B200_PROFILING.md).
set -u
python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_raw.csv \
    python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/ncu_launch.log 2>&1
python tools/ncu_case.py grad 4096 256 > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"infonce_grad_tc5|infonce_fwd_tc|grad_finish_pair|l2norm_pair" -c 8 \
    -o gpurun_out/r2_loss_c2 python tools/ncu_case.py grad 4096 256 > gpurun_out/ncu_a.log 2>&1
python tools/ncu_case.py grad 4096 512 > gpurun_out/plain3.log 2>&1 &&
ncu --set full --clock-control none -k regex:"infonce_grad_tc8" -c 2 -o gpurun_out/r2_tc8 \
    python tools/ncu_case.py grad 4096 512 > gpurun_out/ncu_b.log 2>&1
python tools/ncu_case.py topk 8192 512 262144 > gpurun_out/plain4.log 2>&1 &&
ncu --set full --clock-control none -k regex:"topk_tc_kernel" -c 1 -o gpurun_out/r2_topk \
    python tools/ncu_case.py topk 8192 512 262144 > gpurun_out/ncu_c.log 2>&1
ls -la gpurun_out/*.ncu-rep
# gpurun brings back at most 64 MiB: keep the text summaries (and the source-level stall page of the loss
# kernels), drop the reports
python tools/ncu_summary.py gpurun_out/r2_loss_c2.ncu-rep gpurun_out/r2_tc8.ncu-rep gpurun_out/r2_topk.ncu-rep \
    > gpurun_out/r2_ncu_full_summary.txt 2> gpurun_out/ncu_summary.err
ncu -i gpurun_out/r2_loss_c2.ncu-rep --page details --csv 2>/dev/null | grep -E "infonce_grad_tc5|infonce_fwd_tc" | \
    grep -E "Duration|Tensor|XU|Issue Slots|Executed Ipc|Registers|Achieved Occupancy|L2 Cache Throughput|DRAM Throughput|Warp Cycles Per Issued|Stall" \
    > gpurun_out/r2_ncu_details_loss.csv
rm -f gpurun_out/*.ncu-rep
tail -n 2 gpurun_out/ncu_a.log
tail -n 2 gpurun_out/ncu_launch.log
wc -c gpurun_out/r2_ncu_full_summary.txt gpurun_out/r2_launches_raw.csv
