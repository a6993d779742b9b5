#!/bin/bash
# source-level (SASS) stall sampling of the d <= 256 backward at the bench shape; run under gpurun
python tools/ncu_case.py grad 4096 256 > gpurun_out/plain_g.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"infonce_grad_tc5" -c 1 -o gpurun_out/grad_src python tools/ncu_case.py grad 4096 256 > gpurun_out/ncu_g.log 2>&1
ncu -i gpurun_out/grad_src.ncu-rep --page source --csv --print-source sass > gpurun_out/grad_source_sass.csv 2>gpurun_out/ncu_g2.log
ncu -i gpurun_out/grad_src.ncu-rep --page details --csv 2>/dev/null | grep -E "Duration|Tensor|XU|Issue Slots|Executed Ipc|Stall|Warp Cycles|ALU|FMA|LSU|SM Frequency|Elapsed Cycles" > gpurun_out/grad_details.csv
rm -f gpurun_out/grad_src.ncu-rep
wc -c gpurun_out/grad_source_sass.csv
