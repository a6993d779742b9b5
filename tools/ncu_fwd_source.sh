python tools/ncu_case.py fwd 4096 256 > gpurun_out/plain_f.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"infonce_fwd_tc" -c 1 -o gpurun_out/fwd_src python tools/ncu_case.py fwd 4096 256 > gpurun_out/ncu_f.log 2>&1
ncu -i gpurun_out/fwd_src.ncu-rep --page source --csv --print-source sass > gpurun_out/fwd_source_sass.csv 2>gpurun_out/ncu_f2.log
ncu -i gpurun_out/fwd_src.ncu-rep --page details --csv 2>/dev/null | grep -E "Duration|Tensor|XU|Issue Slots|Executed Ipc|Stall|Warp Cycles|ALU|FMA|LSU" > gpurun_out/fwd_details.csv
rm -f gpurun_out/fwd_src.ncu-rep
wc -c gpurun_out/fwd_source_sass.csv
