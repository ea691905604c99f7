#!/bin/bash
O=gpurun_out/r02
mkdir -p $O
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02_sam2_hiera_l_launches_final.csv python tools/sam2_bench.py large 1 > $O/sam_prof_a.log 2>&1
M="gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__issue_active.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,launch__grid_size,launch__registers_per_thread"
timeout 600 ncu --metrics $M --clock-control none -k regex:"sam_attention_tc|gemm_tcgen05" -s 400 -c 120 --csv --log-file $O/r02_sam2_gemm_attention_metrics.csv python tools/sam2_bench.py large 1 > $O/sam_prof_b.log 2>&1
tail -n 1 $O/sam_prof_a.log; ls -la $O/r02_sam2_*.csv
