#!/bin/bash
O=gpurun_out/r02
mkdir -p $O
M="dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed,sm__issue_active.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,launch__grid_size"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/r02_launches.csv python bench.py --steps 2 --warmup 3 --batch 254 --e2e-steps 1 --no-cpu-baseline --aux none > $O/r02_launches_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_tcgen05|attention_tc" -s 53 -c 10 -f -o /tmp/r02_vitb python tools/ncu_chunk.py > $O/ncu35_vitb.log 2>&1
ncu -i /tmp/r02_vitb.ncu-rep --page raw --csv > $O/r02_ncu_full_vitb_gemm_attention.csv 2>/dev/null
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_tcgen05|attention_tc" -s 36 -c 14 -f -o /tmp/r02_giant python tools/ncu_chunk.py model=dinov2_giant > $O/ncu35_giant.log 2>&1
ncu -i /tmp/r02_giant.ncu-rep --page raw --csv > $O/r02_ncu_full_giant_gemm_attention.csv 2>/dev/null
timeout 900 ncu --metrics $M --clock-control none -k regex:"gemm_tcgen05|attention_tc|layernorm" -s 0 -c 260 --csv --log-file $O/r02_giant_chunk_metrics.csv python tools/ncu_chunk.py model=dinov2_giant > $O/ncu35_giant2.log 2>&1
for f in $O/ncu35_vitb.log $O/ncu35_giant.log $O/ncu35_giant2.log; do tail -n 2 $f; done
ls -la $O/r02_*.csv
du -sh gpurun_out
