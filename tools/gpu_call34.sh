#!/bin/bash
O=gpurun_out/r02
mkdir -p $O
M="dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed,sm__issue_active.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,launch__grid_size"
# 1. launch list of the bench command
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/r02_launches.csv python bench.py --steps 2 --warmup 3 --batch 254 --e2e-steps 1 --no-cpu-baseline --aux none > $O/r02_launches_bench.log 2>&1
# 2. full captures: ViT-B attention (197 tokens) + a few GEMMs of the second chunk
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_tcgen05|attention_tc" -s 53 -c 10 -f -o $O/r02_vitb_gemm_attention python tools/ncu_chunk.py > $O/ncu34_vitb.log 2>&1
# 3. DINOv2 giant: the GEMM shapes (split layers 7-8, plain layers 9+) and the 257-token attention
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_tcgen05|attention_tc" -s 36 -c 14 -f -o $O/r02_giant_gemm_attention python tools/ncu_chunk.py model=dinov2_giant > $O/ncu34_giant.log 2>&1
timeout 900 ncu --metrics $M --clock-control none -k regex:"gemm_tcgen05|attention_tc|layernorm" -s 0 -c 260 --csv --log-file $O/r02_giant_chunk_metrics.csv python tools/ncu_chunk.py model=dinov2_giant > $O/ncu34_giant2.log 2>&1
tail -2 $O/ncu34_vitb.log $O/ncu34_giant.log $O/ncu34_giant2.log
ls -la $O/*.ncu-rep | tail -5
