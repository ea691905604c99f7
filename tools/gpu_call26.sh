#!/bin/bash
O=gpurun_out/r02
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_sam2.py -x -q > $O/t_sam2_26.log 2>&1; tail -15 $O/t_sam2_26.log
timeout 600 python tools/sam2_bench.py > $O/sam2_bench26.log 2>&1; cat $O/sam2_bench24.log
