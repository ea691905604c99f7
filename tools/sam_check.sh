#!/bin/bash
O=gpurun_out/r02
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_sam2.py -x -q > $O/t_sam2_49.log 2>&1; tail -n 3 $O/t_sam2_49.log
timeout 600 python tools/sam2_bench.py > $O/sam2_bench49.log 2>&1; cat $O/sam2_bench49.log
