#!/bin/bash
O=gpurun_out/r02
mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_sam2.py -x -q > $O/t_29.log 2>&1; tail -4 $O/t_27.log
timeout 600 python tools/sam2_bench.py > $O/sam2_bench29.log 2>&1; cat $O/sam2_bench27.log
