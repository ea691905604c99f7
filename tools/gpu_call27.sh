#!/bin/bash
O=gpurun_out/r02
mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_sam2.py tests/test_gpu_kernels.py tests/test_gpu_encoder.py -x -q > $O/t_27.log 2>&1; tail -4 $O/t_27.log
timeout 600 python tools/sam2_bench.py > $O/sam2_bench27.log 2>&1; cat $O/sam2_bench27.log
