#!/bin/bash
O=gpurun_out/r02
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -k attention > $O/t_attention45.log 2>&1; tail -n 5 $O/t_attention45.log
timeout 600 python tools/attn_time.py > $O/attn_time45.log 2>&1; cat $O/attn_time45.log
timeout 600 python tools/attn_precision.py > $O/attn_precision45.log 2>&1; cat $O/attn_precision45.log
