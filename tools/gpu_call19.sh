#!/bin/bash
O=gpurun_out/r02
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -k attention > $O/t_attention19.log 2>&1; tail -3 $O/t_attention19.log
timeout 600 python tools/attn6_check.py > $O/attn6_check19.log 2>&1; cat $O/attn6_check19.log
