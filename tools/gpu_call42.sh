#!/bin/bash
O=gpurun_out/r02
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -k attention > $O/t_attention42.log 2>&1; tail -n 5 $O/t_attention42.log
timeout 600 python tools/attn6_check.py > $O/attn6_check42.log 2>&1; cat $O/attn6_check42.log
