#!/bin/bash
O=gpurun_out/r02
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -k attention > $O/t_attention20.log 2>&1; tail -3 $O/t_attention20.log
timeout 600 python tools/attn_time.py > $O/attn_time20.log 2>&1; cat $O/attn_time20.log
