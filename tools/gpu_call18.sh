#!/bin/bash
O=gpurun_out/r02
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -k attention > $O/t_attention18.log 2>&1; tail -3 $O/t_attention18.log
python tools/attn6_trace.py 0 > $O/attn6_trace18_emu0.log 2>&1; head -16 $O/attn6_trace18_emu0.log
timeout 600 python tools/attn6_check.py > $O/attn6_check18.log 2>&1; cat $O/attn6_check18.log
