#!/bin/bash
O=gpurun_out/r02
mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_tc6 -s 2 -c 1 -f -o $O/attn_v6b_s197 python tools/attn_one.py > $O/ncu_attn16.log 2>&1
tail -3 $O/ncu_attn16.log
