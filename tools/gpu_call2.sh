#!/bin/bash
O=gpurun_out/r02
mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_tc -s 1 -c 1 -o $O/attn_v2_s197 -f python tools/attn_one.py > $O/ncu_attn.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_tc -s 1 -c 1 -o $O/attn_v2_s257 -f python tools/attn_one.py 127 257 16 >> $O/ncu_attn.log 2>&1
tail -5 $O/ncu_attn.log
