#!/bin/bash
O=gpurun_out/r02
mkdir -p $O
timeout 2400 python tools/dinov2_precision.py dinov2_giant 0:40:0 0:40:8 0:40:20 0:40:39 0:20:20 0:30:10 > $O/giant_precision10.log 2>&1
cat $O/giant_precision10.log
