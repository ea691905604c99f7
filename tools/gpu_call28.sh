#!/bin/bash
O=gpurun_out/r02
mkdir -p $O
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/sam2_large_launches28.csv python tools/sam2_bench.py large 1 > $O/sam2_ncu28.log 2>&1
tail -2 $O/sam2_ncu28.log
