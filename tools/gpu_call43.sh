#!/bin/bash
O=gpurun_out/r02
mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_units -s 2 -c 1 -f -o /tmp/attn_units python tools/attn_one.py > $O/ncu_attn43.log 2>&1
ncu -i /tmp/attn_units.ncu-rep --page source --csv --print-source sass > $O/attn_units_src.csv 2>/dev/null
ncu -i /tmp/attn_units.ncu-rep --page raw --csv > $O/attn_units_raw.csv 2>/dev/null
tail -n 2 $O/ncu_attn43.log; ls -la $O/attn_units_*.csv
