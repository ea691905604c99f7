#!/bin/bash
O=gpurun_out/r02
mkdir -p $O
timeout 2400 python tools/dinov2_precision.py dinov2_giant 0:4:4 0:8:8 0:12:12 0:16:16 0:8:4 > $O/giant_precision11.log 2>&1
for c in 127 254 508; do
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --aux none --e2e-steps 3 --chunk $c --batch 2032 > $O/bench11_c$c.json 2> $O/bench11_c$c.err
done
cat $O/giant_precision11.log; for c in 127 254 508; do python -c "
import json,sys
d=json.load(open('$O/bench11_c$c.json')); print($c, round(d['value']), round(d['e2e']['value']), d['roofline']['per_class_ms_one_step'], d['clocks']['sm_mhz'])"; done
