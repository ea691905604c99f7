#!/bin/bash
O=gpurun_out/r02
mkdir -p $O
for c in 127 508 254 762 1016 127 508; do
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --aux none --e2e-steps 5 --chunk $c --batch 2032 > $O/bench36_c$c.json 2> $O/bench36_c$c.err
  python -c "
import json
d=json.load(open('$O/bench36_c$c.json')); print($c, round(d['value']), round(d['e2e']['value']), d['roofline']['per_class_ms_one_step'], round(d['roofline']['frac'],3), d['clocks']['sm_mhz'])"
done
