#!/bin/bash
O=gpurun_out/r02
mkdir -p $O
for cfg in "508 2032" "1016 4064" "762 3048" "508 4064" "508 2032"; do
  set -- $cfg
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --aux none --e2e-steps 5 --chunk $1 --batch $2 > $O/bench47_c$1_b$2.json 2> $O/bench47_c$1_b$2.err
  python -c "
import json
d=json.load(open('$O/bench47_c$1_b$2.json')); print($1, $2, round(d['value']), round(d['e2e']['value']), d['roofline']['per_class_ms_one_step'], round(d['roofline']['frac'],3), d['clocks']['sm_mhz'])"
done
