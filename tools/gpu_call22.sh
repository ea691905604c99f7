#!/bin/bash
O=gpurun_out/r02
mkdir -p $O
for r in 1 2; do
for v in 32 0; do
  timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --aux none --e2e-steps 3 --opt attn_variant=$v > $O/bench22_v${v}_$r.json 2> $O/bench22_v${v}_$r.err
  python -c "
import json
d=json.load(open('$O/bench22_v${v}_$r.json')); print('variant', $v, 'run', $r, round(d['value']), round(d['e2e']['value']), d['roofline']['per_class_ms_one_step'], d['clocks']['sm_mhz'])"
done; done
