#!/bin/bash
O=gpurun_out/r02
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_encoder.py -x -q -k "golden or host or batch or ragged" 2>&1 | tail -n 2
for r in 1 2; do
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --aux none --e2e-steps 10 > $O/bench48_$r.json 2> $O/bench48_$r.err
  python -c "
import json
d=json.load(open('$O/bench48_$r.json')); print(round(d['value']), round(d['e2e']['value']), round(d['e2e']['value']/d['value'],3), d['clocks']['sm_mhz'])"
done
