#!/bin/bash
O=gpurun_out/r02
mkdir -p $O
timeout 900 python bench.py --steps 10 --warmup 3 > $O/bench21.json 2> $O/bench21.err; tail -2 $O/bench21.err
python -c "
import json
d=json.load(open('$O/bench21.json')); print(round(d['value']), round(d['e2e']['value']), d['roofline']['per_class_ms_one_step'], d['roofline']['frac'], d['clocks'])
print({k:(v.get('patches_per_s'), v.get('max_rel_err_vs_golden'), v.get('ms_per_thumbnail_incl_h2d_d2h')) for k,v in d['aux'].items() if isinstance(v, dict)})"
timeout 1500 python -m pytest tests -m gpu -x -q > $O/t_gpu_all21.log 2>&1; tail -4 $O/t_gpu_all21.log
