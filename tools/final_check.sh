#!/bin/bash
O=gpurun_out/r02
mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -x -q > $O/t_gpu_final.log 2>&1; tail -n 3 $O/t_gpu_final.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_final.log 2>&1; tail -n 2 $O/smoke_final.log
( time timeout 900 python bench.py > $O/bench_final.json 2> $O/bench_final.err ) 2> $O/bench_final.time; tail -n 1 $O/bench_final.time | head -1; grep real $O/bench_final.time
( time timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_final_ref.json 2> $O/bench_final_ref.err ) 2> $O/bench_final_ref.time; grep real $O/bench_final_ref.time
python -c "
import json
d=[json.loads(l) for l in open('$O/bench_final.json') if l.startswith('{')][-1]
print(round(d['value']), round(d['e2e']['value']), d['roofline']['per_class_ms_one_step'], round(d['roofline']['frac'],3), d['clocks'], d['gpu_launches'])
r=[json.loads(l) for l in open('$O/bench_final_ref.json') if l.startswith('{')][-1]
print('ref', r['value'], r['cpu_baseline']['cores'], r['cpu_baseline']['kind'])"
