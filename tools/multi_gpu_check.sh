#!/bin/bash
# one N-GPU run of the bench the way the driver launches it:  bash tools/multi_gpu_check.sh N
N=${1:-2}
O=gpurun_out/r02
mkdir -p $O
( time timeout 850 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29537 bench.py --gpus $N --steps 20 --warmup 5 > $O/bench_${N}gpu.json 2> $O/bench_${N}gpu.err ) 2> $O/bench_${N}gpu.time
grep real $O/bench_${N}gpu.time
python -c "
import json
d=[json.loads(l) for l in open('$O/bench_${N}gpu.json') if l.startswith('{')][-1]
a=d['aux']
print(d['n_gpus'], round(d['value']), round(d['e2e']['value']), d['clocks'], d.get('cpu_baseline',{}).get('cores'))
print('c2', a['c2']['ms_max_over_ranks'], 'c3', round(a['c3']['patches_per_s']), a['c3']['max_rel_err_vs_golden'], 'c4', round(a['c4']['patches_per_s']), a['c4']['max_rel_err_vs_golden'], a['c4'].get('gather'))"
