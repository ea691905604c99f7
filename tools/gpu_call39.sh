#!/bin/bash
O=gpurun_out/r02
mkdir -p $O
( time timeout 850 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29527 bench.py --gpus 8 --steps 20 --warmup 5 > $O/bench39_8gpu.json 2> $O/bench39_8gpu.err ) 2> $O/bench39_8gpu.time
tail -n 3 $O/bench39_8gpu.err; cat $O/bench39_8gpu.time
python -c "
import json
d=[json.loads(l) for l in open('$O/bench39_8gpu.json') if l.startswith('{')][-1]
print(round(d['value']), round(d['e2e']['value']), d['n_gpus'], d['clocks'], d.get('cpu_baseline',{}).get('cores'))
print(json.dumps(d['aux'])[:2500])"
