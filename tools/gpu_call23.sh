#!/bin/bash
O=gpurun_out/r02
mkdir -p $O
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench23_2gpu.json 2> $O/bench23_2gpu.err ) 2> $O/bench23_2gpu.time
tail -3 $O/bench23_2gpu.err; cat $O/bench23_2gpu.time
python -c "
import json
d=json.load(open('$O/bench23_2gpu.json')); print(round(d['value']), round(d['e2e']['value']), d['n_gpus'], d.get('cpu_baseline',{}).get('cores'), json.dumps(d['aux'])[:1500])"
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > $O/bench23_2gpu_ref.json 2> $O/bench23_2gpu_ref.err ) 2> $O/bench23_2gpu_ref.time
cat $O/bench23_2gpu_ref.json | cut -c1-600; cat $O/bench23_2gpu_ref.time
