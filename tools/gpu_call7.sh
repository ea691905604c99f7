#!/bin/bash
O=gpurun_out/r02
mkdir -p $O
timeout 300 python tools/attn_precision.py > $O/attn_precision7.log 2>&1
timeout 1800 python -m pytest tests -m gpu -q > $O/t_gpu_all7.log 2>&1
( time timeout 1500 python bench.py --steps 10 --warmup 3 > $O/bench7.json 2> $O/bench7.err ) 2> $O/bench7.time
( time timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench7_ref.json 2> $O/bench7_ref.err ) 2> $O/bench7_ref.time
cat $O/attn_precision7.log; tail -6 $O/t_gpu_all7.log; tail -3 $O/bench7.err; cat $O/bench7.time; head -c 300 $O/bench7.json; echo; head -c 300 $O/bench7_ref.json
