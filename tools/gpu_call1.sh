#!/bin/bash
# GPU call 1 of round 2: LDTM microbenchmark, attention v2 parity + timing, full GPU suite, bench at three chunk sizes
mkdir -p gpurun_out/r02
O=gpurun_out/r02
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/gpu.txt 2>&1
timeout 120 tools/bin/ldtm_bench > $O/ldtm_bench.log 2>&1
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -k attention > $O/t_attention.log 2>&1
timeout 300 python tools/attn_time.py > $O/attn_time.log 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > $O/t_gpu_all.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_c127.json 2> $O/bench_c127.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --chunk 63 --batch 2016 > $O/bench_c63.json 2> $O/bench_c63.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --chunk 31 --batch 1984 > $O/bench_c31.json 2> $O/bench_c31.err
timeout 600 python tools/encoder_bench.py dinov2_large 224 > $O/encoder_bench_large.log 2>&1
timeout 600 python tools/encoder_bench.py dinov2_giant 512 508 > $O/encoder_bench_giant.log 2>&1
tail -3 $O/t_attention.log $O/t_gpu_all.log; cat $O/attn_time.log $O/encoder_bench_large.log $O/encoder_bench_giant.log; head -c 600 $O/bench_c127.json
