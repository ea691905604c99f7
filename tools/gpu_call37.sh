#!/bin/bash
O=gpurun_out/r02
mkdir -p $O
for mb in 127 254 508; do timeout 600 python tools/encoder_bench.py dinov2_large 224 2032 $mb 2>&1 | tail -1; done | tee $O/encoder_bench37_large.log
for mb in 127 254 508; do timeout 600 python tools/encoder_bench.py dinov2_giant 512 1016 $mb 2>&1 | tail -1; done | tee $O/encoder_bench37_giant.log
for mb in 127 508; do timeout 600 python tools/encoder_bench.py vit_l_16 256 2032 $mb 2>&1 | tail -1; done | tee $O/encoder_bench37_vitl.log
