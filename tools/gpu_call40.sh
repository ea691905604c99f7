#!/bin/bash
O=gpurun_out/r02
mkdir -p $O
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -x -q -k "attention" > $O/memcheck40_attention.log 2>&1; echo "rc=$?" >> $O/memcheck40_attention.log
tail -n 4 $O/memcheck40_attention.log
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_sam2.py -x -q -k "golden or service or batch" > $O/memcheck40_sam2.log 2>&1; echo "rc=$?" >> $O/memcheck40_sam2.log
tail -n 4 $O/memcheck40_sam2.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_encoder.py tests/test_gpu_dinov2.py -x -q -k "tiny or strict" > $O/memcheck40_encoder.log 2>&1; echo "rc=$?" >> $O/memcheck40_encoder.log
tail -n 4 $O/memcheck40_encoder.log
