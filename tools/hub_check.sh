#!/bin/bash
# throughput of the hub families that exercise the new code paths (register tokens -> 261-token sequence, [class || mean] head) and a
# compute-sanitizer pass over the new kernels
O=gpurun_out/r02
mkdir -p $O
for a in "hibou_l 224 1016 254" "phikon_v2 224 1016 254" "midnight 224 508 127"; do
  timeout 200 python tools/encoder_bench.py $a 2>/dev/null | tail -n 1 | tee -a $O/hub_bench.log
done
timeout 200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest -q -x "tests/test_gpu_hub_families.py::test_class_mean_head_skips_register_tokens" "tests/test_gpu_hub_families.py::test_tiny_family_pixels_bit_exact_and_features[phikon_v1_test_tiny-300]" > $O/hub_memcheck.log 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" $O/hub_memcheck.log | tail -n 3
# CLIP tower throughput and a racecheck pass over the pool / class-row kernels (the 4 reports it prints are the known tcgen05.alloc false
# positive inside gemm_tcgen05_kernel: profiles/r01_compute_sanitizer.md)
timeout 100 python tools/encoder_bench.py plip 224 2032 508 2>/dev/null | tail -n 1 | tee -a $O/hub_bench.log
timeout 120 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest -q -x "tests/test_gpu_hub_families.py::test_class_mean_head_skips_register_tokens" > $O/hub_racecheck.log 2>&1; echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed" $O/hub_racecheck.log | tail -n 3
