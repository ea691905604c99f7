O=gpurun_out/r02
mkdir -p $O
timeout 100 python tools/encoder_bench.py plip 224 2032 508 2>/dev/null | tail -n 1 | tee -a $O/hub_bench2.log
timeout 120 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest -q -x "tests/test_gpu_hub_families.py::test_class_mean_head_skips_register_tokens" > $O/hub_racecheck.log 2>&1; echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed" $O/hub_racecheck.log | tail -n 3
