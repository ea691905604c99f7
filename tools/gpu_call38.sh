#!/bin/bash
O=gpurun_out/r02
mkdir -p $O
M="dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"
timeout 900 ncu --metrics $M --clock-control none -k regex:gemm_tcgen05 -s 49 -c 49 --csv --log-file $O/r02_ncu_gemm_dram_traffic.csv python tools/ncu_chunk.py chunk=508 > $O/ncu38.log 2>&1
tail -n 2 $O/ncu38.log
( time timeout 900 python bench.py > $O/bench38.json 2> $O/bench38.err ) 2> $O/bench38.time
tail -n 3 $O/bench38.err; cat $O/bench38.time
python -c "
import json
d=[json.loads(l) for l in open('$O/bench38.json') if l.startswith('{')][-1]
print(round(d['value']), round(d['e2e']['value']), d['roofline']['per_class_ms_one_step'], round(d['roofline']['frac'],3), d['clocks'], d['cpu_baseline']['value'], d['cpu_baseline']['cores'])
print({k:(v.get('patches_per_s'), v.get('max_rel_err_vs_golden'), v.get('ms_per_thumbnail_incl_h2d_d2h'), v.get('mask_iou_vs_golden')) for k,v in d['aux'].items() if isinstance(v, dict)})"
