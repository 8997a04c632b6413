#!/bin/bash
O=gpurun_out; mkdir -p $O
python tools/diag_timeline.py cfg5 0 1 > $O/x2_timeline_T1.log 2>&1
python tools/diag_timeline.py cfg5 0 2 > $O/x2_timeline_T2.log 2>&1
python bench.py --impl reference > $O/r02b_bench_reference.json 2>/dev/null
tail -c 700 $O/r02b_bench_reference.json
grep -v "^$" $O/x2_timeline_T1.log | tail -14
