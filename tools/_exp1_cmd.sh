#!/bin/bash
# class-balance experiment: GPU suite, timelines, RDSP_SIDE_TOPUP sweep
O=gpurun_out; mkdir -p $O
(time python -m pytest tests -m gpu -x -q) > $O/x1_tests.log 2>&1; tail -4 $O/x1_tests.log
one() { RDSP_BENCH_NO_CLOCKS=1 python bench.py --steps 30 --warmup 6 --no-cpu --no-other-configs "${@:2}" 2>$O/x1_err.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('$1', round(d['value']), round(d['ms_per_step']*1e3,1), d['ms_step_min_median_max'][1], {k:round(v['ms_per_launch']*1e3,1) for k,v in d['kernels'].items()})
except Exception as e: print('$1 FAILED', e); print(open('$O/x1_err.log').read()[-800:])"; }
for x in 0 160 320 480 640 1000 1500 -1 0 320; do RDSP_SIDE_TOPUP=$x one "topup=$x"; done 2>&1 | tee $O/x1_sweep.log
for a in 0 2 3; do RDSP_SPEC_AFTER=$a one "rule spec_after=$a"; done 2>&1 | tee -a $O/x1_sweep.log
RDSP_SIDE_TOPUP=0 python tools/diag_timeline.py cfg5 > $O/x1_timeline_0.log 2>&1
python tools/diag_timeline.py cfg5 > $O/x1_timeline_rule.log 2>&1
one "T1 rule" --blocks-per-call 1 | tee -a $O/x1_sweep.log
RDSP_SIDE_TOPUP=0 one "T1 topup=0" --blocks-per-call 1 | tee -a $O/x1_sweep.log
