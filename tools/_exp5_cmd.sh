#!/bin/bash
# channel groups (pipeline_chunks) on the configs whose chain is serial (no spectrum branch beside it)
O=gpurun_out; mkdir -p $O
one() { RDSP_BENCH_NO_CLOCKS=1 python bench.py --steps 30 --warmup 6 --no-cpu --no-other-configs "${@:2}" 2>$O/x5_err.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('$1', round(d['value']), round(d['ms_per_step']*1e3,1), 'e2e', round(d['e2e']['value']))
except Exception as e: print('$1 FAILED', e); print(open('$O/x5_err.log').read()[-800:])"; }
for w in cfg3 cfg4a cfg4b cfg2; do for g in 1 2 3 4; do one "$w G=$g" --workload $w --pipeline-chunks $g; done; done 2>&1 | tee $O/x5_sweep.log
