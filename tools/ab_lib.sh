#!/bin/bash
one() { RDSP_BENCH_NO_CLOCKS=1 python bench.py --steps 60 --warmup 10 --no-cpu "${@:2}" 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['value']), round(d['ms_per_step'],4), [round(x,3) for x in d['ms_step_min_median_max']])"; }
one "G=1"
one "G=2" --pipeline-chunks 2
one "G=3" --pipeline-chunks 3
one "G=4" --pipeline-chunks 4
