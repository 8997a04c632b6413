#!/bin/bash
one() { RDSP_BENCH_NO_CLOCKS=1 python bench.py --steps 40 --warmup 8 --no-cpu "${@:2}" 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['value']), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), [round(x,3) for x in d['ms_step_min_median_max']])"; }
for T in 4 8 12 16 24 32 64; do one "T=$T" --blocks-per-call $T; done
