#!/bin/bash
one() { RDSP_BENCH_NO_CLOCKS=1 python bench.py --steps 60 --warmup 10 --no-cpu "${@:2}" 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['value']), round(d['ms_per_step'],4), [round(x,3) for x in d['ms_step_min_median_max']])"; }
one "default (spec after front)"
RDSP_SPEC_WITH_FRONT=1 one "spec with front"
one "default (spec after front)"
one "cfg5 T=32" --blocks-per-call 32
one "cfg5 T=1" --blocks-per-call 1
