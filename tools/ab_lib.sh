#!/bin/bash
one() { RDSP_BENCH_NO_CLOCKS=1 python bench.py --steps 48 --warmup 10 --no-cpu "${@:2}" 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['value']), round(d['ms_per_step'],4), [round(x,3) for x in d['ms_step_min_median_max']], {k:round(v['ms_per_launch'],4) for k,v in d['kernels'].items()})"; }
RDSP_PDL=0 one "pdl off"
RDSP_PDL=1 one "pdl on"
RDSP_PDL=0 one "pdl off"
RDSP_PDL=1 one "pdl on"
RDSP_PDL=1 one "pdl on cfg4a" --workload cfg4a
RDSP_PDL=0 one "pdl off cfg4a" --workload cfg4a
