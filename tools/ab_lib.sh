#!/bin/bash
one() { RDSP_BENCH_NO_CLOCKS=1 python bench.py --steps 48 --warmup 10 --no-cpu "${@:2}" 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['value']), round(d['ms_per_step'],4), [round(x,3) for x in d['ms_step_min_median_max']], {k:round(v['ms_per_launch'],4) for k,v in d['kernels'].items()})"; }
RDSP_CARVEOUT=-1 one "carveout default G1"
RDSP_CARVEOUT=100 one "carveout 100 G1"
RDSP_CARVEOUT=50 one "carveout 50 G1"
RDSP_CARVEOUT=100 one "carveout 100 G2" --pipeline-chunks 2
RDSP_CARVEOUT=100 one "carveout 100 G4" --pipeline-chunks 4
RDSP_CARVEOUT=-1 one "carveout default G4" --pipeline-chunks 4
