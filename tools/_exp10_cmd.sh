#!/bin/bash
# one block per call, end to end: staging depth (RDSP_HOST_STAGES) and the nvidia-smi sampler
O=gpurun_out; mkdir -p $O
one() { python bench.py --steps 200 --warmup 20 --no-cpu --no-other-configs --blocks-per-call 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$1', round(d['value']), round(d['ms_per_step']*1e3,1), 'e2e', round(d['e2e']['value']), round(d['e2e']['frac_of_ceiling'],3), 'mono', round(d['e2e_mono']['value']), round(d['e2e_mono']['frac_of_ceiling'],3))"; }
for s in 2 3 4; do RDSP_HOST_STAGES=$s one "stages=$s"; RDSP_HOST_STAGES=$s RDSP_BENCH_NO_CLOCKS=1 one "stages=$s no-sampler"; done 2>&1 | tee $O/x10.log
