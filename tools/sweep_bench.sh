#!/bin/bash
# development sweep: one line per configuration (GPU box)
run() { python bench.py --steps 15 --warmup 4 --no-cpu "$@" 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$LABEL', 'MS/s',round(d['value']),'ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']), {k:round(v['ms_per_launch'],4) for k,v in d['kernels'].items()})"; }
LABEL="cfg5" run
LABEL="cfg5 T=1" run --blocks-per-call 1
LABEL="cfg5 T=32" run --blocks-per-call 32
for w in cfg2 cfg3 cfg4a cfg4b; do LABEL="$w" run --workload $w; done
