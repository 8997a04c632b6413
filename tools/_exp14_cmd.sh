#!/bin/bash
O=gpurun_out; mkdir -p $O
(time python -m pytest tests -m gpu -x -q) > $O/r02f_gputests.log 2>&1; tail -4 $O/r02f_gputests.log
python bench.py --blocks-per-call 1 --steps 200 --warmup 20 --no-cpu --no-other-configs > $O/r02f_bench_cfg5_T1.json 2>/dev/null
python bench.py --no-cpu --no-other-configs > $O/r02f_bench_cfg5_quick.json 2>/dev/null
for f in $O/r02f_bench_*.json; do python -c "
import json
d=json.loads(open('$f').read().strip().splitlines()[-1])
print('$f'.split('/')[-1], round(d['value'],1), d.get('ms_per_step'), 'e2e', round(d['e2e']['value'],1), round(d['e2e']['frac_of_ceiling'],3), 'mono', round(d['e2e_mono']['value']))
"; done
