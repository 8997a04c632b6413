#!/bin/bash
# the other BASELINE configs as full bench lines (N = 1)
O=gpurun_out; mkdir -p $O
for w in cfg2 cfg3 cfg4a cfg4b; do python bench.py --workload $w --no-cpu --no-other-configs > $O/r02b_bench_$w.json 2>/dev/null; done
for f in $O/r02b_bench_cfg[234]*.json; do python -c "
import json
d=json.loads(open('$f').read().strip().splitlines()[-1])
print('$f'.split('/')[-1], round(d['value'],1), d.get('ms_per_step'), 'e2e', round(d['e2e']['value'],1), d['roofline']['kernel'], round(d['roofline']['frac'],3), (d['roofline'].get('pipe') or {}).get('frac'), {k:round(v['ms_per_launch']*1e3,1) for k,v in d.get('kernels',{}).items()})
"; done
