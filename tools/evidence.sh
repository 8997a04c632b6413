#!/bin/bash
# evidence pass of a round (GPU box): bench lines of every config + the reference arm + the ncu launch list.
# usage: bash tools/evidence.sh r01g     -> gpurun_out/r01g_*
R=${1:-rXX}; O=gpurun_out; mkdir -p $O
python bench.py > $O/${R}_bench_cfg5_T8.json 2> $O/${R}_bench_cfg5_T8.err
python bench.py --blocks-per-call 1 --no-cpu > $O/${R}_bench_cfg5_T1.json 2>/dev/null
python bench.py --blocks-per-call 32 --no-cpu > $O/${R}_bench_cfg5_T32.json 2>/dev/null
for w in cfg2 cfg3 cfg4a cfg4b; do python bench.py --workload $w --no-cpu > $O/${R}_bench_$w.json 2>/dev/null; done
python bench.py --impl reference > $O/${R}_bench_reference.json 2>/dev/null
for f in $O/${R}_bench_*.json; do python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1])
print('$f'.split('/')[-1], round(d['value'],1), d.get('ms_per_step'), 'e2e', round(d['e2e']['value'],1), (d.get('roofline') or {}).get('frac'), (d.get('cpu_baseline') or {}).get('value'), d.get('clocks',{}).get('sm_mhz'))
"; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${R}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > $O/${R}_ncu_launches.log 2>&1
tail -3 $O/${R}_launches.csv | cut -c1-200
