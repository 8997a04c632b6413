#!/bin/bash
# final state check of the round: GPU suite, smoke, default bench line, 1 / 32 blocks per call, launch list
R=${1:-rXX}; O=gpurun_out; mkdir -p $O
(time python -m pytest tests -m gpu -x -q) > $O/${R}_gputests.log 2>&1; tail -4 $O/${R}_gputests.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/${R}_smoke.log 2>&1; tail -2 $O/${R}_smoke.log
python bench.py > $O/${R}_bench_cfg5.json 2> $O/${R}_bench_cfg5.err; tail -c 300 $O/${R}_bench_cfg5.err
python bench.py --blocks-per-call 1 --steps 200 --warmup 20 --no-cpu --no-other-configs > $O/${R}_bench_cfg5_T1.json 2>/dev/null
python bench.py --blocks-per-call 32 --no-cpu --no-other-configs > $O/${R}_bench_cfg5_T32.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${R}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-other-configs > $O/${R}_ncu_launches.log 2>&1
for f in $O/${R}_bench_*.json; do python -c "
import json
d=json.loads(open('$f').read().strip().splitlines()[-1])
print('$f'.split('/')[-1], round(d['value'],1), d.get('ms_per_step'), 'e2e', round(d['e2e']['value'],1), round(d['e2e']['frac_of_ceiling'],3), 'mono', round(d['e2e_mono']['value']), (d.get('roofline') or {}).get('frac'), {k:(round(v['value']),v['ms_per_step']) for k,v in (d.get('other_configs') or {}).items()})
"; done
