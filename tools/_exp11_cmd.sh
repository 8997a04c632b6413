#!/bin/bash
# the front end alone (cfg2) at channel counts that fill the 148 SMs with whole tiles: what the fused streaming + FIR kernel reaches against HBM
O=gpurun_out; mkdir -p $O
for c in 4096 9472 18944 37888; do RDSP_BENCH_NO_CLOCKS=1 python bench.py --workload cfg2 --channels $c --steps 30 --warmup 6 --no-cpu --no-other-configs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernels']['k_front']; print('cfg2 channels', $c, 'MS/s', round(d['value']), 'us/step', round(d['ms_per_step']*1e3,1), 'k_front us', round(k['ms_per_launch']*1e3,1), 'alg GB/s', round(k['alg_gb_s']), 'hbm frac', round(d['roofline']['frac'],3), 'pipe', d['roofline'].get('pipe'))"; done 2>&1 | tee $O/x11.log
