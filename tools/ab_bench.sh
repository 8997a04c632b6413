#!/bin/bash
# compact bench lines of one or more builds of the library inside ONE gpurun call:
#   tools/ab_bench.sh "<lib.so> [<lib.so> ...]" [bench args...]      (RDSP_GPU_LIB selects the build; same ABI required)
LIBS=$1; shift
for L in $LIBS; do
  RDSP_GPU_LIB=$L RDSP_BENCH_NO_CLOCKS=1 python bench.py --steps 30 --warmup 5 --no-cpu --no-other-configs "$@" 2>gpurun_out/ab_err.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read())
    print('$L'.split('/')[-1], sys.argv[1:], round(d['value']), 'MS/s', round(d['ms_per_step']*1e3,1), 'us/step; e2e', round(d['e2e']['value']), 'mono', round(d['e2e_mono']['value']), {k:round(v['ms_per_launch']*1e3,1) for k,v in d['kernels'].items()})
except Exception as e:
    print('$L', 'FAILED', e); print(open('gpurun_out/ab_err.log').read()[-1500:])" "$@"
done
