#!/bin/bash
# A/B of two builds of the library inside ONE gpurun call: tools/ab_bench.sh <lib-a.so> <lib-b.so> [bench args...]
one() { RDSP_GPU_LIB=$1 RDSP_BENCH_NO_CLOCKS=1 python bench.py --steps 30 --warmup 5 --no-cpu --no-other-configs "${@:2}" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$1'.split('/')[-1], sys.argv[1:], round(d['value']), 'MS/s', round(d['ms_per_step']*1e3,1), 'us/step; e2e', round(d['e2e']['value']), {k:round(v['ms_per_launch']*1e3,1) for k,v in d['kernels'].items()})" "${@:2}"; }
A=$1; B=$2; shift 2
one $A "$@"; one $B "$@"
