#!/bin/bash
one() { RDSP_BENCH_NO_CLOCKS=1 python bench.py --steps 40 --warmup 8 --no-cpu "${@:2}" 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['value']), round(d['ms_per_step'],4), {k:round(v['ms_per_launch'],4) for k,v in d['kernels'].items() if 'spec' in k or 'biquad' in k})"; }
for f in gpurun_lib_*.so; do RDSP_GPU_LIB=$PWD/$f one $f; done
