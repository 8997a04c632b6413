#!/bin/bash
# compute-sanitizer memcheck over the CUDA-core kernels (the tcgen05 front end is swapped for its CUDA-core cross-check kernel:
# RDSP_FRONT_IMPL=cuda-core), ragged channel counts and every stage; bounded by timeout
O=gpurun_out; mkdir -p $O
export RDSP_FRONT_IMPL=cuda-core RDSP_GRAPH=0
timeout 420 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 \
  python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ragged_channel_counts or spec256_bit_exact or spec1024_bit_exact or notch_and_agc_f32_parity or conv_and_nr_kinds or blocks_per_call_invariance" > $O/r02b_memcheck.log 2>&1
echo "exit code $?" >> $O/r02b_memcheck.log
grep -E "ERROR SUMMARY|passed|failed|exit code|Invalid|out of bounds" $O/r02b_memcheck.log | head -20
unset RDSP_FRONT_IMPL RDSP_GRAPH
python bench.py --blocks-per-call 1 --no-cpu --no-other-configs > $O/r02b_bench_cfg5_T1.json 2>/dev/null
python -c "
import json; d=json.loads(open('$O/r02b_bench_cfg5_T1.json').read().strip().splitlines()[-1]); print('T1', round(d['value']), d['ms_per_step'], 'e2e', round(d['e2e']['value']), d['e2e']['frac_of_ceiling'], 'mono', round(d['e2e_mono']['value']))"
