#!/bin/bash
# k_nlms with the normalisers in the pre-pass: GPU suite, then A/B against the previous build (same ABI, RDSP_GPU_LIB)
O=gpurun_out; mkdir -p $O
L=$PWD/radiodsp_sdr_rx_b200
(time python -m pytest tests -m gpu -x -q) > $O/x7_tests.log 2>&1; tail -15 $O/x7_tests.log
: > $O/x7_ab.log
bash tools/ab_bench.sh "$L/librdsp_gpu_prev.so $L/librdsp_gpu.so $L/librdsp_gpu_prev.so $L/librdsp_gpu.so" >> $O/x7_ab.log 2>&1
bash tools/ab_bench.sh "$L/librdsp_gpu_prev.so $L/librdsp_gpu.so" --workload cfg3 >> $O/x7_ab.log 2>&1
bash tools/ab_bench.sh "$L/librdsp_gpu_prev.so $L/librdsp_gpu.so" --workload cfg4a >> $O/x7_ab.log 2>&1
bash tools/ab_bench.sh "$L/librdsp_gpu_prev.so $L/librdsp_gpu.so" --blocks-per-call 1 >> $O/x7_ab.log 2>&1
cat $O/x7_ab.log
