#!/bin/bash
# packed k_nlms with two chunks per loop trip: GPU suite, A/B against the previous build
O=gpurun_out; mkdir -p $O
L=$PWD/radiodsp_sdr_rx_b200
(time python -m pytest tests -m gpu -x -q) > $O/x8_tests.log 2>&1; tail -4 $O/x8_tests.log
: > $O/x8_ab.log
bash tools/ab_bench.sh "$L/librdsp_gpu_prev.so $L/librdsp_gpu.so $L/librdsp_gpu_prev.so $L/librdsp_gpu.so $L/librdsp_gpu_prev.so $L/librdsp_gpu.so" >> $O/x8_ab.log 2>&1
cut -c1-330 $O/x8_ab.log
