#!/bin/bash
# one-block calls with the audio row appended by the emitting kernels: GPU suite, A/B against the previous build at 1 and 8 blocks per call
O=gpurun_out; mkdir -p $O
L=$PWD/radiodsp_sdr_rx_b200
(time python -m pytest tests -m gpu -x -q) > $O/x12_tests.log 2>&1; tail -4 $O/x12_tests.log
: > $O/x12_ab.log
bash tools/ab_bench.sh "$L/librdsp_gpu_prev.so $L/librdsp_gpu.so $L/librdsp_gpu_prev.so $L/librdsp_gpu.so" --blocks-per-call 1 >> $O/x12_ab.log 2>&1
bash tools/ab_bench.sh "$L/librdsp_gpu_prev.so $L/librdsp_gpu.so" >> $O/x12_ab.log 2>&1
cut -c1-330 $O/x12_ab.log
