set -x
L=$PWD/radiodsp_sdr_rx_b200
(time python -m pytest tests -m gpu -x -q) > gpurun_out/t6_tests.log 2>&1
tail -5 gpurun_out/t6_tests.log
O=gpurun_out/t6_ab.log; : > $O
bash tools/ab_bench.sh "$L/librdsp_gpu_prev.so $L/librdsp_gpu.so $L/librdsp_gpu_prev.so $L/librdsp_gpu.so" >> $O 2>&1
bash tools/ab_bench.sh "$L/librdsp_gpu_prev.so $L/librdsp_gpu.so" --workload cfg3 >> $O 2>&1
bash tools/ab_bench.sh "$L/librdsp_gpu_prev.so $L/librdsp_gpu.so" --workload cfg4a >> $O 2>&1
cat $O
