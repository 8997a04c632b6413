set -x
L=$PWD/radiodsp_sdr_rx_b200
(time python -m pytest tests -m gpu -x -q -k "spec256 or full_chain or config") > gpurun_out/t7_tests.log 2>&1
tail -5 gpurun_out/t7_tests.log
O=gpurun_out/t7_ab.log; : > $O
bash tools/ab_bench.sh "$L/librdsp_gpu_prev.so $L/librdsp_gpu.so $L/librdsp_gpu_s256_3.so $L/librdsp_gpu_s256_5.so $L/librdsp_gpu_s1024_10.so $L/librdsp_gpu_s1024_14.so $L/librdsp_gpu_s1024_16.so $L/librdsp_gpu_prev.so $L/librdsp_gpu.so" >> $O 2>&1
cat $O
