set -x
L=$PWD/radiodsp_sdr_rx_b200
O=gpurun_out/t4_ab.log; : > $O
bash tools/ab_bench.sh "$L/librdsp_gpu.so $L/librdsp_gpu_mb5.so $L/librdsp_gpu_mb6.so $L/librdsp_gpu.so" >> $O 2>&1
bash tools/ab_bench.sh "$L/librdsp_gpu.so $L/librdsp_gpu_mb5.so $L/librdsp_gpu_mb6.so" --workload cfg4a >> $O 2>&1
bash tools/ab_bench.sh "$L/librdsp_gpu_ff1.so $L/librdsp_gpu_ff2.so $L/librdsp_gpu_ff3.so $L/librdsp_gpu_ff4.so $L/librdsp_gpu.so" >> $O 2>&1
bash tools/ab_bench.sh "$L/librdsp_gpu_ff1.so $L/librdsp_gpu_ff3.so $L/librdsp_gpu_ff4.so $L/librdsp_gpu.so" --workload cfg4b >> $O 2>&1
bash tools/ab_bench.sh "$L/librdsp_gpu.so" --workload cfg2 >> $O 2>&1
bash tools/ab_bench.sh "$L/librdsp_gpu.so" --blocks-per-call 1 >> $O 2>&1
cat $O
python tools/diag_timeline.py cfg5 0 1 > gpurun_out/t4_timeline_T1.log 2>&1
tail -14 gpurun_out/t4_timeline_T1.log
python tools/prof_one.py cfg5 8192 2 > gpurun_out/t4_prof.log 2>&1 || exit 1
timeout 500 ncu --set full --clock-control none --import-source on --launch-skip 13 -c 13 -f -o gpurun_out/t4_step python tools/prof_one.py cfg5 8192 2 > gpurun_out/t4_ncu.log 2>&1
tail -3 gpurun_out/t4_ncu.log; ls -la gpurun_out/
