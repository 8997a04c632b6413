set -x
L=$PWD/radiodsp_sdr_rx_b200
(time python -m pytest tests -m gpu -x -q) > gpurun_out/t2_tests.log 2>&1
tail -5 gpurun_out/t2_tests.log
O=gpurun_out/t2_ab.log; : > $O
bash tools/ab_bench.sh "$L/librdsp_gpu_prev.so $L/librdsp_gpu.so" >> $O 2>&1
for v in 0 2; do echo "SPEC_AFTER=$v" >> $O; RDSP_SPEC_AFTER=$v bash tools/ab_bench.sh "$L/librdsp_gpu.so" >> $O 2>&1; done
echo "NO_PRIO" >> $O; RDSP_NO_PRIO=1 bash tools/ab_bench.sh "$L/librdsp_gpu.so" >> $O 2>&1
echo "chunks2" >> $O; bash tools/ab_bench.sh "$L/librdsp_gpu.so" --pipeline-chunks 2 >> $O 2>&1
echo "LANES=4 (cfg5)" >> $O; RDSP_NLMS_LANES=4 bash tools/ab_bench.sh "$L/librdsp_gpu.so" >> $O 2>&1
for w in cfg3 cfg4a; do
  bash tools/ab_bench.sh "$L/librdsp_gpu.so" --workload $w >> $O 2>&1
  for pk in 0 1; do echo "LANES=8 PACKED=$pk" >> $O; RDSP_NLMS_LANES=8 RDSP_NLMS_PACKED=$pk bash tools/ab_bench.sh "$L/librdsp_gpu.so" --workload $w >> $O 2>&1; done
done
bash tools/ab_bench.sh "$L/librdsp_gpu_prev.so $L/librdsp_gpu.so" --workload cfg4b >> $O 2>&1
cat $O
