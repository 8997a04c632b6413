#!/bin/bash
# ncu --set full of one un-profiled cfg5 step (third call of tools/prof_one.py, graphs off) + a source-level capture of k_nlms
R=${1:-rXX}; O=gpurun_out; mkdir -p $O
python tools/prof_one.py cfg5 8192 3 > $O/${R}_prof_plain.log 2>&1 || { tail -5 $O/${R}_prof_plain.log; exit 1; }
ncu --set full --clock-control none --launch-skip 24 --launch-count 12 -f -o $O/${R}_full python tools/prof_one.py cfg5 8192 3 > $O/${R}_ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_nlms --launch-skip 4 --launch-count 2 -f -o $O/${R}_nlms_src python tools/prof_one.py cfg5 8192 3 > $O/${R}_ncu_nlms.log 2>&1
ls -la $O/*.ncu-rep; tail -2 $O/${R}_ncu_full.log $O/${R}_ncu_nlms.log
