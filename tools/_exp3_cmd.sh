#!/bin/bash
# two GPUs of one box: the two-devices-in-one-process test and the driver's launch line at N = 2
O=gpurun_out; mkdir -p $O
python -m pytest tests -m gpu -x -q -k "two_devices or sharding" > $O/r02b_two_devices.log 2>&1; tail -3 $O/r02b_two_devices.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-other-configs > $O/r02b_scale_cfg5_N2.json 2> $O/r02b_scale_cfg5_N2.err
tail -c 300 $O/r02b_scale_cfg5_N2.err; python -c "
import json; d=json.loads(open('$O/r02b_scale_cfg5_N2.json').read().strip().splitlines()[-1]); print(d['n_gpus'], round(d['value']), d['ms_per_step'], 'e2e', round(d['e2e']['value']), d['e2e'].get('frac_of_ceiling'))"
