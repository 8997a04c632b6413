#!/bin/bash
# one bench line per invocation at N GPUs of this box: tools/run_scale.sh <N> <tag> [bench args...]   (driver's launch line)
N=$1; TAG=$2; shift 2
mkdir -p gpurun_out/scale
if [ "$N" = "1" ]; then python bench.py --gpus 1 "$@" > gpurun_out/scale/${TAG}_N1.json 2> gpurun_out/scale/${TAG}_N1.err
else python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N "$@" > gpurun_out/scale/${TAG}_N$N.json 2> gpurun_out/scale/${TAG}_N$N.err; fi
python - <<PY
import json
try:
    d=[json.loads(l) for l in open("gpurun_out/scale/${TAG}_N$N.json") if l.startswith("{")][-1]
    print("${TAG} N=$N", round(d["value"]), "MS/s", round(d["ms_per_step"],4), "ms/step; e2e", round(d["e2e"]["value"]), "ceiling", round(d["e2e"]["ceiling"]["value"]), "frac", round(d["e2e"]["frac_of_ceiling"],3), "| mono", round(d["e2e_mono"]["value"]), "ceiling", round(d["e2e_mono"]["ceiling"]["value"]), "| replays", d.get("graph_replays"))
except Exception as e:
    print("${TAG} N=$N FAILED", e); print(open("gpurun_out/scale/${TAG}_N$N.err").read()[-2000:])
PY
