#!/bin/bash
# A/B of environment switches on one box: bash tools/bench_env.sh TAG "ENV=.." "ENV=.." ...   (fp32 bench, 30 steps each)
TAG=$1; shift
mkdir -p gpurun_out
i=0
for ENVS in "$@"; do
  env $ENVS timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_env$i.json 2> gpurun_out/${TAG}_env$i.err
  python - "$ENVS" gpurun_out/${TAG}_env$i.json <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[2]))
    c = d["roofline"]["classes_ms_per_step"]
    print(f"{sys.argv[1]:45s} {d['value']:8.1f} fps  {d['ms_per_step']:.3f} ms  clocks {d['clocks']['sm_mhz']}  " + " ".join(f"{k}={v:.3f}" for k, v in c.items()))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
  i=$((i+1))
done
