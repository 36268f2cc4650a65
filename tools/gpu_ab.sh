#!/bin/bash
# A/B of one env switch on the default bench (fp32-accurate, B=16) and the clip bench (bf16), back to back on one box,
# plus the GPU test-suite and an ncu launch list of a small clip.   gpurun -- 'bash tools/gpu_ab.sh TAG ENVVAR'
TAG=${1:-ab}; VAR=${2:-AP_NETG_APPLY_DEEP}
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -x -q -m gpu > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/${TAG}_pytest.log; tail -4 $OUT/${TAG}_pytest.log
for v in 0 1 0 1; do
  env $VAR=$v timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > $OUT/${TAG}_fp32_$v.json 2> $OUT/${TAG}_fp32_$v.err
  python - <<PY
import json; d=json.load(open("$OUT/${TAG}_fp32_$v.json")); print("fp32 $VAR=$v", round(d["value"],1), round(d["e2e"]["value"],1), d["roofline"]["classes_ms_per_step"], d["clocks"]["sm_mhz"])
PY
done
for v in 0 1; do
  env $VAR=$v timeout 300 python bench.py --workload clip --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_clip_$v.json 2> $OUT/${TAG}_clip_$v.err
  python - <<PY
import json; d=json.load(open("$OUT/${TAG}_clip_$v.json")); print("clip $VAR=$v", round(d["value"],1), round(d["e2e"]["value"],1), d["roofline"]["classes_ms_per_batch"], d["roofline"]["conditioning_ms_per_batch"])
PY
done
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 500 --csv \
  --log-file $OUT/${TAG}_clip_launches.csv python bench.py --workload clip --frames 128 --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_clip_list.log 2>&1
echo "ncu list rc=$?"
