#!/bin/bash
TAG=${1:-flow6}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests -x -q -m gpu > $OUT/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 $OUT/${TAG}_pytest.log
timeout 120 python tools/flow_probe.py > $OUT/${TAG}_probe.json 2> $OUT/${TAG}_probe.err; cat $OUT/${TAG}_probe.json
timeout 120 python tools/flow_probe.py --batch 1 --reps 50 > $OUT/${TAG}_probe_b1.json 2>> $OUT/${TAG}_probe.err; cat $OUT/${TAG}_probe_b1.json
timeout 400 python bench.py --workload clip --steps 3 --warmup 3 --flow-net 32,2,4,batch --no-cpu-baseline > $OUT/${TAG}_clip_bf16_netF.json 2> $OUT/${TAG}_clip_bf16_netF.err
python - <<'PY'
import json
for l in open('gpurun_out/'+"${TAG}"+'_clip_bf16_netF.json'):
    if l.startswith('{'):
        d=json.loads(l); print('clip+netF', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['roofline'].get('conditioning_ms_per_batch'))
PY
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active --clock-control none --csv --log-file $OUT/${TAG}_flow_launches.csv \
  python tools/flow_probe.py --once > $OUT/${TAG}_flow_list.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:fconv_umma_kernel --launch-skip 1 --launch-count 9 -o /tmp/${TAG}_umma -f python tools/flow_probe.py --once > $OUT/${TAG}_umma_full.log 2>&1
python tools/ncu_summary.py /tmp/${TAG}_umma.ncu-rep > $OUT/${TAG}_umma_full_summary.txt 2>&1
tail -12 $OUT/${TAG}_umma_full_summary.txt
