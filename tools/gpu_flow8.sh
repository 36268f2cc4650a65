#!/bin/bash
TAG=${1:-flow8}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_flownet.py tests/test_clip.py -x -q -m gpu > $OUT/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 $OUT/${TAG}_pytest.log
timeout 300 python tools/clip_flow_diag.py > $OUT/${TAG}_diag.json 2>&1; tail -1 $OUT/${TAG}_diag.json
timeout 400 python bench.py --workload clip --steps 3 --warmup 3 --flow-net 32,2,4,batch --no-cpu-baseline > $OUT/${TAG}_clip_bf16_netF.json 2> $OUT/${TAG}_clip_bf16_netF.err
python tools/oneline.py $OUT/${TAG}_clip_bf16_netF.json
