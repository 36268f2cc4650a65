#!/bin/bash
TAG=${1:-flow7}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_flownet.py tests/test_clip.py tests/test_conditioning.py -x -q -m gpu > $OUT/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -15 $OUT/${TAG}_pytest.log
timeout 120 python tools/flow_probe.py > $OUT/${TAG}_probe.json 2> $OUT/${TAG}_probe.err; cat $OUT/${TAG}_probe.json; tail -3 $OUT/${TAG}_probe.err
timeout 300 python tools/clip_flow_diag.py > $OUT/${TAG}_diag.json 2>&1; tail -1 $OUT/${TAG}_diag.json
