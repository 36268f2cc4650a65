#!/bin/bash
# netF pipelined under the generator: parity of the clip renderer, then the clip bench with netF in the loop.
TAG=${1:-flow3}
OUT=gpurun_out
mkdir -p $OUT
timeout 400 python -m pytest tests/test_flownet.py tests/test_clip.py -x -q -m gpu > $OUT/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -5 $OUT/${TAG}_pytest.log
timeout 400 python bench.py --workload clip --steps 3 --warmup 3 --flow-net 32,2,4,batch --no-cpu-baseline > $OUT/${TAG}_clip_bf16_netF.json 2> $OUT/${TAG}_clip_bf16_netF.err
tail -c 2200 $OUT/${TAG}_clip_bf16_netF.json
