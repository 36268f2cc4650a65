#!/bin/bash
# netF follow-up visit: parity, probe, one --set full capture (with source) of the register-tiled conv.
TAG=${1:-flow2}
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python -m pytest tests/test_flownet.py -x -q -m gpu > $OUT/${TAG}_pytest_flow.log 2>&1
echo "flow pytest rc=$?"; tail -5 $OUT/${TAG}_pytest_flow.log
timeout 120 python tools/flow_probe.py > $OUT/${TAG}_probe.json 2> $OUT/${TAG}_probe.err; cat $OUT/${TAG}_probe.json
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $OUT/${TAG}_flow_launches.csv \
  python tools/flow_probe.py --once > $OUT/${TAG}_flow_list.log 2>&1
echo "ncu list rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:fconv_tiled_kernel --launch-skip 1 --launch-count 2 \
  -o $OUT/${TAG}_tiled -f python tools/flow_probe.py --once > $OUT/${TAG}_tiled.log 2>&1
echo "ncu full rc=$?"
ls -la $OUT/${TAG}_*
