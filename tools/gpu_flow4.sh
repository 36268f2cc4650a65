#!/bin/bash
# netF tensor-core convs: parity, probe (umma on/off), launch list.
TAG=${1:-flow4}
OUT=gpurun_out
mkdir -p $OUT
timeout 400 python -m pytest tests/test_flownet.py -x -q -m gpu > $OUT/${TAG}_pytest_flow.log 2>&1
echo "flow pytest rc=$?"; tail -25 $OUT/${TAG}_pytest_flow.log
timeout 120 python tools/flow_probe.py > $OUT/${TAG}_probe.json 2> $OUT/${TAG}_probe.err; cat $OUT/${TAG}_probe.json; tail -3 $OUT/${TAG}_probe.err
AP_FLOW_UMMA=0 timeout 120 python tools/flow_probe.py --reps 3 > $OUT/${TAG}_probe_ffma.json 2>> $OUT/${TAG}_probe.err; cat $OUT/${TAG}_probe_ffma.json
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active --clock-control none --csv --log-file $OUT/${TAG}_flow_launches.csv \
  python tools/flow_probe.py --once > $OUT/${TAG}_flow_list.log 2>&1
echo "ncu list rc=$?"
