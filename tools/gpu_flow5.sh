#!/bin/bash
TAG=${1:-flow5}
OUT=gpurun_out
mkdir -p $OUT
timeout 400 python -m pytest tests/test_flownet.py -x -q -m gpu > $OUT/${TAG}_pytest_flow.log 2>&1
echo "flow pytest rc=$?"; tail -5 $OUT/${TAG}_pytest_flow.log
timeout 120 python tools/flow_probe.py > $OUT/${TAG}_probe.json 2> $OUT/${TAG}_probe.err; cat $OUT/${TAG}_probe.json; tail -3 $OUT/${TAG}_probe.err
AP_FLOW_BN256=1 timeout 120 python tools/flow_probe.py > $OUT/${TAG}_probe_bn256.json 2>> $OUT/${TAG}_probe.err; cat $OUT/${TAG}_probe_bn256.json
AP_FLOW_BN256=1 timeout 300 python -m pytest tests/test_flownet.py -x -q -m gpu -k "fast_kernels or matches_the_oracle" > $OUT/${TAG}_pytest_bn256.log 2>&1
echo "bn256 pytest rc=$?"; tail -3 $OUT/${TAG}_pytest_bn256.log
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active --clock-control none --csv --log-file $OUT/${TAG}_flow_launches.csv \
  python tools/flow_probe.py --once > $OUT/${TAG}_flow_list.log 2>&1
AP_FLOW_BN256=1 timeout 300 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active --clock-control none --csv --log-file $OUT/${TAG}_flow_launches_bn256.csv \
  python tools/flow_probe.py --once > $OUT/${TAG}_flow_list2.log 2>&1
echo "ncu list rc=$?"
