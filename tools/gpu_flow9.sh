#!/bin/bash
TAG=${1:-flow9}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_flownet.py -x -q -m gpu > $OUT/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 $OUT/${TAG}_pytest.log
timeout 120 python tools/flow_probe.py > $OUT/${TAG}_probe.json 2> $OUT/${TAG}_probe.err; cat $OUT/${TAG}_probe.json; tail -3 $OUT/${TAG}_probe.err
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum --clock-control none --csv --log-file $OUT/${TAG}_flow_launches.csv \
  python tools/flow_probe.py --once > $OUT/${TAG}_flow_list.log 2>&1
echo "ncu rc=$?"
