#!/bin/bash
# netF visit: parity tests, probe timings (default / fp32 kernels only / generic kernel), ncu launch list, clip bench with
# netF in the loop, clip diagnosis.
#   gpurun --timeout 700 -- 'bash tools/gpu_flow.sh TAG'
TAG=${1:-flow}
OUT=gpurun_out
mkdir -p $OUT
timeout 400 python -m pytest tests/test_flownet.py tests/test_clip.py tests/test_conditioning.py -x -q -m gpu > $OUT/${TAG}_pytest_flow.log 2>&1
echo "flow pytest rc=$?"; tail -5 $OUT/${TAG}_pytest_flow.log
timeout 120 python tools/flow_probe.py > $OUT/${TAG}_probe.json 2> $OUT/${TAG}_probe.err; cat $OUT/${TAG}_probe.json
AP_FLOW_UMMA=0 timeout 120 python tools/flow_probe.py --reps 3 > $OUT/${TAG}_probe_fp32_kernels.json 2>> $OUT/${TAG}_probe.err; cat $OUT/${TAG}_probe_fp32_kernels.json
AP_FLOW_UMMA=0 AP_FLOW_SPARSE=0 AP_FLOW_TILED=0 timeout 120 python tools/flow_probe.py --reps 2 > $OUT/${TAG}_probe_generic.json 2>> $OUT/${TAG}_probe.err; cat $OUT/${TAG}_probe_generic.json
timeout 120 python tools/flow_probe.py --batch 1 --reps 50 > $OUT/${TAG}_probe_b1.json 2>> $OUT/${TAG}_probe.err; cat $OUT/${TAG}_probe_b1.json
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active \
  --clock-control none --csv --log-file $OUT/${TAG}_flow_launches.csv python tools/flow_probe.py --once > $OUT/${TAG}_flow_list.log 2>&1
echo "ncu rc=$?"
timeout 400 python bench.py --workload clip --steps 3 --warmup 3 --flow-net 32,2,4,batch --no-cpu-baseline > $OUT/${TAG}_clip_bf16_netF.json 2> $OUT/${TAG}_clip_bf16_netF.err
python tools/oneline.py $OUT/${TAG}_clip_bf16_netF.json
timeout 300 python tools/clip_flow_diag.py > $OUT/${TAG}_diag.json 2>&1; tail -1 $OUT/${TAG}_diag.json
