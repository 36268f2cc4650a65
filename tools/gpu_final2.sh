#!/bin/bash
# Final visit of the round: every GPU test, smoke, the bench lines of the final build.
TAG=${1:-final}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -x -q -m gpu > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/${TAG}_pytest.log; tail -3 $OUT/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" >> $OUT/${TAG}_smoke.log; tail -5 $OUT/${TAG}_smoke.log
timeout 600 python bench.py > $OUT/${TAG}_bench_fp32.json 2> $OUT/${TAG}_bench_fp32.err
timeout 300 python bench.py --precision bf16 --no-cpu-baseline > $OUT/${TAG}_bench_bf16.json 2> $OUT/${TAG}_bench_bf16.err
timeout 300 python bench.py --workload clip --no-cpu-baseline > $OUT/${TAG}_clip_bf16.json 2> $OUT/${TAG}_clip_bf16.err
timeout 300 python bench.py --workload clip --steps 3 --warmup 3 --flow-net 32,2,4,batch --no-cpu-baseline > $OUT/${TAG}_clip_bf16_netF.json 2> $OUT/${TAG}_clip_bf16_netF.err
timeout 300 python bench.py --workload clip --steps 2 --warmup 2 --precision fp32 --flow-net 32,2,4,batch --no-cpu-baseline > $OUT/${TAG}_clip_fp32_netF.json 2> $OUT/${TAG}_clip_fp32_netF.err
python tools/oneline.py $OUT/${TAG}_bench_fp32.json $OUT/${TAG}_bench_bf16.json $OUT/${TAG}_clip_bf16.json $OUT/${TAG}_clip_bf16_netF.json $OUT/${TAG}_clip_fp32_netF.json
