#!/bin/bash
# Round-end style visit: all GPU tests, smoke, default bench + reference arm, clip bench + its reference arm,
# ncu launch lists of both workloads, ncu --set full of the conditioning / output-stage kernels.
TAG=${1:-final}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/${TAG}_env.txt 2>&1
nproc >> $OUT/${TAG}_env.txt
timeout 900 python -m pytest tests -x -q -m gpu > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/${TAG}_pytest.log; tail -3 $OUT/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" >> $OUT/${TAG}_smoke.log; tail -3 $OUT/${TAG}_smoke.log
timeout 600 python bench.py > $OUT/${TAG}_bench_fp32.json 2> $OUT/${TAG}_bench_fp32.err; head -c 300 $OUT/${TAG}_bench_fp32.json; echo
timeout 300 python bench.py --precision bf16 --no-cpu-baseline --steps 50 > $OUT/${TAG}_bench_bf16.json 2> $OUT/${TAG}_bench_bf16.err; head -c 200 $OUT/${TAG}_bench_bf16.json; echo
timeout 300 python bench.py --workload clip > $OUT/${TAG}_bench_clip_bf16.json 2> $OUT/${TAG}_bench_clip_bf16.err; head -c 300 $OUT/${TAG}_bench_clip_bf16.json; echo
timeout 300 python bench.py --workload clip --precision fp32 --no-cpu-baseline > $OUT/${TAG}_bench_clip_fp32.json 2> $OUT/${TAG}_bench_clip_fp32.err; head -c 200 $OUT/${TAG}_bench_clip_fp32.json; echo
timeout 300 python bench.py --impl reference > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err; head -c 200 $OUT/${TAG}_bench_ref.json; echo
timeout 300 python bench.py --impl reference --workload clip > $OUT/${TAG}_bench_clip_ref.json 2> $OUT/${TAG}_bench_clip_ref.err; head -c 200 $OUT/${TAG}_bench_clip_ref.json; echo
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${TAG}_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_list.log 2>&1
echo "ncu list rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 500 --csv \
  --log-file $OUT/${TAG}_clip_launches.csv python bench.py --workload clip --frames 128 --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_clip_list.log 2>&1
echo "ncu clip list rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'delaunay|raster|draw_kernel|compose' --launch-skip 8 --launch-count 4 \
  -o $OUT/${TAG}_cond_full -f python bench.py --workload clip --frames 128 --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_cond_full.log 2>&1
echo "ncu full rc=$?"
