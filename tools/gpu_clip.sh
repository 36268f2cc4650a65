#!/bin/bash
# GPU-box visit for the conditioning producers + clip workload: tests, clip bench lines, ncu launch list of one small clip.
#   gpurun --timeout 700 -- 'bash tools/gpu_clip.sh TAG [stages...]'      stages: test bench batch64 ref list sanitize
TAG=${1:-clip}
shift
STAGES=${*:-test bench list}
OUT=gpurun_out
mkdir -p $OUT
has() { [[ " $STAGES " == *" $1 "* ]]; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/${TAG}_env.txt 2>&1
if has test; then
  timeout 500 python -m pytest tests/test_conditioning.py tests/test_clip.py -x -q -m gpu > $OUT/${TAG}_pytest.log 2>&1
  echo "pytest rc=$?" >> $OUT/${TAG}_pytest.log
  tail -15 $OUT/${TAG}_pytest.log
fi
if has sanitize; then
  timeout 300 compute-sanitizer --error-exitcode 9 python -m pytest tests/test_conditioning.py -x -q -m gpu -k "not full_clip" > $OUT/${TAG}_sanitize.log 2>&1
  echo "sanitize rc=$?" >> $OUT/${TAG}_sanitize.log
  tail -5 $OUT/${TAG}_sanitize.log
fi
if has bench; then
  timeout 300 python bench.py --workload clip --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_clip_bf16.json 2> $OUT/${TAG}_bench_clip_bf16.err
  tail -c 2500 $OUT/${TAG}_bench_clip_bf16.json; tail -3 $OUT/${TAG}_bench_clip_bf16.err
fi
if has ab; then
  timeout 300 python bench.py --workload clip --steps 3 --warmup 3 --no-share-photo --no-cpu-baseline > $OUT/${TAG}_bench_clip_bf16_noshare.json 2> $OUT/${TAG}_bench_clip_bf16_noshare.err
  tail -c 1200 $OUT/${TAG}_bench_clip_bf16_noshare.json | head -c 700; echo
  timeout 300 python bench.py --workload clip --steps 3 --warmup 3 --batch 64 --no-cpu-baseline > $OUT/${TAG}_bench_clip_bf16_b64.json 2> $OUT/${TAG}_bench_clip_bf16_b64.err
  head -c 300 $OUT/${TAG}_bench_clip_bf16_b64.json; echo
fi
if has batch64; then
  timeout 300 python bench.py --workload clip --steps 3 --warmup 3 --batch 64 --no-cpu-baseline > $OUT/${TAG}_bench_clip_bf16_b64.json 2> $OUT/${TAG}_bench_clip_bf16_b64.err
  tail -c 600 $OUT/${TAG}_bench_clip_bf16_b64.json
  timeout 300 python bench.py --workload clip --steps 3 --warmup 3 --precision fp32 --no-cpu-baseline > $OUT/${TAG}_bench_clip_fp32.json 2> $OUT/${TAG}_bench_clip_fp32.err
  tail -c 600 $OUT/${TAG}_bench_clip_fp32.json
fi
if has ref; then
  timeout 300 python bench.py --workload clip --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_bench_clip_ref.json 2> $OUT/${TAG}_bench_clip_ref.err
  cat $OUT/${TAG}_bench_clip_ref.json
fi
if has list; then
  timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
    --log-file $OUT/${TAG}_launches.csv python bench.py --workload clip --frames 64 --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_list.log 2>&1
  echo "ncu list rc=$?"
fi
