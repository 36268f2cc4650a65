#!/bin/bash
# One GPU-box visit: parity tests, bench (ours + reference arm), ncu launch list, ncu --set full capture.
#   gpurun --timeout 1500 -- 'bash tools/gpu_round.sh TAG [stages...]'
# Stages (default all): test bench ref list full
# Everything lands in gpurun_out/TAG_*.
TAG=${1:-run}
shift
STAGES=${*:-test bench ref list full}
OUT=gpurun_out
mkdir -p $OUT
has() { [[ " $STAGES " == *" $1 "* ]]; }

nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/${TAG}_env.txt 2>&1
nproc >> $OUT/${TAG}_env.txt

if has test; then
  timeout 1200 python -m pytest tests -x -q -m gpu > $OUT/${TAG}_pytest.log 2>&1
  echo "pytest rc=$?" >> $OUT/${TAG}_pytest.log
  tail -5 $OUT/${TAG}_pytest.log
fi
if has bench; then
  timeout 600 python bench.py --steps 50 --warmup 5 > $OUT/${TAG}_bench_fp32.json 2> $OUT/${TAG}_bench_fp32.err
  tail -c 3000 $OUT/${TAG}_bench_fp32.json
  timeout 600 python bench.py --steps 50 --warmup 5 --precision bf16 --no-cpu-baseline > $OUT/${TAG}_bench_bf16.json 2> $OUT/${TAG}_bench_bf16.err
  tail -c 1500 $OUT/${TAG}_bench_bf16.json
fi
if has ref; then
  timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err
  cat $OUT/${TAG}_bench_ref.json
fi
if has list; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_list.log 2>&1
  echo "ncu list rc=$?"
fi
if has full; then
  # one whole forward (66 launches) with the full metric set; the .ncu-rep is reduced to CSV on the box because
  # gpurun only brings back 64 MiB
  timeout 1200 ncu --set full --clock-control none --launch-skip 300 --launch-count 66 \
    -o /tmp/${TAG}_full -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_full.log 2>&1
  echo "ncu full rc=$?"
  ncu -i /tmp/${TAG}_full.ncu-rep --page raw --csv > $OUT/${TAG}_full_raw.csv 2>/dev/null
  python tools/ncu_summary.py /tmp/${TAG}_full.ncu-rep > $OUT/${TAG}_full_summary.txt 2>&1
fi
if has src; then
  # source-level capture of the dominant kernel only (3 launches), small enough to travel
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:${NCU_KERNEL:-conv_umma} --launch-skip ${NCU_SKIP:-40} --launch-count 3 \
    -o $OUT/${TAG}_src -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_src.log 2>&1
  echo "ncu src rc=$?"
fi
ls -la $OUT
