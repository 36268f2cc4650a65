#!/bin/bash
# One GPU-box visit for the numbers and ncu evidence of a round:  bash tools/gpu_profile.sh TAG [stages...]
# Stages (default all): smoke bench cfg4 clip list full full4
TAG=${1:-run}; shift
STAGES=${*:-smoke bench cfg4 clip list full full4}
OUT=gpurun_out
mkdir -p $OUT
has() { [[ " $STAGES " == *" $1 "* ]]; }
line() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d.get("roofline") or {}
    c = r.get("classes_ms_per_batch") or {}
    print(f"{sys.argv[1]}: {d['value']:.1f} fps e2e {d['e2e']['value']:.1f} {d['ms_per_step']:.3f} ms trunk {r.get('achieved', 0):.0f} TF frac {r.get('frac', 0):.3f} clk {d['clocks']['sm_mhz']} "
          + " ".join(f"{k}={v:.3f}" for k, v in c.items()), d.get("batch1", ""), (r.get("conditioning_ms_per_batch") or ""))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
    try: print(open(sys.argv[1].replace(".json", ".err")).read()[-1500:])
    except Exception: pass
PY
}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/${TAG}_env.txt 2>&1; nproc >> $OUT/${TAG}_env.txt
if has smoke; then timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; tail -4 $OUT/${TAG}_smoke.log; fi
if has bench; then
  timeout 600 python bench.py --steps 100 --warmup 5 > $OUT/${TAG}_bench_fp32.json 2> $OUT/${TAG}_bench_fp32.err; line $OUT/${TAG}_bench_fp32.json
  timeout 600 python bench.py --steps 100 --warmup 5 --precision bf16 --no-cpu-baseline > $OUT/${TAG}_bench_bf16.json 2> $OUT/${TAG}_bench_bf16.err; line $OUT/${TAG}_bench_bf16.json
  timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err; tail -c 700 $OUT/${TAG}_bench_ref.json
fi
if has cfg4; then
  timeout 600 python bench.py --steps 50 --warmup 5 --output-nc 3 --precision bf16 --batch 32 --no-cpu-baseline > $OUT/${TAG}_cfg4_bf16_b32.json 2> $OUT/${TAG}_cfg4_bf16_b32.err; line $OUT/${TAG}_cfg4_bf16_b32.json
  timeout 600 python bench.py --steps 30 --warmup 5 --batch 64 --no-cpu-baseline > $OUT/${TAG}_cfg3_fp32_b64.json 2> $OUT/${TAG}_cfg3_fp32_b64.err; line $OUT/${TAG}_cfg3_fp32_b64.json
fi
if has clip; then
  timeout 600 python bench.py --workload clip --steps 5 --warmup 3 > $OUT/${TAG}_clip_bf16.json 2> $OUT/${TAG}_clip_bf16.err; line $OUT/${TAG}_clip_bf16.json
  timeout 600 python bench.py --workload clip --steps 3 --warmup 3 --precision fp32 --no-cpu-baseline > $OUT/${TAG}_clip_fp32.json 2> $OUT/${TAG}_clip_fp32.err; line $OUT/${TAG}_clip_fp32.json
  timeout 600 python bench.py --workload clip --steps 3 --warmup 3 --flow-net 32,2,4,batch --no-cpu-baseline > $OUT/${TAG}_clip_bf16_netF.json 2> $OUT/${TAG}_clip_bf16_netF.err; line $OUT/${TAG}_clip_bf16_netF.json
fi
if has list; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_list.log 2>&1; echo "ncu list rc=$?"
fi
if has full; then
  timeout 1200 ncu --set full --clock-control none --graph-profiling node --launch-skip 340 --launch-count 67 \
    -o /tmp/${TAG}_full -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_full.log 2>&1; echo "ncu full rc=$?"
  python tools/ncu_summary.py /tmp/${TAG}_full.ncu-rep > $OUT/${TAG}_full_summary.txt 2>&1; tail -3 $OUT/${TAG}_full_summary.txt
fi
if has full4; then
  timeout 1200 ncu --set full --clock-control none --graph-profiling node --launch-skip 340 --launch-count 67 \
    -o /tmp/${TAG}_full4 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --output-nc 3 --precision bf16 --batch 32 > $OUT/${TAG}_full4.log 2>&1; echo "ncu full4 rc=$?"
  python tools/ncu_summary.py /tmp/${TAG}_full4.ncu-rep > $OUT/${TAG}_full4_summary.txt 2>&1; tail -3 $OUT/${TAG}_full4_summary.txt
fi
ls $OUT | grep ${TAG} | head -40
