#!/bin/bash
# One multi-GPU box visit: bash tools/gpu_multi.sh TAG N [stages...]   (stages: test check bench benchnccl clip)
TAG=$1; N=$2; shift; shift
STAGES=${*:-test check bench}
OUT=gpurun_out
mkdir -p $OUT
has() { [[ " $STAGES " == *" $1 "* ]]; }
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port"
nvidia-smi topo -m > $OUT/${TAG}_topo.txt 2>&1
if has test; then
  timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "two_devices" > $OUT/${TAG}_pytest2dev.log 2>&1
  tail -3 $OUT/${TAG}_pytest2dev.log
fi
if has check; then
  for G in nccl peer; do
    timeout 600 $RUN 29533 tools/check_sharded.py $G > $OUT/${TAG}_check_$G.log 2>&1
    grep -E "OK|FAIL|PASS|Error|error|frames\]" $OUT/${TAG}_check_$G.log | tail -12
  done
fi
if has bench; then
  timeout 900 $RUN 29511 bench.py --gpus $N --steps 20 --warmup 3 > $OUT/${TAG}_bench_n$N.json 2> $OUT/${TAG}_bench_n$N.err
  tail -c 2500 $OUT/${TAG}_bench_n$N.json; tail -5 $OUT/${TAG}_bench_n$N.err
fi
if has benchnccl; then
  timeout 900 $RUN 29512 bench.py --gpus $N --steps 20 --warmup 3 --gather nccl > $OUT/${TAG}_bench_nccl_n$N.json 2> $OUT/${TAG}_bench_nccl_n$N.err
  tail -c 1200 $OUT/${TAG}_bench_nccl_n$N.json; tail -5 $OUT/${TAG}_bench_nccl_n$N.err
fi
if has clip; then
  timeout 900 $RUN 29513 bench.py --gpus $N --workload clip --steps 5 --warmup 3 > $OUT/${TAG}_clip_n$N.json 2> $OUT/${TAG}_clip_n$N.err
  tail -c 1200 $OUT/${TAG}_clip_n$N.json; tail -5 $OUT/${TAG}_clip_n$N.err
fi
