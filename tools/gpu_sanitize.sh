#!/bin/bash
# compute-sanitizer memcheck over the flow network (two configurations) and over smoke() (generator, conditioning, netF).
OUT=gpurun_out
TAG=${1:-san}
mkdir -p $OUT
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 python tools/flow_probe.py --batch 3 --once > $OUT/${TAG}_memcheck_netF.log 2>&1
echo "memcheck netF rc=$?"; tail -3 $OUT/${TAG}_memcheck_netF.log
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 python tools/flow_probe.py --batch 2 --cfg 16,4,3,instance --once > $OUT/${TAG}_memcheck_netF2.log 2>&1
echo "memcheck netF (16,4,3,instance) rc=$?"; tail -3 $OUT/${TAG}_memcheck_netF2.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_memcheck_smoke.log 2>&1
echo "memcheck smoke rc=$?"; tail -6 $OUT/${TAG}_memcheck_smoke.log
