#!/bin/bash
# A/B timing of the tcgen05 conv kernel under diagnostic switches (kernel durations from ncu).
#   bash tools/conv_probe.sh TAG "Cin Cout S stride pad transposed B impl" "ENV1=.. ENV2=.." "ENV..." ...
TAG=$1; LAYER=$2; shift; shift
mkdir -p gpurun_out
: > gpurun_out/${TAG}_probe.txt
for ENVS in "$@"; do
  echo "== $LAYER :: $ENVS" >> gpurun_out/${TAG}_probe.txt
  env $ENVS timeout 300 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,l1tex__m_xbar2l1tex_read_bytes.sum.per_second \
    --clock-control none -k regex:conv_umma --csv python tools/conv_probe.py $LAYER 2>&1 | grep -E '^"[0-9]' | \
    awk -F'","' '{print $5, $(NF-2), $(NF)}' | sed 's/"//g' >> gpurun_out/${TAG}_probe.txt
done
cat gpurun_out/${TAG}_probe.txt
