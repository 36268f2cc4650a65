OUT=gpurun_out
TAG=r03d
timeout 400 ncu --set full --clock-control none -k regex:"fconv_umma_kernel|fconv_sparse_kernel|kp_kernel" --launch-count 13 -o /tmp/${TAG}_umma -f python tools/flow_probe.py --once > $OUT/${TAG}_umma_full.log 2>&1
echo "ncu rc=$?"
python tools/ncu_summary.py /tmp/${TAG}_umma.ncu-rep > $OUT/${TAG}_umma_full_summary.txt 2>&1
cat $OUT/${TAG}_umma_full_summary.txt
