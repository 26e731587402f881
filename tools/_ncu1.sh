set -x
CMD="python bench.py --no-cpu-baseline --no-sensitivity --steps 4 --warmup 3 --streams 1 --no-graph"
$CMD > gpurun_out/plain_tc.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_fused|k_merge" -s 8 -c 4 -f -o gpurun_out/r2a_tc $CMD > gpurun_out/ncu_tc.log 2>&1
export JPEGB200_DCT=butterfly
$CMD > gpurun_out/plain_bf.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_fused" -s 8 -c 2 -f -o gpurun_out/r2a_bf $CMD > gpurun_out/ncu_bf.log 2>&1
tail -5 gpurun_out/ncu_tc.log gpurun_out/ncu_bf.log
