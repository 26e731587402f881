CMD="python bench.py --no-cpu-baseline --no-sensitivity --workload batch1080p --batch 64 --steps 2 --warmup 3"
export JPEGB200_DCT=tc
$CMD > gpurun_out/plain4.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_fused|k_strip|k_merge" -s 9 -c 3 -f -o gpurun_out/r2e_batch64 $CMD > gpurun_out/ncu4.log 2>&1
tail -3 gpurun_out/ncu4.log
