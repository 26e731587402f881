#!/bin/bash
# ncu captures of round 2 (run under gpurun; every command is first run plain and must exit 0).
# usage: bash tools/ncu_round2.sh   -> gpurun_out/r02_*.ncu-rep, gpurun_out/r02_launches_*.csv
set -x
COMMON="--no-cpu-baseline --no-sensitivity --no-extras --streams 1 --no-graph --steps 4 --warmup 3"
K="regex:k_fused|k_strip|k_merge"
CMD4K="python bench.py $COMMON"
CMDB="python bench.py $COMMON --workload batch1080p --batch 64"
$CMD4K > gpurun_out/r02_plain_4k.log 2>&1 && ncu --set full --clock-control none --import-source on -k "$K" -s 21 -c 3 -f -o gpurun_out/r02_4k_tc $CMD4K > gpurun_out/r02_ncu_4k.log 2>&1
$CMD4K > gpurun_out/r02_plain_4k.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_launches_uhd4k.csv $CMD4K > /dev/null 2>&1
$CMDB > gpurun_out/r02_plain_b.log 2>&1 && ncu --set full --clock-control none --import-source on -k "$K" -s 21 -c 3 -f -o gpurun_out/r02_batch64_tc $CMDB > gpurun_out/r02_ncu_b.log 2>&1
JPEGB200_DCT=butterfly $CMDB > gpurun_out/r02_plain_bf.log 2>&1 && JPEGB200_DCT=butterfly ncu --set full --clock-control none --import-source on -k "regex:k_fused" -s 7 -c 1 -f -o gpurun_out/r02_batch64_butterfly $CMDB > gpurun_out/r02_ncu_bf.log 2>&1
CMD8K="python tools/encode_once.py 7680 4320 6"
$CMD8K > gpurun_out/r02_plain_8k.log 2>&1 && ncu --set full --clock-control none --import-source on -k "$K" -s 12 -c 3 -f -o gpurun_out/r02_8k_tc $CMD8K > gpurun_out/r02_ncu_8k.log 2>&1
ls -la gpurun_out/r02_*
