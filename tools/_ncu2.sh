CMD="python bench.py --no-cpu-baseline --no-sensitivity --steps 4 --warmup 3 --streams 1 --no-graph"
export JPEGB200_DCT=butterfly
$CMD > gpurun_out/plain2.log 2>&1 && ncu --metrics gpu__time_duration.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum --clock-control none -k regex:"k_fused|k_merge|k_strip" -s 9 -c 6 --csv --log-file gpurun_out/r2d_launches_bf.csv $CMD > /dev/null 2>&1
export JPEGB200_DCT=tc
$CMD > gpurun_out/plain3.log 2>&1 && ncu --metrics gpu__time_duration.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum --clock-control none -k regex:"k_fused|k_merge|k_strip" -s 9 -c 6 --csv --log-file gpurun_out/r2d_launches_tc.csv $CMD > /dev/null 2>&1
python - <<'PY'
import csv
for f in ('gpurun_out/r2d_launches_bf.csv','gpurun_out/r2d_launches_tc.csv'):
    rows=[r for r in csv.reader(open(f)) if len(r)>10]
    h=rows[0]
    for r in rows[1:]:
        print(f[-6:-4], r[h.index('Kernel Name')][:40], r[h.index('Metric Name')], r[h.index('Metric Value')])
PY
