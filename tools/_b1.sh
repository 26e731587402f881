set -x
B="python bench.py --no-cpu-baseline --no-sensitivity"
timeout 300 $B --steps 2000 --warmup 50 > gpurun_out/r2a_tc_4k.json 2> gpurun_out/r2a_tc_4k.err
JPEGB200_DCT=butterfly timeout 300 $B --steps 2000 --warmup 50 > gpurun_out/r2a_bf_4k.json 2> gpurun_out/r2a_bf_4k.err
timeout 300 $B --steps 2000 --warmup 50 --streams 1 > gpurun_out/r2a_tc_4k_s1.json 2>> gpurun_out/r2a_tc_4k.err
JPEGB200_DCT=butterfly timeout 300 $B --steps 2000 --warmup 50 --streams 1 > gpurun_out/r2a_bf_4k_s1.json 2>> gpurun_out/r2a_bf_4k.err
timeout 300 $B --workload batch1080p --steps 30 --warmup 3 > gpurun_out/r2a_tc_batch.json 2> gpurun_out/r2a_tc_batch.err
JPEGB200_DCT=butterfly timeout 300 $B --workload batch1080p --steps 30 --warmup 3 > gpurun_out/r2a_bf_batch.json 2> gpurun_out/r2a_bf_batch.err
for f in gpurun_out/r2a_*.json; do echo $f; python - $f <<'PY'
import json,sys
for l in open(sys.argv[1]):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print({k:d.get(k) for k in ('value','ms_per_step','roofline','e2e')}); print(d.get('config',{}).get('kernel_us'))
PY
done
tail -3 gpurun_out/r2a_*.err
