timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
B="python bench.py --no-cpu-baseline --no-sensitivity"
for mode in tc butterfly; do
  export JPEGB200_DCT=$mode
  timeout 300 $B --steps 2000 --warmup 50 > gpurun_out/r2c_${mode}_4k.json 2> gpurun_out/r2c_${mode}_4k.err
  timeout 300 $B --workload batch1080p --steps 30 --warmup 3 > gpurun_out/r2c_${mode}_batch.json 2> gpurun_out/r2c_${mode}_batch.err
done
for f in gpurun_out/r2c_*.json; do echo $f; python - $f <<'PY'
import json,sys
for l in open(sys.argv[1]):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print(d.get('value'), d.get('ms_per_step'), d['roofline'].get('per_kernel_ms'), d['roofline'].get('frac'))
PY
done
