B="python bench.py --no-cpu-baseline --no-sensitivity --no-extras"
run() { name=$1; shift; "$@" > gpurun_out/r2i_$name.json 2> gpurun_out/r2i_$name.err; python - gpurun_out/r2i_$name.json $name <<'PY'
import json,sys
for l in open(sys.argv[1]):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print(sys.argv[2], d.get('value'), d.get('ms_per_step'), d['roofline'].get('per_kernel_ms'), d['roofline'].get('frac'), d['config'].get('single_stream_ms_per_step'))
PY
}
run carve_4k timeout 300 $B --steps 200 --warmup 20
run carve_batch timeout 300 $B --workload batch1080p --steps 20 --warmup 3
export JPEGB200_NO_CARVEOUT=1
run nocarve_4k timeout 300 $B --steps 200 --warmup 20
run nocarve_batch timeout 300 $B --workload batch1080p --steps 20 --warmup 3
