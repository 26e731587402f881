timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
B="python bench.py --no-cpu-baseline --no-sensitivity"
run() { name=$1; shift; "$@" > gpurun_out/r2p_$name.json 2> gpurun_out/r2p_$name.err; python - gpurun_out/r2p_$name.json $name <<'PY'
import json,sys
for l in open(sys.argv[1]):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print(sys.argv[2], d.get('value'), d.get('ms_per_step'), d['roofline'].get('per_kernel_ms'), d['roofline'].get('frac'))
PY
}
export JPEGB200_DCT=tc
run tc_s2 timeout 300 $B --steps 3000 --warmup 50 --streams 2
run tc_s4 timeout 300 $B --steps 3000 --warmup 50 --streams 4
run tc_batch timeout 300 $B --workload batch1080p --steps 30 --warmup 3
export JPEGB200_DCT=butterfly
run bf_batch timeout 300 $B --workload batch1080p --steps 30 --warmup 3
