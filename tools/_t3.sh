B="python bench.py --no-cpu-baseline --no-sensitivity --steps 3000 --warmup 50"
run() { name=$1; shift; "$@" > gpurun_out/r2d_$name.json 2> gpurun_out/r2d_$name.err; python - gpurun_out/r2d_$name.json $name <<'PY'
import json,sys
for l in open(sys.argv[1]):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print(sys.argv[2], d.get('value'), d.get('ms_per_step'), d['roofline'].get('per_kernel_ms'))
PY
}
export JPEGB200_DCT=butterfly
run bf_s2 timeout 300 $B --streams 2
run bf_s3 timeout 300 $B --streams 3
run bf_s4 timeout 300 $B --streams 4
JPEGB200_K1_GRID=148 run bf_g148_s3 timeout 300 $B --streams 3
JPEGB200_K1B_GRID=296 run bf_k1b296_s3 timeout 300 $B --streams 3
export JPEGB200_DCT=tc
run tc_s2 timeout 300 $B --streams 2
run tc_s3 timeout 300 $B --streams 3
run tc_s4 timeout 300 $B --streams 4
