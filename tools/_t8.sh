B="python bench.py --no-cpu-baseline --no-sensitivity --no-extras"
run() { name=$1; shift; "$@" > gpurun_out/r2k_$name.json 2> gpurun_out/r2k_$name.err; python - gpurun_out/r2k_$name.json $name <<'PY'
import json,sys
for l in open(sys.argv[1]):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print(sys.argv[2], d.get('value'), d.get('ms_per_step'), d['roofline'].get('per_kernel_ms'), d['roofline'].get('frac'))
PY
}
export JPEGB200_LIB=$PWD/jpeg_image_compression_b200/libjpegb200_noflag.so
run noflag_batch timeout 300 $B --workload batch1080p --steps 20 --warmup 3
JPEGB200_DCT=butterfly run noflag_bf_batch timeout 300 $B --workload batch1080p --steps 20 --warmup 3
