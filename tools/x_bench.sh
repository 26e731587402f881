#!/bin/bash
# timing experiment: K1 with half / none of the coefficient stores (results wrong, time only)
B="python bench.py --no-cpu-baseline --no-sensitivity --no-extras --no-shared-device --workload batch1080p --steps 20 --warmup 3"
for v in "" _x_NOFLAG; do
  JPEGB200_LIB=$PWD/jpeg_image_compression_b200/libjpegb200$v.so timeout 300 $B > gpurun_out/xb$v.json 2> gpurun_out/xb$v.err
  python - gpurun_out/xb$v.json "base$v" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print(sys.argv[2], d.get('value'), d.get('ms_per_step'), d['roofline'].get('per_kernel_ms'))
PY
done
JPEGB200_LIB=$PWD/jpeg_image_compression_b200/libjpegb200_x_NOFLAG.so timeout 100 python bench.py --no-cpu-baseline --no-sensitivity --no-extras --no-shared-device --steps 2400 --warmup 100 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('4k noflag', d['value'], d['roofline']['per_kernel_ms'])"
