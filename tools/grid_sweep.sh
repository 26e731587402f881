#!/bin/bash
# A/B: K1 / K1b / K2 grid sizes on the 4K workload (8 streams), via the tuning environment variables
B="python bench.py --no-cpu-baseline --no-sensitivity --no-extras --steps 4800 --warmup 200"
run() { name=$1; shift; env "$@" timeout 300 $B $XTRA > gpurun_out/gs_$name.json 2> gpurun_out/gs_$name.err; python - gpurun_out/gs_$name.json $name <<'PY'
import json,sys
for l in open(sys.argv[1]):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print(sys.argv[2], d.get('value'), d.get('ms_per_step'), d['roofline'].get('per_kernel_ms'), d['config'].get('single_stream_ms_per_step'))
PY
}
XTRA="" run base A=1
XTRA="" run k1_64_b127_m127 JPEGB200_K1_GRID=64 JPEGB200_K1B_GRID=127 JPEGB200_K2_GRID=127
XTRA="" run k1_51_b127_m127 JPEGB200_K1_GRID=51 JPEGB200_K1B_GRID=127 JPEGB200_K2_GRID=127
XTRA="" run k1_74_b127_m127 JPEGB200_K1_GRID=74 JPEGB200_K1B_GRID=127 JPEGB200_K2_GRID=127
XTRA="" run k1_64_b148_m127 JPEGB200_K1_GRID=64 JPEGB200_K1B_GRID=148 JPEGB200_K2_GRID=127
XTRA="" run k1_64_b100_m127 JPEGB200_K1_GRID=64 JPEGB200_K1B_GRID=100 JPEGB200_K2_GRID=127
XTRA="" run k1_64_b127_m148 JPEGB200_K1_GRID=64 JPEGB200_K1B_GRID=127 JPEGB200_K2_GRID=148
XTRA="" run k1_64_b127_m100 JPEGB200_K1_GRID=64 JPEGB200_K1B_GRID=127 JPEGB200_K2_GRID=100
XTRA="--streams 12" run s12_k1_64_b127_m127 JPEGB200_K1_GRID=64 JPEGB200_K1B_GRID=127 JPEGB200_K2_GRID=127
XTRA="--streams 16" run s16_k1_64_b127_m127 JPEGB200_K1_GRID=64 JPEGB200_K1B_GRID=127 JPEGB200_K2_GRID=127
XTRA="--streams 16" run s16_k1_43_b85_m85 JPEGB200_K1_GRID=43 JPEGB200_K1B_GRID=85 JPEGB200_K2_GRID=85
XTRA="--streams 4" run s4_k1_64_b127_m127 JPEGB200_K1_GRID=64 JPEGB200_K1B_GRID=127 JPEGB200_K2_GRID=127
XTRA="--streams 4" run s4_base A=1
XTRA="--streams 16" run s16_base A=1
