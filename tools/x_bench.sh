#!/bin/bash
# Timing experiments on K1 (results are WRONG by construction, only the times mean something; the numbers are in
# profiles/r02_design_experiments.md section 6): the library built with one of
#   -DJB_EXPERIMENT_NOSTORE    none of the coefficient stores (the "half of them" variant of section 6 became the two-plane layout)
#   -DJB_EXPERIMENT_NOFLAG     no exact re-evaluation of flagged coefficients
# usage:  bash tools/x_bench.sh build     (here: nvcc cross-compiles the variants next to the library)
#         gpurun -- 'bash tools/x_bench.sh'   (on the B200: batch of 512 x 1080p with each variant, per-kernel times)
cd "$(dirname "$0")/.."
P=jpeg_image_compression_b200
VARIANTS="NOSTORE NOFLAG"
if [ "$1" = "build" ]; then
  make -C $P/csrc all > /dev/null || exit 1
  for v in $VARIANTS; do
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -DJB_EXPERIMENT_$v \
      -Iinclude -I$P/csrc -c $P/csrc/jpegb200.cu -o $P/csrc/build/x_$v.o 2> /dev/null &&
    nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $P/libjpegb200_x_$v.so $P/csrc/build/x_$v.o $P/csrc/build/tables.o $P/csrc/build/bmp_io.o $P/csrc/build/jfif_io.o -lcudart || exit 1
  done
  exit 0
fi
B="python bench.py --no-cpu-baseline --no-sensitivity --no-extras --no-shared-device --workload batch1080p --steps 20 --warmup 3"
for v in "" $VARIANTS; do
  lib=$PWD/$P/libjpegb200${v:+_x_$v}.so
  [ -f "$lib" ] || continue
  JPEGB200_LIB=$lib timeout 300 $B 2> /dev/null | python -c "
import json, sys
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('${v:-shipped}', d['value'], d['ms_per_step'], d['roofline']['per_kernel_ms'])"
done
