#!/bin/bash
# quick A/B: GPU parity tests, then 4K (8 streams) and batch1080p numbers with per-kernel times
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
B="python bench.py --no-cpu-baseline --no-sensitivity --no-extras"
run() { name=$1; shift; "$@" > gpurun_out/qb_$name.json 2> gpurun_out/qb_$name.err; python - gpurun_out/qb_$name.json $name <<'PY'
import json,sys
for l in open(sys.argv[1]):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print(sys.argv[2], d.get('value'), d.get('ms_per_step'), d['roofline'].get('per_kernel_ms'), d['roofline'].get('frac'), d['config'].get('single_stream_ms_per_step'), d['config'].get('shared_device'))
PY
}
run 4k timeout 300 $B --steps 240 --warmup 24
run batch timeout 300 $B --workload batch1080p --steps 20 --warmup 3
