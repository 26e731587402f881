B="python bench.py --no-cpu-baseline --no-sensitivity --no-extras --steps 240 --warmup 24"
for s in 4 6 8; do
timeout 300 $B --streams $s > gpurun_out/r2n_s$s.json 2> gpurun_out/r2n_s$s.err; python - gpurun_out/r2n_s$s.json s$s <<'PY'
import json,sys
for l in open(sys.argv[1]):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print(sys.argv[2], d.get('value'), d.get('ms_per_step'), d['e2e']['value'], d['e2e'].get('host_threads'))
PY
done
timeout 600 python bench.py --no-cpu-baseline --no-sensitivity --no-giga --steps 20 --warmup 5 > gpurun_out/r2n_x.json 2>gpurun_out/r2n_x.err; python -c "
import json
d=json.loads(open('gpurun_out/r2n_x.json').read().strip().splitlines()[-1]); print(d['config']['extra']['batch1080p_4096'])"
