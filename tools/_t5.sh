timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
( time timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2h_driver.json 2> gpurun_out/r2h_driver.err ) 2>&1 | grep real
tail -3 gpurun_out/r2h_driver.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2h_driver.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','repeats','timed_steps','gpu_launches')})
print('roofline',d['roofline'])
print('e2e',d['e2e'])
print('extra',json.dumps(d['config']['extra'],indent=1))
print('cpu',d['cpu_baseline'])
print('sens',d['config']['sensitivity'], d['config']['single_stream_ms_per_step'])
PY
( time timeout 300 python bench.py --impl reference --steps 4 --warmup 1 ) 2>&1 | tail -5
