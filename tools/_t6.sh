N=$1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2h_n$N.json 2> gpurun_out/r2h_n$N.err
tail -5 gpurun_out/r2h_n$N.err
python - $N <<'PY'
import json,sys
d=json.loads(open(f'gpurun_out/r2h_n{sys.argv[1]}.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','repeats','n_gpus')})
print('e2e',d['e2e'])
print('extra',json.dumps(d['config']['extra'],indent=1))
PY
