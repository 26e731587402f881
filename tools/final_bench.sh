#!/bin/bash
# usage (under gpurun --gpus N): bash tools/final_bench.sh N   -> gpurun_out/r02_final_n$N.json
N=$1
if [ "$N" = "1" ]; then
  timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_final_n1.json 2> gpurun_out/r02_final_n1.err
  timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_final_ref.json 2> gpurun_out/r02_final_ref.err
  timeout 900 python bench.py --workload batch1080p --steps 20 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r02_final_batch.json 2> gpurun_out/r02_final_batch.err
else
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_final_n$N.json 2> gpurun_out/r02_final_n$N.err
fi
python - $N <<'PY'
import json,sys,glob
n=sys.argv[1]
for f in sorted(glob.glob(f'gpurun_out/r02_final_n{n}.json')+(glob.glob('gpurun_out/r02_final_ref.json')+glob.glob('gpurun_out/r02_final_batch.json') if n=='1' else [])):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, d.get('impl','ours'), d['value'], d['ms_per_step'], d.get('repeats'))
    if 'roofline' in d and d['roofline']: print('  roofline', d['roofline']['frac'], d['roofline'].get('step_frac'), d['roofline']['per_kernel_ms'])
    print('  e2e', d['e2e'])
    ex=(d.get('config') or {}).get('extra')
    if ex: print('  extra', {k:(v['value'], v.get('ms_per_image', v.get('ms_per_batch')), v.get('gather_to_rank0_ms'), v.get('verified')) for k,v in ex.items()})
    if d.get('cpu_baseline'): print('  cpu', {k:(v if not isinstance(v,dict) else v.get('value')) for k,v in d['cpu_baseline'].items() if k!='sample'})
PY
