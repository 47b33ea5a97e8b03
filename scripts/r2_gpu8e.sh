mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$TR --nproc-per-node 8 --master-port 29701 scripts/mgpu_check.py 2>&1 | grep -E "MGPU|world=8" | tail -3
for n in 8 2; do
  $TR --nproc-per-node $n --master-port $((29800 + n)) bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/r2e_n${n}.json 2> gpurun_out/r2e_n${n}.err
  python - <<PY
import json
ok=False
for l in open('gpurun_out/r2e_n${n}.json'):
    if l.startswith('{'):
        d=json.loads(l); ok=True
        print('N=$n default:', round(d['ms_per_step'],2), 'ms e2e', round(d['e2e']['ms_per_step'],2), d['root'][:12], d['config']['parallelism'][:60], {k:round(v,2) for k,v in d['kernel_ms_rank0'].items()})
if not ok: print('N=$n: no json'); print(open('gpurun_out/r2e_n${n}.err').read()[-1500:])
PY
done
