mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$TR --nproc-per-node 8 --master-port 29701 scripts/mgpu_check.py 2>&1 | grep -E "MGPU|world=8" | tail -8
run() {  # n, tag, env...
  n=$1; tag=$2; shift 2
  env "$@" $TR --nproc-per-node $n --master-port $((29800 + RANDOM % 100)) bench.py --gpus $n --steps 8 --warmup 3 > gpurun_out/r2b_n${n}_$tag.json 2> gpurun_out/r2b_n${n}_$tag.err
  python - <<PY
import json
ok=False
for l in open('gpurun_out/r2b_n${n}_$tag.json'):
    if l.startswith('{'):
        d=json.loads(l); ok=True
        print('N=$n $tag:', round(d['ms_per_step'],2), 'ms  e2e', (round(d['e2e']['ms_per_step'],2) if d.get('e2e') else None), d['root'][:12], {k:round(v,2) for k,v in d['kernel_ms_rank0'].items()})
if not ok: print('N=$n $tag: no json'); print(open('gpurun_out/r2b_n${n}_$tag.err').read()[-1200:])
PY
}
run 8 deferred X=1
run 8 deferred_sub2 LG_MGPU_SUB=2 LG_BENCH_SKIP_E2E=1
run 8 deferred_e2e2 LG_MGPU_E2E_PIPELINE=2
run 8 nopipe LG_SHARD_PIPELINE=0 LG_MGPU_E2E_PIPELINE=0
run 4 deferred X=1
run 4 deferred_sub2 LG_MGPU_SUB=2 LG_BENCH_SKIP_E2E=1
run 4 nopipe LG_SHARD_PIPELINE=0 LG_MGPU_E2E_PIPELINE=0
run 2 deferred LG_SHARD_PIPELINE=2 LG_MGPU_E2E_PIPELINE=2
