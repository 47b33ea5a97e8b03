mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
run() {  # n, tag, env...
  n=$1; tag=$2; shift 2
  env "$@" $TR --nproc-per-node $n --master-port $((29800 + RANDOM % 100)) bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/r2d_n${n}_$tag.json 2> gpurun_out/r2d_n${n}_$tag.err
  python - <<PY
import json
ok=False
for l in open('gpurun_out/r2d_n${n}_$tag.json'):
    if l.startswith('{'):
        d=json.loads(l); ok=True
        print('N=$n $tag:', round(d['ms_per_step'],2), 'ms e2e', (round(d['e2e']['ms_per_step'],2) if d.get('e2e') else None), d['root'][:12], {k:round(v,2) for k,v in d['kernel_ms_rank0'].items()})
if not ok: print('N=$n $tag: no json'); print(open('gpurun_out/r2d_n${n}_$tag.err').read()[-1200:])
PY
}
run 8 default X=1
run 8 sub3 LG_MGPU_SUB=3 LG_BENCH_SKIP_E2E=1
run 8 sub4 LG_MGPU_SUB=4 LG_BENCH_SKIP_E2E=1
run 8 sub2_g3 LG_SHARD_GROUPS=3 LG_BENCH_SKIP_E2E=1
run 4 default X=1
$TR --nproc-per-node 8 --master-port 29702 scripts/mgpu_prove_check.py 14 18 2>&1 | grep -E "MGPU|gates" | tail -3
