mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
run() {  # n, tag, env...
  n=$1; tag=$2; shift 2
  env LG_BENCH_SKIP_E2E=1 "$@" $TR --nproc-per-node $n --master-port $((29800 + RANDOM % 100)) bench.py --gpus $n --steps 8 --warmup 3 > gpurun_out/r2c_n${n}_$tag.json 2> gpurun_out/r2c_n${n}_$tag.err
  python - <<PY
import json
ok=False
for l in open('gpurun_out/r2c_n${n}_$tag.json'):
    if l.startswith('{'):
        d=json.loads(l); ok=True
        print('N=$n $tag:', round(d['ms_per_step'],2), 'ms', d['root'][:12], {k:round(v,2) for k,v in d['kernel_ms_rank0'].items()})
if not ok: print('N=$n $tag: no json'); print(open('gpurun_out/r2c_n${n}_$tag.err').read()[-1200:])
PY
}
run 8 p0_t0 LG_SHARD_PIPELINE=0 LG_SHARD_TWO_STREAM=0
run 8 p0_t1 LG_SHARD_PIPELINE=0 LG_SHARD_TWO_STREAM=1
run 8 p1_t0_s1 LG_SHARD_PIPELINE=1 LG_SHARD_TWO_STREAM=0
run 8 p1_t1_s1 LG_SHARD_PIPELINE=1 LG_SHARD_TWO_STREAM=1
run 8 p1_t1_s2 LG_SHARD_PIPELINE=1 LG_SHARD_TWO_STREAM=1 LG_MGPU_SUB=2
run 8 p1_t0_s2 LG_SHARD_PIPELINE=1 LG_SHARD_TWO_STREAM=0 LG_MGPU_SUB=2
run 8 p0_t1_s2 LG_SHARD_PIPELINE=0 LG_SHARD_TWO_STREAM=1 LG_MGPU_SUB=2
run 4 p0_t1 LG_SHARD_PIPELINE=0 LG_SHARD_TWO_STREAM=1
run 4 p1_t1_s2 LG_SHARD_PIPELINE=1 LG_SHARD_TWO_STREAM=1 LG_MGPU_SUB=2
run 4 p1_t0_s1 LG_SHARD_PIPELINE=1 LG_SHARD_TWO_STREAM=0
