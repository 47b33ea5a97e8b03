mkdir -p gpurun_out
python scripts/mgpu_single_process_check.py 2 2>&1 | tail -6
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for cfg in "LG_SHARD_PIPELINE=0" "LG_SHARD_GROUPS=3" "LG_SHARD_GROUPS=2" "LG_SHARD_GROUPS=3 LG_MGPU_SUB=2"; do
  tag=$(echo $cfg | tr ' =' '__')
  env $cfg $TR --master-port $((29600 + RANDOM % 200)) bench.py --gpus 2 --steps 6 --warmup 3 > gpurun_out/r2_n2_$tag.json 2> gpurun_out/r2_n2_$tag.err
  echo "== $cfg rc=$?"
  python -c "
import json,sys
try:
    d=json.load(open('gpurun_out/r2_n2_$tag.json')); print(d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['root'][:16], d['kernel_ms_rank0'])
except Exception as e:
    print('no json', e); print(open('gpurun_out/r2_n2_$tag.err').read()[-1500:])
"
done
