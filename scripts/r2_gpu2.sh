# two-GPU correctness + bench pass (gpurun --gpus 2)
mkdir -p gpurun_out
nvidia-smi -L | head -3
python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -15
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555"
$TR bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; tail -c 1800 gpurun_out/r2_bench_n2.json; tail -3 gpurun_out/r2_bench_n2.err
LG_SHARD_PIPELINE=0 $TR bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_n2_nopipe.json 2> gpurun_out/r2_bench_n2_nopipe.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_n2_nopipe.json')); print('nopipe', d['ms_per_step'], d['e2e']['ms_per_step'], d['root'][:16])"; tail -2 gpurun_out/r2_bench_n2_nopipe.err
