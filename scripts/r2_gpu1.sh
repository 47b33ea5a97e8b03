# one-GPU measurement pass of round 2 (run under gpurun from the repo root)
mkdir -p gpurun_out
scripts/_bin/pipe_probe3 > gpurun_out/r2_pipe_probe3.jsonl 2>&1; cat gpurun_out/r2_pipe_probe3.jsonl
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; tail -c 2500 gpurun_out/r2_bench_n1.json; tail -3 gpurun_out/r2_bench_n1.err
python bench.py --workload micro --rows 16384 --n 65536 --rho-inv 4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_micro_16384x65536.json 2> gpurun_out/r2_micro.err; cut -c1-600 gpurun_out/r2_micro_16384x65536.json; tail -2 gpurun_out/r2_micro.err
python bench.py --workload micro --rows 256 --n 1024 --rho-inv 4 --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2_micro_256x1024.json 2>> gpurun_out/r2_micro.err; cut -c1-400 gpurun_out/r2_micro_256x1024.json
ncu --set full --clock-control none --import-source on -k regex:"ntt_persist|ntt_global|hash_columns" -c 4 -f -o gpurun_out/r2_full python scripts/gpu_probe.py 16388x8192x8 > gpurun_out/r2_ncu_full.log 2>&1; tail -2 gpurun_out/r2_ncu_full.log
python scripts/ncu_summary.py gpurun_out/r2_full.ncu-rep > gpurun_out/r2_ncu_full_2p24.jsonl 2>&1; cut -c1-300 gpurun_out/r2_ncu_full_2p24.jsonl
