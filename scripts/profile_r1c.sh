mkdir -p gpurun_out
python bench.py > gpurun_out/r1c_bench_n1.json 2> gpurun_out/r1c_bench_n1.err; tail -c 3000 gpurun_out/r1c_bench_n1.json; tail -3 gpurun_out/r1c_bench_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1c_launches_raw.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --prove-log-gates 20 > gpurun_out/r1c_ncu_bench.log 2>&1
python scripts/launch_list.py gpurun_out/r1c_launches_raw.csv "ncu --metrics gpu__time_duration.sum --clock-control none: python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --prove-log-gates 20 (2^24-gate encode+commit steps, then one 2^20-gate prove+verify leg), 1 B200" > gpurun_out/r1c_launches_n1.csv 2>&1; head -40 gpurun_out/r1c_launches_n1.csv
ncu --set full --clock-control none --import-source on -k regex:hash_columns_quad -c 1 -f -o gpurun_out/r1c_quad python scripts/hash_probe.py > gpurun_out/r1c_ncu_quad.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:trace_level -s 20 -c 2 -f -o gpurun_out/r1c_trace python scripts/prove_probe.py 20 > gpurun_out/r1c_ncu_trace.log 2>&1
python scripts/ncu_summary.py gpurun_out/r1c_quad.ncu-rep gpurun_out/r1c_trace.ncu-rep > gpurun_out/r1c_ncu_full.jsonl 2>&1; cat gpurun_out/r1c_ncu_full.jsonl | cut -c1-1500
