set -x
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r1b_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r1b_smoke.log 2>&1
timeout 600 python bench.py > gpurun_out/r1b_bench.json 2> gpurun_out/r1b_bench.err
timeout 300 python tools/phase_profile.py 20000 3 > gpurun_out/r1b_phase.txt 2>&1
BRR_BENCH_N=62500 timeout 300 python tools/phase_profile.py 20000 3 > gpurun_out/r1b_phase_n62500.txt 2>&1
BRR_BENCH_M=8000 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1b_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r1b_ncu_bench.log 2>&1
BRR_BENCH_M=8000 timeout 900 ncu --set full --clock-control none --import-source on -k sweep_dense_kernel -s 3 -c 1 -o gpurun_out/r1b_sweep_v2 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r1b_ncu_full.log 2>&1
ls -la gpurun_out
cat gpurun_out/r1b_pytest.log gpurun_out/r1b_smoke.log gpurun_out/r1b_bench.json
