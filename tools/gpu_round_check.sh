# GPU check of a build: parity tests, smoke, phase breakdown, bench.  usage: bash tools/gpu_round_check.sh <tag> [quick]
TAG=${1:-chk}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv,noheader
timeout 120 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; tail -2 gpurun_out/${TAG}_smoke.log
timeout 600 python -m pytest tests -m gpu -x -q --timeout 100 -o timeout_method=thread 2>&1 | tail -25 > gpurun_out/${TAG}_pytest.log
tail -5 gpurun_out/${TAG}_pytest.log
if ! grep -q " passed" gpurun_out/${TAG}_pytest.log || grep -q "failed\|Timeout" gpurun_out/${TAG}_pytest.log; then echo "TESTS NOT GREEN: stopping"; exit 1; fi
timeout 120 python tools/phase_profile.py 20000 8 > gpurun_out/${TAG}_phase.txt 2>&1; tail -9 gpurun_out/${TAG}_phase.txt
BRR_BENCH_N=62500 timeout 120 python tools/phase_profile.py 20000 8 > gpurun_out/${TAG}_phase_n62500.txt 2>&1; tail -3 gpurun_out/${TAG}_phase_n62500.txt
BRR_GPU_PROFILE=1 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/${TAG}_bench_prof.json 2> gpurun_out/${TAG}_bench_prof.err; grep -h "phase cycles" gpurun_out/${TAG}_bench_prof.err; python -c "import json;d=json.load(open('gpurun_out/${TAG}_bench_prof.json'));print('profiled bench: %.4f s/sweep' % d['s_per_sweep'])"
if [ "$2" != "quick" ]; then  # full bench also in ncu mode
  timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; cat gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
fi
if [ "$2" == "ncu" ]; then
  BRR_BENCH_M=8000 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"sweep_dense_kernel|eps_add_reduce_kernel" -c 30 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1
  BRR_BENCH_M=64000 timeout 600 ncu --set full --clock-control none --import-source on -k sweep_dense_kernel -s 4 -c 1 -o gpurun_out/${TAG}_sweep -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full.log 2>&1
  ls -la gpurun_out/${TAG}_*
fi
