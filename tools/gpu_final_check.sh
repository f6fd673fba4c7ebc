# Round check of the shipped build on one B200: all GPU tests, smoke, bench (+ profile counters), ncu launch list, one ncu --set full capture of the
# version-6 sweep kernel, DRAM traffic of one full-size launch.  usage: bash tools/gpu_final_check.sh <tag> [full]   (full: with the ncu --set full capture)
TAG=${1:-fin}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv,noheader
timeout 120 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; tail -2 gpurun_out/${TAG}_smoke.log
timeout 900 python -m pytest tests -m gpu -q --timeout 150 -o timeout_method=thread 2>&1 | tail -15 > gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 600 gpurun_out/${TAG}_bench.json; tail -2 gpurun_out/${TAG}_bench.err
BRR_GPU_PROFILE=1 timeout 300 python bench.py --no-cpu-baseline --steps 10 --warmup 5 > gpurun_out/${TAG}_bench_prof.json 2> gpurun_out/${TAG}_bench_prof.err; grep -h "phase cycles" gpurun_out/${TAG}_bench_prof.err | cut -c1-900
BRR_GPU_DENSE_KERNEL=v6 BRR_GPU_PROFILE=1 timeout 200 python tools/bench_quiet.py > gpurun_out/${TAG}_quiet.json 2> gpurun_out/${TAG}_quiet.err; cat gpurun_out/${TAG}_quiet.json; tail -1 gpurun_out/${TAG}_quiet.err
BRR_BENCH_M=8000 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"sweep_dense|inmodel_mask|eps_add_reduce" -c 60 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 4 --warmup 8 --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1
[ "$2" == "full" ] && BRR_BENCH_M=64000 timeout 700 ncu --set full --clock-control none --import-source on -k regex:sweep_dense6 -s 3 -c 1 -o gpurun_out/${TAG}_sweep6 -f python bench.py --steps 2 --warmup 8 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full.log 2>&1
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:sweep_dense6 -s 3 -c 1 --csv --log-file gpurun_out/${TAG}_traffic.csv python bench.py --steps 2 --warmup 8 --no-cpu-baseline > gpurun_out/${TAG}_traffic.log 2>&1
tail -3 gpurun_out/${TAG}_traffic.csv | cut -c1-300
ls -la gpurun_out/${TAG}_*
