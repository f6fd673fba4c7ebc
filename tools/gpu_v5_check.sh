# GPU check of the version-5 dense kernel: parity tests (v5 first), then the bench with each kernel forced and in auto mode.
# usage: bash tools/gpu_v5_check.sh <tag> [full]
TAG=${1:-v5}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv,noheader
timeout 900 python -m pytest tests/test_gpu_dense.py -m gpu -x -q -k "v5" --timeout 150 -o timeout_method=thread 2>&1 | tail -30 > gpurun_out/${TAG}_pytest_v5.log
tail -6 gpurun_out/${TAG}_pytest_v5.log
if ! grep -q " passed" gpurun_out/${TAG}_pytest_v5.log || grep -q "failed\|Timeout\|error" gpurun_out/${TAG}_pytest_v5.log; then echo "V5 TESTS NOT GREEN"; [ "$2" == "force" ] || exit 1; fi
for K in v5 auto; do
  if [ "$K" == "auto" ]; then unset BRR_GPU_DENSE_KERNEL; else export BRR_GPU_DENSE_KERNEL=$K; fi
  BRR_GPU_PROFILE=1 timeout 400 python bench.py --no-cpu-baseline --steps 5 --warmup 4 > gpurun_out/${TAG}_bench_prof_$K.json 2> gpurun_out/${TAG}_bench_prof_$K.err
  grep -h "phase cycles" gpurun_out/${TAG}_bench_prof_$K.err; tail -2 gpurun_out/${TAG}_bench_prof_$K.err
  timeout 400 python bench.py --no-cpu-baseline --steps 8 --warmup 4 > gpurun_out/${TAG}_bench_$K.json 2> gpurun_out/${TAG}_bench_$K.err
  python -c "
import json;d=json.load(open('gpurun_out/${TAG}_bench_$K.json'));print('$K: %.4f s/sweep, frac %.3f, e2e %.4f' % (d['s_per_sweep'], d['roofline']['frac'], d['e2e']['s_per_sweep'])); print(d['config']['chain_start'], d['config']['dense_kernel_timed_sweeps'], d['config']['pipeline_restarts_per_sweep'], d['config']['nonskipped_fraction'])"
done
unset BRR_GPU_DENSE_KERNEL
if [ "$2" == "full" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q --timeout 150 -o timeout_method=thread 2>&1 | tail -15 > gpurun_out/${TAG}_pytest_all.log; tail -5 gpurun_out/${TAG}_pytest_all.log
fi
