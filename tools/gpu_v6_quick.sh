# quick GPU loop for the version-6 kernel: its parity tests, then the bench (forced v6) with and without phase counters
TAG=${1:-q6}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dense.py -m gpu -x -q -k "v6" --timeout 150 -o timeout_method=thread 2>&1 | tail -40 > gpurun_out/${TAG}_pytest_v6.log
tail -5 gpurun_out/${TAG}_pytest_v6.log
if ! grep -q " passed" gpurun_out/${TAG}_pytest_v6.log || grep -q "failed\|Timeout\|error" gpurun_out/${TAG}_pytest_v6.log; then echo "V6 TESTS NOT GREEN"; [ "$2" == "force" ] || exit 1; fi
export BRR_GPU_DENSE_KERNEL=v6
BRR_GPU_PROFILE=1 timeout 400 python bench.py --no-cpu-baseline --steps 5 --warmup 4 > gpurun_out/${TAG}_bench_prof.json 2> gpurun_out/${TAG}_bench_prof.err
grep -h "phase cycles" gpurun_out/${TAG}_bench_prof.err; tail -2 gpurun_out/${TAG}_bench_prof.err | cut -c1-300
unset BRR_GPU_DENSE_KERNEL
timeout 400 python bench.py --no-cpu-baseline --steps 10 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python -c "
import json;d=json.load(open('gpurun_out/${TAG}_bench.json'));print('auto: %.4f s/sweep, frac %.3f, e2e %.4f' % (d['s_per_sweep'], d['roofline']['frac'], d['e2e']['s_per_sweep'])); print(d['config']['chain_start'], d['config']['dense_kernel_timed_sweeps'], d['config']['pipeline_restarts_per_sweep'], d['config']['nonskipped_fraction'])"
