# the bench's chain at several pipeline depths of the version-6 kernel (BRR_GPU_V6_DEPTH)
TAG=${1:-dp}; shift
for d in "$@"; do
  BRR_GPU_V6_DEPTH=$d timeout 400 python bench.py --no-cpu-baseline --steps 10 --warmup 5 > gpurun_out/${TAG}_bench_d$d.json 2> gpurun_out/${TAG}_bench_d$d.err
  python -c "
import json;d=json.load(open('gpurun_out/${TAG}_bench_d$d.json'));print('depth $d: %.4f s/sweep, frac %.3f, e2e %.4f' % (d['s_per_sweep'], d['roofline']['frac'], d['e2e']['s_per_sweep']), d['config']['dense_kernel_timed_sweeps'], d['config']['pipeline_restarts_per_sweep'])"
done
