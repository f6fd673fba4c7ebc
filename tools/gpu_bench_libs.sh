# the bench's chain with several builds of the library (ab/<name>.so through BRR_GPU_LIB), alternating, <reps> times
TAG=${1:-bl}; REPS=${2:-2}; shift; shift
for r in $(seq 1 $REPS); do for L in "$@"; do
  BRR_GPU_LIB=$PWD/ab/$L.so timeout 400 python bench.py --no-cpu-baseline --steps 10 --warmup 5 > gpurun_out/${TAG}_${L}_$r.json 2> gpurun_out/${TAG}_${L}_$r.err
  python -c "
import json;d=json.load(open('gpurun_out/${TAG}_${L}_$r.json'));print('$L #$r: %.4f s/sweep, frac %.3f, e2e %.4f' % (d['s_per_sweep'], d['roofline']['frac'], d['e2e']['s_per_sweep']), d['config']['dense_kernel_timed_sweeps'])"
done; done
