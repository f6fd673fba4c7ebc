# A/B of two builds of the library on one B200: the tree's libbrr_gpu.so (A) and ab/<name>.so (B, through BRR_GPU_LIB):
# version-6 parity tests, the quiet chain with phase counters, the bench.   usage: bash tools/gpu_ab.sh <tag> <B.so>
TAG=${1:-ab}; BLIB=${2:-ab/ctrlB.so}
mkdir -p gpurun_out
for V in A B; do
  if [ $V == B ]; then export BRR_GPU_LIB=$PWD/$BLIB; fi
  timeout 600 python -m pytest tests/test_gpu_dense.py -m gpu -x -q -k "v6" --timeout 150 -o timeout_method=thread 2>&1 | tail -8 > gpurun_out/${TAG}_${V}_pytest.log
  echo "== $V: $(tail -1 gpurun_out/${TAG}_${V}_pytest.log)"
  BRR_GPU_DENSE_KERNEL=v6 BRR_GPU_PROFILE=1 timeout 200 python tools/bench_quiet.py > gpurun_out/${TAG}_${V}_quietprof.json 2> gpurun_out/${TAG}_${V}_quietprof.err; tail -1 gpurun_out/${TAG}_${V}_quietprof.err | cut -c1-700
  BRR_GPU_DENSE_KERNEL=v6 timeout 200 python tools/bench_quiet.py 2>/dev/null | cut -c1-200
  timeout 400 python bench.py --no-cpu-baseline --steps 10 --warmup 5 > gpurun_out/${TAG}_${V}_bench.json 2> gpurun_out/${TAG}_${V}_bench.err
  python -c "
import json;d=json.load(open('gpurun_out/${TAG}_${V}_bench.json'));print('bench: %.4f s/sweep, frac %.3f, e2e %.4f' % (d['s_per_sweep'], d['roofline']['frac'], d['e2e']['s_per_sweep']), d['config']['dense_kernel_timed_sweeps'])"
done
