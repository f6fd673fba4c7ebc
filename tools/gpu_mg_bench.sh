# bench.py on N GPUs of one box (run under `gpurun --gpus N`), the driver's launch line.   usage: bash tools/gpu_mg_bench.sh <tag> <N> [steps]
TAG=${1:-mg}; N=${2:-2}; STEPS=${3:-10}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps $STEPS --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench_${N}gpu.json 2> gpurun_out/${TAG}_bench_${N}gpu.err
tail -2 gpurun_out/${TAG}_bench_${N}gpu.err | cut -c1-300
python -c "
import json;d=json.loads(open('gpurun_out/${TAG}_bench_${N}gpu.json').read().strip().splitlines()[-1]);print('$N GPUs: %.4f s/sweep, e2e %.4f' % (d['s_per_sweep'], d['e2e']['s_per_sweep']), d['config']['dense_kernel_timed_sweeps'], 'ranks identical:', d['config']['ranks_bit_identical'])"
