# the quiet chain with phase counters (role counters of CTA 0 + the per-CTA dot-warp counters), at the shipped depth and at depth 7
TAG=${1:-qp}
mkdir -p gpurun_out
for D in 4 7; do
  echo "== depth $D"
  BRR_GPU_V6_DEPTH=$D BRR_GPU_DENSE_KERNEL=v6 BRR_GPU_PROFILE=1 timeout 200 python tools/bench_quiet.py > gpurun_out/${TAG}_d${D}.json 2> gpurun_out/${TAG}_d${D}.err
  tail -4 gpurun_out/${TAG}_d${D}.err | cut -c1-1200; cat gpurun_out/${TAG}_d${D}.json
  cp gpurun_out/quiet_ctas6.npy gpurun_out/${TAG}_d${D}_ctas.npy
done
