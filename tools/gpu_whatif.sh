# timing experiments on the quiet chain: builds with parts of the version-6 pipeline left out (ab/x<mask>.so: 1 = no MMA loop,
# 2 = no missing-genotype tails, 4 = no dot REDs, 8 = no row copies) -- what the rest of the pipeline costs per block.
# arguments: <mask>[:<depth>] ...
# builds (here, before gpurun):  mkdir -p ab; for x in 0 1 2 4 15; do nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC \
#     -shared -DBRR_X=$x -o ab/x$x.so bayesrrcmd_b200/csrc/brr_gpu.cu; done      (ab/*.so travel with the snapshot, they are not committed)
for a in "$@"; do
  x=${a%%:*}; d=${a##*:}; [ "$d" == "$a" ] && d=4
  echo -n "x$x depth $d: "
  BRR_GPU_V6_DEPTH=$d BRR_GPU_LIB=$PWD/ab/x$x.so BRR_GPU_DENSE_KERNEL=v6 timeout 120 python tools/bench_quiet.py 2>&1 | tail -1 | cut -c30-110
done
