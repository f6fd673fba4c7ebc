# timing experiments on the quiet chain: builds with parts of the version-6 pipeline left out (ab/x<mask>.so: 1 = no MMA loop,
# 2 = no missing-genotype tails, 4 = no dot REDs, 8 = no row copies) -- what the rest of the pipeline costs per block.
# arguments: <mask>[:<depth>] ...
for a in "$@"; do
  x=${a%%:*}; d=${a##*:}; [ "$d" == "$a" ] && d=4
  echo -n "x$x depth $d: "
  BRR_GPU_V6_DEPTH=$d BRR_GPU_LIB=$PWD/ab/x$x.so BRR_GPU_DENSE_KERNEL=v6 timeout 120 python tools/bench_quiet.py 2>&1 | tail -1 | cut -c30-110
done
