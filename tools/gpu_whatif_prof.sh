# the role counters of the what-if builds (see tools/gpu_whatif.sh).  arguments: <mask>[:<depth>] ...
for a in "$@"; do
  x=${a%%:*}; d=${a##*:}; [ "$d" == "$a" ] && d=4
  echo "== x$x depth $d"
  BRR_GPU_PROFILE=1 BRR_GPU_V6_DEPTH=$d BRR_GPU_LIB=$PWD/ab/x$x.so BRR_GPU_DENSE_KERNEL=v6 timeout 120 python tools/bench_quiet.py 2>&1 | tail -5 | cut -c1-900
done
