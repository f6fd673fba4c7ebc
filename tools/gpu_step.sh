# one development step on the GPU: version-6 parity tests of the tree's build, then quiet-chain timings (what-if builds in ab/)
TAG=${1:-st}; shift
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dense.py -m gpu -x -q -k "v6" --timeout 150 -o timeout_method=thread 2>&1 | tail -8 > gpurun_out/${TAG}_pytest.log
echo "== tests: $(tail -1 gpurun_out/${TAG}_pytest.log)"
grep -q "failed\|error" gpurun_out/${TAG}_pytest.log && tail -8 gpurun_out/${TAG}_pytest.log
bash tools/gpu_whatif.sh "$@"
