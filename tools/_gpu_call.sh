mkdir -p gpurun_out
export BENCH_NO_SMI=1
run() { ( env "$@" timeout 300 python tools/variant_bench.py 20 3 2>&1 | grep -E "VARIANT|rror" | tail -2 ) >> gpurun_out/c17_variants.log 2>&1; }
: > gpurun_out/c17_variants.log
run DG_LDE_PREFOLD=1
run DG_LDE_PREFOLD=0
cat gpurun_out/c17_variants.log
( timeout 900 python -m pytest tests/test_gpu_prove.py tests/test_gpu_baseline_sizes.py -x -q -m gpu 2>&1 | tail -4 ) > gpurun_out/c17_pytest.log 2>&1
cat gpurun_out/c17_pytest.log
