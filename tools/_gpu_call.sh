mkdir -p gpurun_out
export BENCH_NO_SMI=1
run() { ( env "$@" timeout 300 python tools/variant_bench.py 20 3 2>&1 | grep -E "VARIANT|Error|error" | tail -2 ) >> gpurun_out/c14_variants.log 2>&1; }
: > gpurun_out/c14_variants.log
run DG_AIR_CFG=6
run DG_AIR_CFG=7
run DG_AIR_CFG=8
cat gpurun_out/c14_variants.log
