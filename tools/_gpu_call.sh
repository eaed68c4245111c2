mkdir -p gpurun_out
export BENCH_NO_SMI=1
run() { ( env "$@" timeout 300 python tools/variant_bench.py 20 3 2>&1 | grep -E "VARIANT|Error|error" | tail -2 ) >> gpurun_out/c9_variants.log 2>&1; }
: > gpurun_out/c9_variants.log
run DG_X=default
run DG_NTT_ZFAST=0
run DG_NTT_SCRATCH_MB=1024
run DG_NTT_SCRATCH_MB=1024 DG_NTT_ZFAST=0
run DG_NTT_SCRATCH_MB=14000
cat gpurun_out/c9_variants.log
unset BENCH_NO_SMI
bash tools/capture_profiles.sh r02b 20
( timeout 900 python bench.py --steps 5 --warmup 3 2>&1 | tail -1 ) > gpurun_out/c9_bench1.json
cut -c1-700 gpurun_out/c9_bench1.json
