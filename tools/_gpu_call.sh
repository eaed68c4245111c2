mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -25 ) > gpurun_out/c3_pytest.log 2>&1
cat gpurun_out/c3_pytest.log
export BENCH_NO_SMI=1
run() { ( env "$@" timeout 300 python tools/variant_bench.py 20 3 2>&1 | grep -E "VARIANT|SUBSTAGE|Error|error" | tail -3 ) >> gpurun_out/c3_variants.log 2>&1; }
: > gpurun_out/c3_variants.log
run DG_X=default
run DG_SCAN_CHAINED=0
run DG_SUBSTAGE=1
( DG_SUBSTAGE=1 timeout 300 python tools/variant_bench.py 8 3 2>&1 | grep -E "VARIANT|SUBSTAGE|rror" | tail -2 ) >> gpurun_out/c3_variants.log 2>&1
( DG_SUBSTAGE=1 timeout 300 python tools/variant_bench.py 14 3 2>&1 | grep -E "VARIANT|SUBSTAGE|rror" | tail -2 ) >> gpurun_out/c3_variants.log 2>&1
cat gpurun_out/c3_variants.log
unset BENCH_NO_SMI
( DG_SUBSTAGE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --steps 5 --warmup 3 2>&1 | grep -E "^\{|SUBSTAGE|rror" | tail -4 | cut -c1-1800 ) > gpurun_out/c3_bench2.log 2>&1
cat gpurun_out/c3_bench2.log
