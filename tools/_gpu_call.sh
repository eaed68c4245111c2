mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -25 ) > gpurun_out/c5_pytest.log 2>&1
cat gpurun_out/c5_pytest.log
export BENCH_NO_SMI=1
( timeout 300 python tools/variant_bench.py 20 3 2>&1 | grep -E "VARIANT|SUBSTAGE|Error|error" | tail -3 ) > gpurun_out/c5_variants.log 2>&1
cat gpurun_out/c5_variants.log
unset BENCH_NO_SMI
( DG_SUBSTAGE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --steps 4 --warmup 3 --no-cpu-baseline 2>&1 | grep -E "^\{|SUBSTAGE|rror" | tail -3 | cut -c1-1500 ) > gpurun_out/c5_bench2.log 2>&1
cat gpurun_out/c5_bench2.log
