mkdir -p gpurun_out
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29546 tools/multi_gpu_check.py 2>&1 | grep -E "MULTI_GPU_CHECK|identical=False|rror" | tail -5 ) > gpurun_out/c12_check4.log 2>&1
cat gpurun_out/c12_check4.log
( DG_SUBSTAGE=1 BENCH_NO_SMI=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 4 --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | grep -E "^\{|SUBSTAGE|rror" | cut -c1-700 | tail -3 ) > gpurun_out/c12_bench4.log 2>&1
cat gpurun_out/c12_bench4.log
( timeout 600 python tools/single_process_check.py 4 14 2>&1 | tail -3 ) > gpurun_out/c12_single4.log 2>&1
cat gpurun_out/c12_single4.log
