mkdir -p gpurun_out
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29546 tools/multi_gpu_check.py 2>&1 | grep -E "MULTI_GPU_CHECK|identical=False|rror" | tail -5 ) > gpurun_out/c4_check8.log 2>&1
cat gpurun_out/c4_check8.log
( DG_SUBSTAGE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 8 --steps 4 --warmup 3 2>&1 | grep -E "^\{|SUBSTAGE|rror" | tail -12 | cut -c1-2500 ) > gpurun_out/c4_bench8.log 2>&1
cat gpurun_out/c4_bench8.log
