mkdir -p gpurun_out
( DG_SUBSTAGE=2 BENCH_NO_SMI=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 8 --steps 4 --warmup 3 --no-cpu-baseline 2>&1 | grep -E "^\{|SUBSTAGE|rror" ) > gpurun_out/c13_bench8.log 2>&1
grep "^{" gpurun_out/c13_bench8.log | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print(d['value'], d['ms_steps'], d['e2e']['value'], d['e2e']['pageable_ms'], [round(x,2) for x in d['stage_ms']])"
grep SUBSTAGE gpurun_out/c13_bench8.log | sed -n 41,48p | cut -c1-330
