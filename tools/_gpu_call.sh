set -x
mkdir -p gpurun_out
nvidia-smi -L
( timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 ) > gpurun_out/c2_pytest.log 2>&1
cat gpurun_out/c2_pytest.log
export BENCH_NO_SMI=1
run() { ( env "$@" timeout 300 python tools/variant_bench.py 20 3 2>&1 | grep -E "VARIANT|Error|error" ) >> gpurun_out/c2_variants.log 2>&1; }
: > gpurun_out/c2_variants.log
run DG_X=default
run DG_NTT_B=2,512,1
run DG_NTT_A=2,512,1
run DG_NTT_A=2,512,1 DG_NTT_B=2,512,1 DG_NTT_AF=3,512,1
run DG_NTT_TILE=13 DG_NTT_A=3,1024,1 DG_NTT_B=3,1024,1 DG_NTT_AF=3,1024,1
run DG_NO_STAGING=1
cat gpurun_out/c2_variants.log
unset BENCH_NO_SMI
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --steps 5 --warmup 3 2>&1 | tail -3 ) > gpurun_out/c2_bench2.log 2>&1
tail -c 3000 gpurun_out/c2_bench2.log
( timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -2 ) > gpurun_out/c2_bench1.log 2>&1
tail -c 3500 gpurun_out/c2_bench1.log
python - <<'PY' > gpurun_out/c2_oracle_threads.log 2>&1
import time, sys
sys.path.insert(0, '.')
import bench
from oracle import pyoracle as po
tr, name = bench.build_trace(16)
for T in (1, 8, 16, 32, 64, 128):
    po.set_threads(T)
    t0 = time.time()
    r = po.prove(tr.registers, tr.ctx_depth, tr.loop_depth, tr.public_inputs, tr.outputs)
    print(T, round(time.time() - t0, 2), [round(x) for x in r.stage_ms], flush=True)
PY
cat gpurun_out/c2_oracle_threads.log
