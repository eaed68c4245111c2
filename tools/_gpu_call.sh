mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 ) > gpurun_out/c18_pytest.log 2>&1
cat gpurun_out/c18_pytest.log
( timeout 900 python bench.py 2>&1 | tail -1 ) > gpurun_out/c18_bench1.json
( BENCH_NO_SMI=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | grep "^{" | tail -1 ) > gpurun_out/c18_bench2.json
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29545 bench.py --gpus 2 --microbench --quick 2>&1 | grep "^{" | tail -1 ) > gpurun_out/c18_microbench_n2.json
python - <<'PY'
import json
for f in ("c18_bench1", "c18_bench2"):
    d = json.loads(open("gpurun_out/%s.json" % f).read())
    print(f, round(d["value"], 2), round(d["e2e"]["value"], 2), round(d["e2e"]["pageable_ms"], 2), d["gpu_launches"], d["proof_check"]["matches_oracle_golden"], [round(x, 2) for x in d["stage_ms"]], d["roofline"]["frac"], d["roofline"]["traffic"], d["clocks"])
m = json.loads(open("gpurun_out/c18_microbench_n2.json").read())
for r in m["microbench"]:
    print(r["kernel"], r["shape"], round(r["ms"], 3), round(r["alg_gbs_aggregate"], 1))
PY
