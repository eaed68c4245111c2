mkdir -p gpurun_out
export BENCH_NO_SMI=1
( timeout 300 python tools/variant_bench.py 20 3 2>&1 | grep -E "VARIANT|rror" | tail -2 ) > gpurun_out/c16_variants.log 2>&1
cat gpurun_out/c16_variants.log
( timeout 900 python -m pytest tests/test_gpu_blocks.py tests/test_gpu_prove.py -x -q -m gpu 2>&1 | tail -4 ) > gpurun_out/c16_pytest.log 2>&1
cat gpurun_out/c16_pytest.log
tag=r02d; out=gpurun_out; tmp=/tmp/ncu_$tag; mkdir -p $tmp
ncu --set full --clock-control none --import-source on -k regex:ntt_pass_kernel -s 18 -c 2 -o $tmp/ntt -f python tools/prove_once.py 20 2 > $out/${tag}_prove.log 2>&1
ncu -i $tmp/ntt.ncu-rep --page raw --csv > $out/${tag}_ntt_raw.csv 2>/dev/null
ls -la $out/${tag}_ntt_raw.csv
