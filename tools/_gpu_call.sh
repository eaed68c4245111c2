mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -40 ) > gpurun_out/c6_pytest.log 2>&1
cat gpurun_out/c6_pytest.log
( timeout 600 python bench.py --microbench --verbose 2> gpurun_out/c6_micro_verbose.log | tail -1 ) > gpurun_out/c6_microbench_n1.json
cat gpurun_out/c6_micro_verbose.log | tail -40
