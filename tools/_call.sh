set -x
(timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -25) > gpurun_out/c2_pytest.log
tail -5 gpurun_out/c2_pytest.log
timeout 600 python tools/bench_kernels.py 32 gpurun_out/c2_kernels.json > gpurun_out/c2_kernels.log 2>&1
tail -40 gpurun_out/c2_kernels.log
timeout 900 python bench.py > gpurun_out/c2_bench.json 2> gpurun_out/c2_bench.err
cat gpurun_out/c2_bench.json; tail -5 gpurun_out/c2_bench.err
