set -x
(timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -25) > gpurun_out/c6_pytest.log
tail -8 gpurun_out/c6_pytest.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/c6_bench.json 2> gpurun_out/c6_bench.err
cat gpurun_out/c6_bench.json | cut -c1-400; tail -3 gpurun_out/c6_bench.err
timeout 600 python tools/bench_kernels.py 32 gpurun_out/c6_kernels.json 2>&1 | grep "^| gemm\|^| attention\|^| nms\|^| ransac\|^| conv1a\|^| layernorm" 
