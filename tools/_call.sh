set -x
(timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -25) > gpurun_out/c13_pytest.log
tail -12 gpurun_out/c13_pytest.log
