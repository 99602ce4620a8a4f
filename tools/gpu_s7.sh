set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/s7_pytest.log 2>&1; tail -15 gpurun_out/s7_pytest.log
