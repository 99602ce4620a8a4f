set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/s16_pytest.log 2>&1; tail -3 gpurun_out/s16_pytest.log
timeout 600 python tools/run_bwd.py resnet50:0 resnet50:3 resnet50:7 resnet50:13 alexnet:0 alexnet:1 2>&1 | cut -c1-200 | tee gpurun_out/s16_bwd.txt
