set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/s6_pytest.log 2>&1; tail -5 gpurun_out/s6_pytest.log
timeout 600 python tools/run_many.py resnet50:0:0,40,41,74,75,76,77 resnet50:3:0,40,74,76 resnet50:7:0,41,75,77,78,79 resnet50:13:0,42,78,79 alexnet:1:0,45,78,79 alexnet:0:0,49,80 googlenet:0:0,40,74,75 > gpurun_out/s6_layers.txt 2>&1; cat gpurun_out/s6_layers.txt | cut -c1-150
