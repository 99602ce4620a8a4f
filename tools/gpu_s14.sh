set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "forward" > gpurun_out/s14_pytest.log 2>&1; tail -2 gpurun_out/s14_pytest.log
timeout 600 python tools/run_many.py alexnet:1:0,27,33,34,35 alexnet:3:0,27,33,34,35 resnet50:13:0,29,36,37 googlenet:17:0,27,36,37 resnet50:7:0,2,38,39 > gpurun_out/s14_layers.txt 2>&1; cat gpurun_out/s14_layers.txt | cut -c1-170
