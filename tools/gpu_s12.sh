set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "forward" > gpurun_out/s12_pytest.log 2>&1; tail -2 gpurun_out/s12_pytest.log
timeout 600 python tools/run_many.py alexnet:1:0,30,23,4,5,25 alexnet:3:0,30,23,4,5,25 resnet50:7:0,2,4,5,25 resnet50:13:0,31,4,5,25 googlenet:17:0,30,4,5,25 > gpurun_out/s12_layers.txt 2>&1; cat gpurun_out/s12_layers.txt | cut -c1-150
