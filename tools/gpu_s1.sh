set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/s1_pytest.log 2>&1; tail -5 gpurun_out/s1_pytest.log
timeout 300 python tools/interp_bench.py --from 32 12 30 100 > gpurun_out/s1_interp.txt 2>&1; cat gpurun_out/s1_interp.txt
timeout 600 python tools/run_many.py resnet50:0:0,21,32,33,34,38,40 resnet50:3:0,21,32,33,34,38,40 resnet50:7:0,5,33,35,37,39,41,42,43,45,46 resnet50:13:0,33,35,37,41,42,43,44,45,46 alexnet:1:0,33,35,37,41,42,43,44,45,46 alexnet:0:0,48,49,50 googlenet:0:0,32,34,36 > gpurun_out/s1_layers.txt 2>&1; cat gpurun_out/s1_layers.txt
