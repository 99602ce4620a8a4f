set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/s2_pytest.log 2>&1; tail -5 gpurun_out/s2_pytest.log
timeout 300 python tools/interp_bench.py --from 55 12 30 100 > gpurun_out/s2_interp.txt 2>&1; cat gpurun_out/s2_interp.txt
timeout 600 python tools/run_many.py resnet50:0:0,41,55,56,57,58,59,60,61,62,63,64,65,66,67,68 resnet50:3:0,41,55,56,57,58,59,60,61,62,63,65,67 resnet50:7:0,42,55,56,57,58,59,60,61,62,64,66,67,68 resnet50:13:0,43,55,56,57,58,59,60,61,62,64,66,67,68 alexnet:1:0,46,55,56,57,58,59,60,61,62,64,66,67,68 alexnet:0:0,50,69,70,71,72 > gpurun_out/s2_layers.txt 2>&1; cat gpurun_out/s2_layers.txt | cut -c1-150
