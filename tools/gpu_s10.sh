set -x
mkdir -p gpurun_out
timeout 600 python tools/run_many.py resnet50:0:0,29,30,31,32,33 resnet50:3:0,29,30,31,32,33 resnet50:7:0,29,31,32 resnet50:13:0,29,31,32 alexnet:1:0,29,31,32 > gpurun_out/s10_layers.txt 2>&1; cat gpurun_out/s10_layers.txt | cut -c1-150
