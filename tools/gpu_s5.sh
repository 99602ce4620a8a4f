set -x
mkdir -p gpurun_out
timeout 600 python tools/run_many.py resnet50:0:0,32,33,40,41,56 resnet50:3:0,32,40,56 resnet50:7:0,33,41,42,56 resnet50:13:0,42,43,56 alexnet:1:0,45,41,56 alexnet:0:0,49,71 alexnet:2:0,45,56 googlenet:0:0,32,40 > gpurun_out/s5_layers.txt 2>&1; cat gpurun_out/s5_layers.txt | cut -c1-150
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sconv_tile -s 2 -c 1 -f -o gpurun_out/s5_res2a_v40 python tools/run_many.py resnet50:0:40 > gpurun_out/s5_a.log 2>&1; tail -2 gpurun_out/s5_a.log
