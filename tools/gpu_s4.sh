set -x
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sconv_tile -s 2 -c 1 -f -o gpurun_out/s4_res2a_v56 python tools/run_many.py resnet50:0:56 > gpurun_out/s4_a.log 2>&1; tail -2 gpurun_out/s4_a.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sconv_tile -s 2 -c 1 -f -o gpurun_out/s4_res2a_v41 python tools/run_many.py resnet50:0:41 > gpurun_out/s4_b.log 2>&1; tail -2 gpurun_out/s4_b.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sconv_tile -s 2 -c 1 -f -o gpurun_out/s4_conv3_v56 python tools/run_many.py alexnet:1:56 > gpurun_out/s4_c.log 2>&1; tail -2 gpurun_out/s4_c.log
