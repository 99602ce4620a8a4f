set -x
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sconv_tile_kernel -s 2 -c 1 -f -o gpurun_out/s18_res4a_interp_p2 python tools/run_many.py resnet50:7:2 > gpurun_out/s18.log 2>&1; tail -2 gpurun_out/s18.log | cut -c1-200
