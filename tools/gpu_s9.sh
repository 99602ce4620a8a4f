set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k backward > gpurun_out/s9_pytest.log 2>&1; tail -3 gpurun_out/s9_pytest.log
timeout 600 python tools/run_bwd.py resnet50:0 resnet50:3 resnet50:7 resnet50:13 alexnet:0 alexnet:1 > gpurun_out/s9_bwd.txt 2>&1; cat gpurun_out/s9_bwd.txt | cut -c1-200
ESCORT_BWDW_VARIANT=sconv_tile_wa_o4_y7_x4_k3x3_s1_w8_r232 timeout 600 python tools/run_bwd.py resnet50:0 resnet50:3 2>&1 | cut -c1-200
ESCORT_BWDW_VARIANT=sconv_tile_wb_o4_y7_x4_k3x3_s1_w8_r232 timeout 600 python tools/run_bwd.py resnet50:7 resnet50:13 alexnet:1 2>&1 | cut -c1-200
ESCORT_BWDW_VARIANT=sconv_tile_wb_o3_y7_x4_k3x3_s1_w12_r152 timeout 600 python tools/run_bwd.py resnet50:0 resnet50:13 alexnet:1 2>&1 | cut -c1-200
ESCORT_BWDW_VARIANT=sconv_tile_wb_o4_y4_x4_k3x3_s1_w12_r152 timeout 600 python tools/run_bwd.py resnet50:7 resnet50:13 2>&1 | cut -c1-200
