mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "forward or backward" > gpurun_out/t2_pytest.log 2>&1; tail -2 gpurun_out/t2_pytest.log
timeout 600 python tools/run_many.py alexnet:1:0,32,30 alexnet:2:0,32,30 alexnet:3:0,32,30 resnet50:13:0,34,31 resnet50:7:0,24,29 resnet50:0:0,24,29 googlenet:17:0,32,30 2>&1 | cut -c1-130
for v in sconv_tile_wb_o4_y4_x4_k3x3_s1_w12_r152 sconv_tile_wt_o4_y4_x4_k3x3_s1_w12_r152; do ESCORT_BWDW_VARIANT=$v python tools/run_bwd.py alexnet:1 2>&1 | cut -c95-160; done
for v in sconv_tile_wb_o5_y4_x4_k3x3_s1_w12_r152 sconv_tile_wt_o5_y4_x4_k3x3_s1_w12_r152; do ESCORT_BWDW_VARIANT=$v python tools/run_bwd.py resnet50:13 2>&1 | cut -c95-160; done
for v in sconv_tile_wb_o3_y7_x4_k3x3_s1_w12_r152 sconv_tile_wt_o3_y7_x4_k3x3_s1_w12_r152; do ESCORT_BWDW_VARIANT=$v python tools/run_bwd.py resnet50:7 2>&1 | cut -c95-160; done
