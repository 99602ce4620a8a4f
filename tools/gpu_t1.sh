mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/t1_pytest.log 2>&1; tail -2 gpurun_out/t1_pytest.log
timeout 600 python tools/run_many.py resnet50:0:0,23,25,27,28 resnet50:3:0,23,25,27,28 googlenet:0:0,23,27 googlenet:3:0,23,27 2>&1 | cut -c1-130
for v in sconv_tile_wa_o3_y7_x4_k3x3_s1_w12_r152 sconv_tile_ws_o3_y7_x4_k3x3_s1_w12_r152 sconv_tile_wa_o4_y7_x4_k3x3_s1_w8_r232 sconv_tile_ws_o4_y7_x4_k3x3_s1_w8_r232; do ESCORT_BWDW_VARIANT=$v python tools/run_bwd.py resnet50:0 resnet50:3 2>&1 | cut -c95-160; done
