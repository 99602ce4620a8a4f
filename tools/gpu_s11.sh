set -x
mkdir -p gpurun_out
timeout 600 python tools/run_many.py alexnet:0:0,15,29,30,31,22,3 googlenet:4:0,15,29,30,31,22 googlenet:14:0,15,29,30,31,22 googlenet:18:0,29,30,31,22 > gpurun_out/s11_layers.txt 2>&1; cat gpurun_out/s11_layers.txt | cut -c1-150
for v in sconv_tile_wb_o4_y4_x4_k5x5_s1_w12_r152 sconv_tile_wb_o2_y4_x4_k5x5_s1_w12_r152 sconv_tile_wb_o3_y4_x4_k5x5_s1_w12_r152; do ESCORT_BWDW_VARIANT=$v python tools/run_bwd.py alexnet:0 2>&1 | cut -c60-200; done
