set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/s13_pytest.log 2>&1; tail -3 gpurun_out/s13_pytest.log
timeout 600 python tools/run_many.py resnet50:13:0,28,29 googlenet:15:0,28,29 googlenet:17:0,28,29 alexnet:1:0,28,29 > gpurun_out/s13_layers.txt 2>&1; cat gpurun_out/s13_layers.txt | cut -c1-150
python tools/run_bwd.py resnet50:13 2>&1 | cut -c1-200
ESCORT_BWDW_VARIANT=sconv_tile_wb_o5_y4_x4_k3x3_s1_w12_r152 python tools/run_bwd.py resnet50:13 2>&1 | cut -c1-200
