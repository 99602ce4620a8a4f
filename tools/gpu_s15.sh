set -x
mkdir -p gpurun_out
timeout 600 python tools/run_many.py sweep:188:0,-1,36,37,27 sweep:148:0,-1,36,37,27 sweep:146:0,-1,36,37,27 alexnet:1:0,27,36,37 sweep:85:0,19,20,33,34,35 sweep:31:0,19,20,33,34,35 sweep:83:0,19,20,33,34,35 > gpurun_out/s15_layers.txt 2>&1; cat gpurun_out/s15_layers.txt | cut -c1-170
