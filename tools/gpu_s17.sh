set -x
mkdir -p gpurun_out
timeout 600 python tools/run_many.py alexnet:0:0,22,23,24 googlenet:4:0,22,23,24 googlenet:14:0,22,23,24 > gpurun_out/s17_layers.txt 2>&1; cat gpurun_out/s17_layers.txt | cut -c1-170
