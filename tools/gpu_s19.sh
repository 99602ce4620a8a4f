set -x
timeout 600 python tools/run_many.py resnet50:7:0,2,4,4r1,4r2,4r3,4r4,4r5 alexnet:1:0,4,4r1,4r2,4r3,4r4 resnet50:13:0,4,4r1,4r2,4r3 2>&1 | cut -c1-210
