for ci in 3 4 6 8 10; do echo "== W_CI=$ci"; ESCORT_W_CI=$ci python tools/run_bwd.py resnet50:0 resnet50:7 resnet50:13 alexnet:1 2>&1 | cut -c100-175; done
