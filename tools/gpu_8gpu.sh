set -x
mkdir -p gpurun_out/final
O=gpurun_out/final
nvidia-smi -L | wc -l
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 > $O/bench_alexnet_8gpu.json 2> $O/bench_alexnet_8gpu.err; tail -2 $O/bench_alexnet_8gpu.err; cut -c1-260 $O/bench_alexnet_8gpu.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --workload resnet50 --train --steps 5 --warmup 3 > $O/bench_resnet50_train_8gpu.json 2> $O/bench_resnet50_train_8gpu.err; tail -2 $O/bench_resnet50_train_8gpu.err; cut -c1-260 $O/bench_resnet50_train_8gpu.json
