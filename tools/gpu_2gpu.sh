set -x
mkdir -p gpurun_out/final
O=gpurun_out/final
nvidia-smi -L
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_alexnet_2gpu.json 2> $O/bench_alexnet_2gpu.err; tail -3 $O/bench_alexnet_2gpu.err; cut -c1-400 $O/bench_alexnet_2gpu.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload resnet50 --train --steps 5 --warmup 3 > $O/bench_resnet50_train_2gpu.json 2> $O/bench_resnet50_train_2gpu.err; tail -3 $O/bench_resnet50_train_2gpu.err; cut -c1-700 $O/bench_resnet50_train_2gpu.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --impl reference --steps 1 --warmup 1 > $O/bench_ref_2gpu.json 2> $O/bench_ref_2gpu.err; cut -c1-300 $O/bench_ref_2gpu.json
