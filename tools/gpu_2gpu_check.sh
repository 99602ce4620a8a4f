mkdir -p gpurun_out/final
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/final/check_2gpu.out 2> gpurun_out/final/check_2gpu.err
echo "stdout lines: $(wc -l < gpurun_out/final/check_2gpu.out)"; head -c 200 gpurun_out/final/check_2gpu.out; echo; grep -c "NCCL version" gpurun_out/final/check_2gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --impl reference --steps 1 --warmup 1 > gpurun_out/final/check_2gpu_ref.out 2>/dev/null; echo "ref stdout lines: $(wc -l < gpurun_out/final/check_2gpu_ref.out)"; head -c 150 gpurun_out/final/check_2gpu_ref.out
