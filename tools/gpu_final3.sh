set -x
mkdir -p gpurun_out/final
O=gpurun_out/final
timeout 900 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1; tail -2 $O/pytest_gpu.log
python bench.py --steps 10 --warmup 3 > $O/bench_alexnet.json 2> $O/bench_alexnet.err; tail -2 $O/bench_alexnet.err; cut -c1-300 $O/bench_alexnet.json
python bench.py --workload resnet50 --train --steps 5 --warmup 3 --no-cpu > $O/bench_resnet50_train.json 2> $O/bench_resnet50_train.err; tail -2 $O/bench_resnet50_train.err; cut -c1-300 $O/bench_resnet50_train.json
python bench.py --workload alexnet --train --steps 5 --warmup 3 --no-cpu > $O/bench_alexnet_train.json 2> $O/bench_alexnet_train.err; tail -2 $O/bench_alexnet_train.err
python bench.py --workload resnet50 --steps 5 --warmup 3 --no-cpu > $O/bench_resnet50.json 2> $O/bench_resnet50.err; tail -2 $O/bench_resnet50.err
python bench.py --workload googlenet --steps 5 --warmup 3 --no-cpu > $O/bench_googlenet.json 2> $O/bench_googlenet.err; tail -2 $O/bench_googlenet.err
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
