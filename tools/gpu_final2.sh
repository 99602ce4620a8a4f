set -x
mkdir -p gpurun_out/final
O=gpurun_out/final
rm -f $O/tune_*.json
timeout 600 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1; tail -2 $O/pytest_gpu.log
python bench.py --steps 10 --warmup 3 --tune-cache $O/tune_alexnet.json > $O/bench_alexnet.json 2> $O/bench_alexnet.err; tail -2 $O/bench_alexnet.err
python bench.py --workload googlenet --steps 5 --warmup 3 --no-cpu --tune-cache $O/tune_googlenet.json > $O/bench_googlenet.json 2> $O/bench_googlenet.err; tail -2 $O/bench_googlenet.err
python bench.py --workload alexnet --train --steps 5 --warmup 3 --no-cpu --tune-cache $O/tune_alexnet.json > $O/bench_alexnet_train.json 2> $O/bench_alexnet_train.err; tail -2 $O/bench_alexnet_train.err
python bench.py --workload lenet --steps 5 --warmup 3 --tune-cache $O/tune_lenet.json > $O/bench_lenet.json 2> $O/bench_lenet.err; tail -2 $O/bench_lenet.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/ncu_launches_alexnet_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu --tune-cache $O/tune_alexnet.json > $O/ncu_launch.log 2>&1; tail -1 $O/ncu_launch.log | cut -c1-200
ncu --set full --clock-control none --import-source on -k regex:sconv_tile_kernel -s 12 -c 4 -f -o $O/ncu_full_alexnet_step python bench.py --steps 2 --warmup 3 --no-cpu --tune-cache $O/tune_alexnet.json > $O/ncu_full.log 2>&1; tail -1 $O/ncu_full.log | cut -c1-200
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
