set -x
mkdir -p gpurun_out/final
O=gpurun_out/final
python bench.py --steps 10 --warmup 3 --tune-cache $O/tune_alexnet.json > $O/bench_alexnet.json 2> $O/bench_alexnet.err; tail -2 $O/bench_alexnet.err
python bench.py --workload resnet50 --steps 5 --warmup 3 --no-cpu --tune-cache $O/tune_resnet50.json > $O/bench_resnet50.json 2> $O/bench_resnet50.err; tail -2 $O/bench_resnet50.err
python bench.py --workload googlenet --steps 5 --warmup 3 --no-cpu --tune-cache $O/tune_googlenet.json > $O/bench_googlenet.json 2> $O/bench_googlenet.err; tail -2 $O/bench_googlenet.err
python bench.py --workload resnet50 --train --steps 5 --warmup 3 --no-cpu --tune-cache $O/tune_resnet50.json > $O/bench_resnet50_train.json 2> $O/bench_resnet50_train.err; tail -2 $O/bench_resnet50_train.err
python bench.py --workload alexnet --train --steps 5 --warmup 3 --no-cpu --tune-cache $O/tune_alexnet.json > $O/bench_alexnet_train.json 2> $O/bench_alexnet_train.err; tail -2 $O/bench_alexnet_train.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_alexnet_reference.json 2> $O/bench_ref.err; tail -2 $O/bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/ncu_launches_alexnet_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu --tune-cache $O/tune_alexnet.json > $O/ncu_launch.log 2>&1; tail -1 $O/ncu_launch.log | cut -c1-200
ncu --set full --clock-control none --import-source on -k regex:sconv_tile_kernel -s 12 -c 4 -f -o $O/ncu_full_alexnet_step python bench.py --steps 2 --warmup 3 --no-cpu --tune-cache $O/tune_alexnet.json > $O/ncu_full.log 2>&1; tail -1 $O/ncu_full.log | cut -c1-200
ncu --set full --clock-control none --import-source on -k regex:sconv_tile_bwdw -s 1 -c 1 -f -o $O/ncu_full_res2a_bwdw python tools/run_bwd.py resnet50:0 > $O/ncu_bwdw.log 2>&1; tail -1 $O/ncu_bwdw.log | cut -c1-200
ls -la $O
