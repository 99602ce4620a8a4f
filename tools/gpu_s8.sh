set -x
mkdir -p gpurun_out
( time python bench.py --steps 5 --warmup 3 --tune-cache gpurun_out/tune_alexnet.json > gpurun_out/s8_bench_alexnet.json 2> gpurun_out/s8_bench_alexnet.err ) 2>&1 | tail -4; tail -3 gpurun_out/s8_bench_alexnet.err; cat gpurun_out/s8_bench_alexnet.json | cut -c1-1500
( time python bench.py --workload resnet50 --train --steps 3 --warmup 3 --no-cpu --tune-cache gpurun_out/tune_resnet50.json > gpurun_out/s8_bench_resnet50_train.json 2> gpurun_out/s8_bench_resnet50_train.err ) 2>&1 | tail -4; tail -3 gpurun_out/s8_bench_resnet50_train.err; cat gpurun_out/s8_bench_resnet50_train.json | cut -c1-1200
