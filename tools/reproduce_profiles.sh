#!/bin/bash
# How profiles/r01_final_* were produced (each block is one `gpurun -- 'bash tools/reproduce_profiles.sh <block>'` call).
set -x
O=gpurun_out/final
mkdir -p $O
case "${1:-bench}" in
bench)   # 1 GPU: parity tests, the bench lines, the reference arm
  python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1; tail -2 $O/pytest_gpu.log
  python bench.py --steps 10 --warmup 3 --tune-cache $O/tune_alexnet.json > $O/bench_alexnet.json
  python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_alexnet_reference.json
  for w in resnet50 googlenet lenet; do python bench.py --workload $w --steps 5 --warmup 3 --no-cpu > $O/bench_$w.json; done
  for w in resnet50 alexnet; do python bench.py --workload $w --train --steps 5 --warmup 3 --no-cpu > $O/bench_${w}_train.json; done
  ;;
ncu)     # 1 GPU: launch list of the default bench command + one full capture per kernel of the step
  python bench.py --steps 2 --warmup 3 --no-cpu --tune-cache $O/tune_alexnet.json > /dev/null
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/ncu_launches_alexnet_bench.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu --tune-cache $O/tune_alexnet.json > $O/ncu_launch.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:sconv_tile_kernel -s 12 -c 4 -f -o $O/ncu_full_alexnet_step \
      python bench.py --steps 2 --warmup 3 --no-cpu --tune-cache $O/tune_alexnet.json > $O/ncu_full.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:sconv_tile_bwdw -s 1 -c 1 -f -o $O/ncu_full_res2a_bwdw \
      python tools/run_bwd.py resnet50:0 > $O/ncu_bwdw.log 2>&1
  # then, on the CPU box: python tools/ncu_summary.py $O/<rep>.ncu-rep --top 16 > profiles/r01_final_ncu_<...>_stalls.txt
  ;;
sweep)   # 1 GPU: BASELINE configs[4]
  python tools/sweep.py > $O/sweep.jsonl
  python tools/sweep.py --autotune > $O/sweep_autotuned.jsonl
  ;;
r02)     # 1 GPU, round 2: tests, bench lines, launch list + one full capture per kernel of the default step
  O=gpurun_out/r02
  mkdir -p $O
  python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1; tail -2 $O/pytest_gpu.log
  python bench.py --steps 10 --warmup 3 --tune-cache $O/tune_alexnet.json > $O/bench_alexnet.json 2> $O/bench_alexnet.err
  python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_alexnet_reference.json
  for w in resnet50 googlenet lenet; do python bench.py --workload $w --steps 5 --warmup 3 --no-cpu > $O/bench_$w.json 2> $O/bench_$w.err; done
  python bench.py --workload resnet50 --train --steps 5 --warmup 3 --no-cpu > $O/bench_resnet50_train.json 2> $O/bench_resnet50_train.err
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/ncu_launches_alexnet_bench.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu --no-train-leg --tune-cache $O/tune_alexnet.json > $O/ncu_launch.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:sconv_tmem -s 12 -c 4 -f -o $O/ncu_full_alexnet_step \
      python bench.py --steps 2 --warmup 3 --no-cpu --no-train-leg --no-parity --tune-cache $O/tune_alexnet.json > $O/ncu_full.log 2>&1
  ;;
r02b)    # 1 GPU, end of round 2: everything above on the final code + the dense (f1) kernels
  O=gpurun_out/r02b
  mkdir -p $O
  python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1; tail -2 $O/pytest_gpu.log
  python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_alexnet_reference.json
  python bench.py --steps 10 --warmup 3 --tune-cache $O/tune_alexnet.json > $O/bench_alexnet.json 2> $O/bench_alexnet.err
  for w in resnet50 googlenet lenet; do python bench.py --workload $w --steps 5 --warmup 3 --no-cpu > $O/bench_$w.json 2> $O/bench_$w.err; done
  python bench.py --workload resnet50 --train --steps 5 --warmup 3 --no-cpu > $O/bench_resnet50_train.json 2> $O/bench_resnet50_train.err
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/ncu_launches_alexnet_bench.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu --no-train-leg --tune-cache $O/tune_alexnet.json > $O/ncu_launch.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:sconv_tmem -s 12 -c 4 -f -o $O/ncu_full_alexnet_step \
      python bench.py --steps 2 --warmup 3 --no-cpu --no-train-leg --no-parity --tune-cache $O/tune_alexnet.json > $O/ncu_full.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:dense_conv_tf32 -c 22 -f -o $O/ncu_full_dense_conv \
      python tools/run_dense.py > $O/ncu_dense.log 2>&1
  for tool in memcheck racecheck synccheck; do
    compute-sanitizer --tool $tool python tools/sanitize_dense.py > $O/sanitizer_dense_$tool.log 2>&1
    grep -E "rel_l2|ERROR SUMMARY|RACECHECK SUMMARY" $O/sanitizer_dense_$tool.log | tail -9
  done
  python tools/small_time.py > $O/small_maps.txt 2>&1
  ;;
sanitize) # 1 GPU, round 2: compute-sanitizer on a thin layer through one kernel of each family (generic, cp.async and
          # TMA tile variants, 51 TMEM producer/consumer, 59 TMEM self-fill) + the default backward kernels
  O=gpurun_out/r02
  mkdir -p $O
  for tool in memcheck racecheck synccheck; do
    compute-sanitizer --tool $tool python tools/sanitize_case.py 0 2 11 51 59 > $O/sanitizer_$tool.log 2>&1
    SANITIZE_H=28 compute-sanitizer --tool $tool python tools/sanitize_case.py 4 6 27 51 59 >> $O/sanitizer_$tool.log 2>&1
    grep -E "forward|backward|ERROR SUMMARY|RACECHECK SUMMARY" $O/sanitizer_$tool.log | tail -16
  done
  ;;
multi)   # N GPUs (gpurun --gpus N): the same bench under torchrun
  N=${2:-2}
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 \
      bench.py --gpus $N --steps 10 --warmup 3 > $O/bench_alexnet_${N}gpu.json
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 \
      bench.py --gpus $N --workload resnet50 --train --steps 5 --warmup 3 > $O/bench_resnet50_train_${N}gpu.json
  ;;
esac
