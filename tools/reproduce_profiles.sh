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
  python tools/run_dense.py > $O/dense.txt 2>&1          # -> profiles/r02_dense_tcgen05.txt
  python tools/run_block.py > $O/block.txt 2>&1          # -> profiles/r02_bottleneck_block.txt
  python tools/run_stride2.py > $O/stride2.txt 2>&1      # -> profiles/r02_stride2_space_to_depth.txt
  python tools/run_level1.py > $O/level1.txt 2>&1        # -> profiles/r02_level1_vs_level2.txt
  ;;
r02sweep) # 1 GPU, ~10 min: BASELINE configs[4] + every variant x layout against the generic kernel at full size
  O=gpurun_out/r02b
  mkdir -p $O
  python tools/sweep.py --autotune > $O/sweep_autotuned.jsonl    # -> profiles/r02_sweep_config5_autotuned.jsonl (+ _summary.txt)
  for a in "alexnet:0 256" "alexnet:1 256" "alexnet:2 256" "alexnet:3 256" "resnet50:0 64" "resnet50:3 256" "resnet50:7 256" \
           "resnet50:13 256" "googlenet:0 128" "googlenet:1 128" "googlenet:2 128" "googlenet:6 128" "googlenet:12 128" "googlenet:17 128"; do
    set -- $a; python tools/check_variants.py $1 $2
  done > $O/check_variants.txt 2>&1                              # -> profiles/r02_check_variants_full_size.txt
  ;;
r02multi) # N GPUs (gpurun --gpus N): the default command under torchrun, as the driver's scaling run launches it
  N=${2:-8}
  O=gpurun_out/r02b
  mkdir -p $O
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 \
      bench.py --gpus $N --steps 10 --warmup 3 > $O/bench_alexnet_${N}gpu.json 2> $O/bench_alexnet_${N}gpu.err
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
