mkdir -p gpurun_out/final
timeout 1500 python tools/sweep.py --autotune > gpurun_out/final/sweep_autotuned.jsonl 2> gpurun_out/final/sweep2.err; tail -3 gpurun_out/final/sweep2.err; wc -l gpurun_out/final/sweep_autotuned.jsonl
