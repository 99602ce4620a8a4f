mkdir -p gpurun_out/final
timeout 1200 python tools/sweep.py > gpurun_out/final/sweep.jsonl 2> gpurun_out/final/sweep.err; tail -3 gpurun_out/final/sweep.err; wc -l gpurun_out/final/sweep.jsonl; head -3 gpurun_out/final/sweep.jsonl | cut -c1-400
