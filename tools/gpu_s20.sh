mkdir -p gpurun_out/final
O=gpurun_out/final
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "autotune or backward" > $O/s20_pytest.log 2>&1; tail -2 $O/s20_pytest.log
python bench.py --workload resnet50 --train --steps 5 --warmup 3 --no-cpu > $O/bench_resnet50_train.json 2> $O/bench_resnet50_train.err; tail -2 $O/bench_resnet50_train.err
python bench.py --workload alexnet --train --steps 5 --warmup 3 --no-cpu > $O/bench_alexnet_train.json 2> $O/bench_alexnet_train.err; tail -2 $O/bench_alexnet_train.err
python - <<'PY'
import json
for f in ["bench_resnet50_train","bench_alexnet_train"]:
    d=json.loads(open("gpurun_out/final/"+f+".json").read().strip().splitlines()[-1])
    print(f, "value %.0f ms %.2f"%(d["value"], d["ms_per_step"]))
    seen=set()
    for L in d["layers"]:
        if L["op"]=="bwd_weight" and L["kernel"] not in seen or L["layer"].endswith("a_branch2b") and L["op"]=="bwd_weight" or "alexnet" in L["layer"] and L["op"]=="bwd_weight":
            seen.add(L["kernel"]); print("   %-26s %.3f ms %5.2f TF %s"%(L["layer"],L["ms"],L["tflops"],L["kernel"]))
PY
