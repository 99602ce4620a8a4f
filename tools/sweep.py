#!/usr/bin/env python
"""BASELINE.json configs[4]: single-layer sweep, sparsity 50-95 % x C=M 64-512 x H=W 7-56 x stride 1/2 (3x3, pad 1, N=64):
our direct sparse conv (default plan, no autotune) against the reference's own GPU direct sconv (oracle/_ref, a slice of
the batch: it syncs the device after every launch), with the per-point roofline fraction.
    python tools/sweep.py [--autotune] [--stride 1|2] > gpurun_out/sweep.jsonl"""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from caffe_escoin_b200 import capi, workloads as wl  # noqa: E402
from oracle import pyoracle as po  # noqa: E402  (checker + comparator only)


def time_cuda(fn, warmup=2, iters=5):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def main():
    autotune = "--autotune" in sys.argv
    peak = max(capi.measure_fp32_peak(v, 8192)[0] for v in (0, 1, 2))
    hbm = 6546.2
    try:
        hbm = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    R = C.CDLL(po.ref_gpu_path()) if po.have_ref_gpu() else None
    print(json.dumps({"fp32_peak_tflops": peak, "hbm_gbs": hbm, "reference_gpu": R is not None}), flush=True)
    only_stride = int(sys.argv[sys.argv.index("--stride") + 1]) if "--stride" in sys.argv else None
    for li, spec in enumerate(wl.sweep_specs(64)):
        if only_stride is not None and spec.stride != only_stride:
            continue
        d = wl.make_layer_data(spec, li)
        geom = capi.make_geom(spec.Cin, spec.Cout, spec.H, spec.H, spec.k, spec.stride, spec.pad, 1, spec.group)
        csr = capi.weight_align(torch.from_numpy(d["w"]).cuda(), geom)
        x = torch.from_numpy(d["x"]).cuda()
        b = torch.from_numpy(d["bias"]).cuda()
        plan = capi.Plan(geom, csr)
        if autotune:
            plan.autotune(spec.N)
        y = torch.empty((spec.N, spec.Cout, plan.Ho, plan.Wo), device="cuda")
        flops, byts = wl.alg_work(spec, plan.nnz)
        ms = time_cuda(lambda: plan.forward(x, b, top=y))
        t_roof = max(flops / (peak * 1e12), byts / (hbm * 1e9))
        rec = {"point": spec.name, "sparsity": spec.sparsity, "C": spec.Cin, "H": spec.H, "stride": spec.stride,
               "nnz": int(plan.nnz), "kernel": plan.kernel_name, "s2d": "(s2d)" in plan.describe(), "ms": ms, "tflops": flops / ms / 1e9,
               "img_s": spec.N / ms * 1e3, "roofline_frac": t_roof / (ms * 1e-3),
               "bound": "fp32_fma" if flops / (peak * 1e12) >= byts / (hbm * 1e9) else "hbm"}
        if R is not None:
            Hp = spec.H + spec.pad
            plen = spec.Cin * Hp * Hp + spec.pad * (spec.H + 2 * spec.pad)
            padded = torch.zeros(plen, device="cuda")
            top = torch.zeros_like(y)
            p = lambda t: C.c_void_p(0 if t is None else t.data_ptr())
            nref = 4

            def ref_run():
                R.refgpu_conv_forward(p(x), nref, spec.Cin, spec.H, spec.H, spec.Cout, spec.group, spec.k, spec.k,
                                      spec.pad, spec.pad, spec.stride, spec.stride, 1, 1, p(csr["values"]),
                                      p(csr["colidx"]), p(csr["rowptr"]), p(b), 0, p(top), p(padded))
            rms = time_cuda(ref_run, warmup=1, iters=2)
            rec["reference_gpu_img_s"] = nref / rms * 1e3
            rec["rel_l2_vs_reference_gpu"] = float((top[:nref] - y[:nref]).norm() / y[:nref].norm())
            rec["speedup_vs_reference_gpu"] = rec["img_s"] / rec["reference_gpu_img_s"]
        print(json.dumps(rec), flush=True)
        del plan


if __name__ == "__main__":
    main()
