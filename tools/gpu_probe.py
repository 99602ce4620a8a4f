#!/usr/bin/env python
"""One-off GPU probe: FP32 FMA peak (three operand patterns), and per-layer timing of our forward vs the
reference's own GPU kernels (oracle/_ref).  Writes gpurun_out/probe.json."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

ge.build()
from caffe_escoin_b200 import capi, workloads as wl  # noqa: E402
from oracle import pyoracle as po  # noqa: E402


def time_cuda(fn, warmup=3, iters=10):
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
    out = {"device": torch.cuda.get_device_name(0)}
    peaks = {}
    for v, name in ((0, "ffma_shared_operand"), (1, "ffma2_packed"), (2, "ffma_3reg")):
        tf, sms, khz = capi.measure_fp32_peak(v, 8192)
        peaks[name] = tf
        out["sm_count"], out["clock_khz"] = sms, khz
    out["fp32_peak_tflops"] = peaks
    print(json.dumps(out), flush=True)
    variants = [int(v) for v in os.environ.get("ESCORT_VARIANTS", "-1,0").split(",")]
    layers = []
    specs = wl.ALEXNET + [wl.GOOGLENET[0], wl.GOOGLENET[1], wl.GOOGLENET[13], wl.GOOGLENET[17]] + \
        [wl.RESNET50[0], wl.RESNET50[3], wl.RESNET50[7], wl.RESNET50[13]]
    R = C.CDLL(po.ref_gpu_path()) if po.have_ref_gpu() else None
    for li, spec in enumerate(specs):
        d = wl.make_layer_data(spec, li)
        geom = capi.make_geom(spec.Cin, spec.Cout, spec.H, spec.H, spec.k, spec.stride, spec.pad, 1, spec.group)
        csr = capi.weight_align(torch.from_numpy(d["w"]).cuda(), geom)
        x = torch.from_numpy(d["x"]).cuda()
        b = torch.from_numpy(d["bias"]).cuda() if d["bias"] is not None else None
        plan = capi.Plan(geom, csr)
        nnz = plan.nnz
        flops, byts = wl.alg_work(spec, nnz)
        rec = {"layer": spec.name, "N": spec.N, "nnz": nnz, "gflop": flops / 1e9, "mb": byts / 1e6}
        y = torch.empty((spec.N, spec.Cout, plan.Ho, plan.Wo), device="cuda")
        yref = None
        for v in variants:
            try:
                if v != -1:
                    plan.set_variant(v)
            except capi.EscortError as e:
                rec["v%d" % v] = "unsupported"
                continue
            ms = time_cuda(lambda: plan.forward(x, b, top=y))
            if yref is None:
                yref = y.clone()
                err = 0.0
            else:
                err = float((y - yref).norm() / yref.norm())
            rec["v%d" % v] = {"kernel": plan.kernel_name, "ms": ms, "tflops": flops / ms / 1e9, "img_s": spec.N / ms * 1e3,
                              "rel_vs_first": err}
        if R is not None:
            Hp, Wp = spec.H + spec.pad, spec.H + spec.pad
            plen = spec.Cin * Hp * Wp + spec.pad * (spec.H + 2 * spec.pad)
            padded = torch.zeros(plen, device="cuda")
            top = torch.zeros_like(y)
            p = lambda t: C.c_void_p(0 if t is None else t.data_ptr())
            nref = min(spec.N, 16)   # the reference syncs the device after every launch; time a slice of the batch

            def ref_run():
                R.refgpu_conv_forward(p(x), nref, spec.Cin, spec.H, spec.H, spec.Cout, spec.group, spec.k, spec.k,
                                      spec.pad, spec.pad, spec.stride, spec.stride, 1, 1, p(csr["values"]),
                                      p(csr["colidx"]), p(csr["rowptr"]), p(b), 0, p(top), p(padded))
            ms = time_cuda(ref_run, warmup=1, iters=3)
            rec["reference_gpu"] = {"ms_per_image": ms / nref, "img_s": nref / ms * 1e3,
                                    "rel_l2_vs_ours": float((top[:nref] - yref[:nref]).norm() / yref[:nref].norm())}
        layers.append(rec)
        print(json.dumps(rec), flush=True)
    out["layers"] = layers
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "probe.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
