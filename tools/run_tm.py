#!/usr/bin/env python
"""TMEM-window kernel bring-up: parity of every applicable TMEM variant against the CPU oracle on small cases, then timing
at full BASELINE size against the previous default.  python tools/run_tm.py [check|time <net> [idx ...]]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from caffe_escoin_b200 import capi, workloads as wl  # noqa: E402
from oracle import pyoracle as po  # noqa: E402


def tm_variants(plan):
    out = []
    for v in range(1, 100):
        try:
            plan.set_config(v, 0)
        except capi.EscortError:
            continue
        if plan.kernel_name.startswith("sconv_tmem"):
            out.append((v, plan.kernel_name))
    return out


def check():
    C = wl.ConvSpec
    cases = [
        C("thin_conv3", 5, 32, 48, 13, 3, 1, 1, 1, 0.88, True, True),
        C("thin_conv2_g2", 3, 16, 32, 27, 5, 1, 2, 2, 0.85, True, True),
        C("lenet2", 4, 20, 50, 12, 5, 1, 0, 1, 0.80, True, False),
        C("res_56", 2, 16, 16, 56, 3, 1, 1, 1, 0.70, False, False),
        C("res_7", 9, 64, 40, 7, 3, 1, 1, 1, 0.70, False, True),
        C("pointwise", 3, 40, 24, 14, 1, 1, 0, 1, 0.6, True, False),
        C("dense_k3", 2, 8, 70, 10, 3, 1, 1, 1, 0.0, True, False),
        C("many_ch", 2, 300, 130, 6, 3, 1, 1, 1, 0.9, True, True),
    ]
    bad = 0
    for i, spec in enumerate(cases):
        d = wl.make_layer_data(spec, i)
        g = po.Geom(spec.N, spec.Cin, spec.H, spec.H, spec.Cout, spec.k, spec.stride, spec.pad, 1, spec.group)
        ocsr = po.weight_align(d["w"], g)
        y_ref = po.conv_forward(d["x"], ocsr, g, d["bias"], relu=spec.relu)
        geom = capi.make_geom(spec.Cin, spec.Cout, spec.H, spec.H, spec.k, spec.stride, spec.pad, 1, spec.group)
        csr = capi.weight_align(torch.from_numpy(d["w"]).cuda(), geom)
        x = torch.from_numpy(d["x"]).cuda()
        b = torch.from_numpy(d["bias"]).cuda() if d["bias"] is not None else None
        plan = capi.Plan(geom, csr)
        print(spec.name, "default:", plan.kernel_name, flush=True)
        for v, name in tm_variants(plan):
            plan.set_config(v, 0)
            y = torch.full(y_ref.shape, float("nan"), device="cuda")
            plan.forward(x, b, relu=spec.relu, top=y)
            torch.cuda.synchronize()
            err = po.rel_l2(y.cpu().numpy(), y_ref)
            ok = err < 1e-4
            bad += 0 if ok else 1
            print("  %-28s v%-3d rel_l2 %.2e %s   %s" % (name, v, err, "ok" if ok else "MISMATCH", plan.describe()[:150]), flush=True)
    print("check: %d mismatches" % bad)
    return bad


def time_net(net, idxs):
    specs = wl.NETWORKS[net]
    for idx in idxs or range(len(specs)):
        spec = specs[idx]
        d = wl.make_layer_data(spec, idx)
        geom = capi.make_geom(spec.Cin, spec.Cout, spec.H, spec.H, spec.k, spec.stride, spec.pad, 1, spec.group)
        csr = capi.weight_align(torch.from_numpy(d["w"]).cuda(), geom)
        x = torch.from_numpy(d["x"]).cuda()
        b = torch.from_numpy(d["bias"]).cuda() if d["bias"] is not None else None
        plan = capi.Plan(geom, csr)
        flops, _ = wl.alg_work(spec, plan.nnz)
        y0 = None
        os.environ["ESCORT_NO_TMEM"] = "1"
        plan.set_config(-1, 0)
        del os.environ["ESCORT_NO_TMEM"]
        cands = [(-2, plan.kernel_name)] + tm_variants(plan)
        for v, name in cands:
            if v == -2:
                os.environ["ESCORT_NO_TMEM"] = "1"
                plan.set_config(-1, 0)
                del os.environ["ESCORT_NO_TMEM"]
            else:
                plan.set_config(v, 0)
            y = torch.full((spec.N, spec.Cout, plan.Ho, plan.Wo), float("nan"), device="cuda")
            best = 1e9
            for _ in range(4):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                plan.forward(x, b, relu=spec.relu, top=y)
                e1.record()
                e1.synchronize()
                best = min(best, e0.elapsed_time(e1))
            if y0 is None:
                y0 = y
                err = 0.0
            else:
                err = float(torch.linalg.vector_norm((y - y0).double()) / torch.linalg.vector_norm(y0.double()))
            print("RESULT %-28s %-44s %8.3f ms %6.2f TFLOP/s  rel_l2 vs first %.1e   %s" % (spec.name, name, best, flops / best / 1e9, err, plan.describe()[len(name):110]), flush=True)
        del plan


def dump_debug():
    import ctypes as C
    buf = (C.c_int * 16)()
    capi.lib.escort_tmem_debug(buf)
    print("tmem debug words {code, block, warp, barrier offset, parity}:", list(buf)[:6], flush=True)


if __name__ == "__main__":
    torch.zeros(1).cuda()
    try:
        if len(sys.argv) < 2 or sys.argv[1] == "check":
            rc = 1 if check() else 0
        else:
            time_net(sys.argv[2], [int(a) for a in sys.argv[3:]])
            rc = 0
    except Exception as e:  # noqa: BLE001
        print("FAILED:", repr(e)[:300], flush=True)
        dump_debug()
        rc = 2
    sys.exit(rc)
