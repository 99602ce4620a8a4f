#!/usr/bin/env python
"""Event trace of the TMEM-window kernel (library built with `make -C caffe_escoin_b200/csrc TMTRACE=1`): runs one layer
once and prints, for CTA 0, how each warp's time splits between waiting and working.
python tools/tm_trace.py <net> <idx> <variant> [dump]

Event codes (tmem_kernel.cuh): compute warps 1/2 = before/after the chunk's smem_full wait, 3 = slot group's tm_full wait
done, 4 = its taps done, 5/6 = epilogue begin/end; producers 11/12 = before/after smem_full, 13 = tm_empty wait done,
14 = slot group filled; loader 7 = issue begin, 8 = stage free (smem_empty wait done), 9 = copies issued."""
import ctypes as C
import os
import sys
from collections import defaultdict

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from caffe_escoin_b200 import capi, workloads as wl  # noqa: E402

NW, NE = 24, 4096


def fetch():
    buf = np.zeros(NW * NE, dtype=np.uint64)
    cnt = np.zeros(NW, dtype=np.int32)
    rc = capi.lib.escort_tmem_trace(buf.ctypes.data_as(C.c_void_p), cnt.ctypes.data_as(C.c_void_p))
    assert rc == 0
    return buf.reshape(NW, NE), cnt


def main():
    net, idx, variant = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    dump = len(sys.argv) > 4
    spec = wl.NETWORKS[net][idx]
    d = wl.make_layer_data(spec, idx)
    geom = capi.make_geom(spec.Cin, spec.Cout, spec.H, spec.H, spec.k, spec.stride, spec.pad, 1, spec.group)
    csr = capi.weight_align(torch.from_numpy(d["w"]).cuda(), geom)
    x = torch.from_numpy(d["x"]).cuda()
    b = torch.from_numpy(d["bias"]).cuda() if d["bias"] is not None else None
    plan = capi.Plan(geom, csr)
    plan.set_variant(variant)
    print(plan.describe())
    y = torch.empty((spec.N, spec.Cout, plan.Ho, plan.Wo), device="cuda")
    plan.forward(x, b, top=y)
    torch.cuda.synchronize()
    fetch()  # warm-up launch discarded
    plan.forward(x, b, top=y)
    torch.cuda.synchronize()
    ev, cnt = fetch()
    t0 = min(int(ev[w, 0] >> 8) for w in range(NW) if cnt[w])
    tend = max(int(ev[w, cnt[w] - 1] >> 8) for w in range(NW) if cnt[w])
    print("CTA 0: %d clocks traced, events per warp: %s" % (tend - t0, list(cnt)))
    # interval accounting: time between consecutive events of a warp is attributed to the LATER event's code
    names = {1: "other(compute)", 2: "wait smem_full", 3: "wait tm_full", 4: "taps", 5: "other", 6: "epilogue",
             11: "other(producer)", 12: "wait smem_full", 13: "wait tm_empty", 14: "fill", 7: "other/between", 8: "wait smem_empty(+table)",
             9: "issue copies"}
    for w in range(NW):
        n = int(cnt[w])
        if n < 2:
            continue
        acc = defaultdict(int)
        num = defaultdict(int)
        for i in range(1, n):
            code = int(ev[w, i] & 0xff)
            acc[code] += int(ev[w, i] >> 8) - int(ev[w, i - 1] >> 8)
            num[code] += 1
        tot = sum(acc.values())
        parts = ", ".join("%s %.1f%% (%d x %.0f)" % (names.get(c, str(c)), 100.0 * acc[c] / tot, num[c], acc[c] / max(num[c], 1))
                          for c in sorted(acc, key=lambda c: -acc[c]))
        print("warp %2d  span %8d  %s%s" % (w, tot, parts, "  [LOG FULL]" if n >= NE else ""))
    if dump:
        for w in (0, 16):
            print("warp", w)
            for i in range(min(int(cnt[w]), 400)):
                print("  %8d  %d" % (int(ev[w, i] >> 8) - t0, int(ev[w, i] & 0xff)))


if __name__ == "__main__":
    main()
