#!/usr/bin/env python
"""Interpreter-only ceiling of every tile variant (no loader, no barriers): python tools/interp_bench.py [per_load ...]"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: F401,E402
from caffe_escoin_b200 import capi  # noqa: E402

torch.zeros(1).cuda()
lib = capi.lib
lib.escort_interp_bench.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                    C.POINTER(C.c_int)]
# python tools/interp_bench.py [--from V] [per_load ...]; for the sieve variants per_load = density in percent
args = sys.argv[1:]
v = 1
if args and args[0] == "--from":
    v = int(args[1])
    args = args[2:]
per_loads = [int(a) for a in args] or [0, 16, 6]
while True:
    row = []
    for pl in per_loads:
        for aw in (0,):
            ms, tf, nc = C.c_double(0), C.c_double(0), C.c_int(0)
            rc = lib.escort_interp_bench(v, pl, aw, 20, C.byref(ms), C.byref(tf), C.byref(nc))
            if rc != 0:
                break
            row.append("pl%-2d %6.2f TF" % (pl, tf.value))
        if rc != 0:
            break
    if rc != 0:
        break
    p = capi  # name via a throwaway query is not exposed; print the index
    print("v%-2d NC=%-3d %s" % (v, nc.value, " | ".join(row)), flush=True)
    v += 1
