#!/usr/bin/env python
"""One interpreter-only microbenchmark launch set (for ncu): python tools/interp_one.py VARIANT PER_LOAD [ACTIVE_WARPS]"""
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
v, pl = int(sys.argv[1]), int(sys.argv[2])
aw = int(sys.argv[3]) if len(sys.argv) > 3 else 0
ms, tf, nc = C.c_double(0), C.c_double(0), C.c_int(0)
rc = lib.escort_interp_bench(v, pl, aw, 20, C.byref(ms), C.byref(tf), C.byref(nc))
print("v%d pl%d aw%d rc=%d %.3f ms %.2f TF NC=%d" % (v, pl, aw, rc, ms.value, tf.value, nc.value))
