#!/usr/bin/env python
"""Magnitude-prune the convolution (and optionally inner-product) weights of a `.caffemodel` through the C ABI
(escort_caffemodel_* / escort_prune_magnitude) and write the pruned model: the input WeightAlign expects from the
SkimCaffe checkpoints the reference's run.sh:13 names.  Biases and every other field are written back unchanged.

python tools/prune_caffemodel.py in.caffemodel out.caffemodel --sparsity 0.88 [--layer conv2=0.85 ...] [--skip conv1] [--fc]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from caffe_escoin_b200 import capi  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("src")
    ap.add_argument("dst")
    ap.add_argument("--sparsity", type=float, default=0.88, help="default target sparsity of a pruned layer")
    ap.add_argument("--layer", action="append", default=[], help="name=sparsity override (repeatable)")
    ap.add_argument("--skip", action="append", default=["conv1"], help="layers left dense (the reference keeps conv1 dense)")
    ap.add_argument("--fc", action="store_true", help="also prune InnerProduct layers")
    args = ap.parse_args()
    over = dict((kv.split("=")[0], float(kv.split("=")[1])) for kv in args.layer)
    m = capi.CaffeModel(args.src)
    for i in range(len(m)):
        L = m.layer(i)
        if L["num_blobs"] == 0 or not (L["is_conv"] or (args.fc and L["is_inner_product"])):
            continue
        w = m.blob(i, 0)
        if L["name"] in args.skip and L["name"] not in over:
            print("%-28s %-14s %-18s kept dense (%d weights)" % (L["name"], L["type"], w.shape, w.size))
            continue
        s = over.get(L["name"], args.sparsity)
        thr, nnz = capi.prune_magnitude(w.reshape(-1), s)
        print("%-28s %-14s %-18s sparsity %.3f -> nnz %d of %d (threshold %.4g)" % (L["name"], L["type"], w.shape, 1 - nnz / max(w.size, 1),
                                                                                  nnz, w.size, thr))
    m.save(args.dst)
    print("wrote", args.dst)


if __name__ == "__main__":
    main()
