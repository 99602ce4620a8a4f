#!/usr/bin/env python
"""f1 + f2 together: a ResNet-50 bottleneck block (branch2a 1x1 -> BN -> Scale -> ReLU -> pruned branch2b 3x3 -> BN -> Scale -> ReLU
-> branch2c 1x1 -> BN -> Scale -> Eltwise SUM -> ReLU; models/resnet/test_sconv.prototxt) as THREE launches -- dense tcgen05,
sparse direct, dense tcgen05 with the residual in its epilogue -- at batch 256, next to the same 13 layers through
PyTorch (cuDNN fp32 / TF32 convolutions + eager BN / add / ReLU = one HBM pass per layer, as Caffe runs them).
python tools/run_block.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from caffe_escoin_b200 import capi, workloads as wl  # noqa: E402

N, EPS = 256, 1e-5
F = torch.nn.functional


def best_ms(fn, n=5):
    fn()
    best = 1e9
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


print("bottleneck block forward, batch %d (ms; ours = 3 launches)" % N)
for name, Cio, Cmid, H in [("res2b", 256, 64, 56), ("res3b", 512, 128, 28), ("res4b", 1024, 256, 14), ("res5b", 2048, 512, 7)]:
    g = torch.Generator(device="cuda").manual_seed(H)
    x = torch.rand((N, Cio, H, H), device="cuda", generator=g) * 2 - 1
    w_a = torch.randn((Cmid, Cio, 1, 1), device="cuda", generator=g) / Cio ** 0.5
    rng = np.random.default_rng(H)
    w_b = torch.from_numpy(wl.prune_magnitude((rng.standard_normal((Cmid, Cmid, 3, 3)) / (Cmid * 9) ** 0.5).astype(np.float32), 0.7)).cuda()
    w_c = torch.randn((Cio, Cmid, 1, 1), device="cuda", generator=g) / Cmid ** 0.5
    bns = []
    for M in (Cmid, Cmid, Cio):
        bns.append((torch.randn(M, device="cuda", generator=g) * 0.3, torch.rand(M, device="cuda", generator=g) + 0.5, 1.0,
                    torch.rand(M, device="cuda", generator=g) + 0.5, torch.randn(M, device="cuda", generator=g) * 0.2))

    def bn(t, p):
        return F.batch_norm(t, p[0], p[1], p[3], p[4], False, 0.0, EPS)

    def torch_block():
        r = torch.relu(bn(F.conv2d(x, w_a), bns[0]))
        r = torch.relu(bn(F.conv2d(r, w_b, padding=1), bns[1]))
        return torch.relu(bn(F.conv2d(r, w_c), bns[2]) + x)

    affine = [capi.bn_scale_to_affine(p[0], p[1], p[2], EPS, p[3], p[4]) for p in bns]
    g_a = capi.make_geom(Cio, Cmid, H, H, 1, 1, 0, 1, 1)
    g_b = capi.make_geom(Cmid, Cmid, H, H, 3, 1, 1, 1, 1)
    g_c = capi.make_geom(Cmid, Cio, H, H, 1, 1, 0, 1, 1)
    wa_f, ba_f = capi.dense_fold_affine(w_a, *affine[0])
    wc_f, bc_f = capi.dense_fold_affine(w_c, *affine[2])
    plan = capi.Plan(g_b, capi.weight_align(w_b, g_b))
    plan.autotune(N)
    bb_f = capi.fold_affine(plan, w_b, *affine[1])
    ws = torch.empty(max(capi.dense_conv_workspace_bytes(g_a, N), capi.dense_conv_workspace_bytes(g_c, N)), dtype=torch.uint8, device="cuda")
    t1 = torch.empty((N, Cmid, H, H), device="cuda")
    t2 = torch.empty_like(t1)
    y = torch.empty_like(x)

    def ours():
        capi.dense_conv_forward(g_a, x, wa_f, ba_f, relu=True, top=t1, workspace=ws)
        plan.forward(t1, bb_f, relu=True, top=t2)
        capi.dense_conv_forward(g_c, t2, wc_f, bc_f, relu=True, residual=x, top=y, workspace=ws)

    t_ours = best_ms(ours)
    t_a = best_ms(lambda: capi.dense_conv_forward(g_a, x, wa_f, ba_f, relu=True, top=t1, workspace=ws))
    t_b = best_ms(lambda: plan.forward(t1, bb_f, relu=True, top=t2))
    t_c = best_ms(lambda: capi.dense_conv_forward(g_c, t2, wc_f, bc_f, relu=True, residual=x, top=y, workspace=ws))
    torch.backends.cudnn.allow_tf32 = False
    ref = torch_block()
    t_fp32 = best_ms(torch_block)
    torch.backends.cudnn.allow_tf32 = True
    t_tf32 = best_ms(torch_block)
    ours()
    torch.cuda.synchronize()
    err = float((y - ref).norm() / ref.norm())
    print("  %-6s %4d -> %3d -> %4d @%2d | ours %6.3f (2a %.3f + 2b %.3f [%s] + 2c %.3f) rel_l2 %.1e | torch cuDNN fp32 %6.3f (%.1fx) | cuDNN tf32 %6.3f (%.1fx)"
          % (name, Cio, Cmid, Cio, H, t_ours, t_a, t_b, plan.kernel_name, t_c, err, t_fp32, t_fp32 / t_ours, t_tf32, t_tf32 / t_ours), flush=True)
