"""World-size-2 gloo tests (CPU) of the N > 1 host logic: image sharding, the flat CSR-ordered gradient buffer and the
all-reduce + 1/N exchange (reference src/caffe/parallel.cpp:238-256).  The per-rank gradients come from the oracle
(test infrastructure), so the check is: exchange(shard gradients) == full-batch gradient / world."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _layer_grads(po, wl, spec, x, dy, w):
    g = po.Geom(x.shape[0], spec.Cin, spec.H, spec.H, spec.Cout, spec.k, spec.stride, spec.pad, 1, spec.group)
    wd, bd, _ = po.conv_backward(x, dy, w, g, mask_only=True, want_x=False)
    csr = po.weight_align(w, g, stretch=False)
    # CSR order = row-major order of the nonzero positions (per group, groups concatenated)
    return wd.reshape(-1)[np.flatnonzero(w.reshape(-1))], bd, int(sum(csr["nz_num"]))


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from caffe_escoin_b200 import sharding, workloads as wl
    from oracle import pyoracle as po
    specs = [wl.ALEXNET[1]._replace(N=6, Cin=8, Cout=12), wl.ALEXNET[0]._replace(N=6, Cin=8, Cout=8, H=9)]
    rng = np.random.default_rng(7)
    full, mine, layout = [], [], []
    for li, spec in enumerate(specs):
        d = wl.make_layer_data(spec, li)
        Ho = wl.out_dim(spec.H, spec.pad, spec.k, spec.stride)
        dy = rng.standard_normal((spec.N, spec.Cout, Ho, Ho)).astype(np.float32)
        s, c = sharding.shard_range(spec.N, world, rank)
        fw, fb, nnz = _layer_grads(po, wl, spec, d["x"], dy, d["w"])
        mw, mb, _ = _layer_grads(po, wl, spec, d["x"][s:s + c], dy[s:s + c], d["w"])
        full.append((fw, fb))
        mine.append((mw, mb))
        layout.append((spec.name, nnz, spec.Cout))
    segs, total = sharding.flat_layout(layout)
    flat = torch.zeros(total)
    want = torch.zeros(total)
    it = iter(segs)
    for (mw, mb), (fw, fb) in zip(mine, full):
        sw, sb = next(it), next(it)
        assert sw.count == mw.size and sb.count == mb.size
        flat[sw.offset:sw.offset + sw.count] = torch.from_numpy(mw)
        flat[sb.offset:sb.offset + sb.count] = torch.from_numpy(mb)
        want[sw.offset:sw.offset + sw.count] = torch.from_numpy(fw) / world
        want[sb.offset:sb.offset + sb.count] = torch.from_numpy(fb) / world
    sharding.exchange_gradients(flat, world)
    err = float((flat - want).norm() / want.norm())
    # weights broadcast from rank 0
    wflat = torch.full((16,), float(rank + 1))
    sharding.broadcast_weights(wflat, src=0)
    ok_b = bool((wflat == 1.0).all())
    # the bench's timing reduction: max over ranks
    t = torch.tensor([1.0 + rank])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    out[rank] = (err, ok_b, float(t.item()))
    dist.destroy_process_group()


def test_shard_range_covers_batch():
    from caffe_escoin_b200 import sharding
    for total in (0, 1, 7, 256, 257):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_range(total, world, r) for r in range(world)]
            assert sum(c for _, c in spans) == total
            pos = 0
            for s, c in spans:
                assert s == pos
                pos += c
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
    with pytest.raises(ValueError):
        sharding.shard_range(4, 2, 2)


def test_flat_layout_alignment():
    from caffe_escoin_b200 import sharding
    segs, total = sharding.flat_layout([("a", 5, 3), ("b", 8, 0), ("c", 1, 1)])
    assert [s.kind for s in segs] == ["weight_csr", "bias", "weight_csr", "weight_csr", "bias"]
    assert all(s.offset % 4 == 0 for s in segs) and total % 4 == 0
    for a, b in zip(segs, segs[1:]):
        assert a.offset + a.count <= b.offset


def test_gradient_exchange_world2_gloo():
    from oracle import pyoracle as po
    po.build()
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert len(out) == world
    for rank in range(world):
        err, ok_b, tmax = out[rank]
        assert err < 1e-5, "rank %d: exchanged gradient differs from full-batch / world: %g" % (rank, err)
        assert ok_b, "weights were not broadcast from rank 0"
        assert tmax == 2.0
