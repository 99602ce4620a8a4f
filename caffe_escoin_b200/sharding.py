"""Host-side logic of the path's one multi-GPU dimension: the batch is sharded by image across ranks (one process
per GPU), inference needs no collective, and a training step has exactly one exchange -- all-reduce(sum) of the flat
gradient buffer followed by a 1/world scale, the reference's NCCL<Dtype>::on_gradients_ready
(src/caffe/parallel.cpp:238-256) over GPUParams' contiguous diff buffer (:75-108).  Sparse layers contribute their
gradient in CSR order (nnz floats, identical mask on every replica) instead of the dense weight count.

torch.distributed is the plumbing (NCCL on GPUs, gloo in the CPU tests); nothing here computes a convolution."""
from collections import namedtuple

import torch
import torch.distributed as dist

Segment = namedtuple("Segment", "name kind offset count")


def shard_range(total, world, rank):
    """Contiguous image shard [start, start+count) of rank; the first total % world ranks get one extra image."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError("bad world/rank")
    base, rem = divmod(total, world)
    start = rank * base + min(rank, rem)
    return start, base + (1 if rank < rem else 0)


def flat_layout(layers):
    """layers: iterable of (name, nnz, num_bias).  Returns (segments, total) for the flat diff buffer:
    per layer the CSR-ordered weight gradient (nnz floats) then the bias gradient, 4-float aligned segments."""
    segs, off = [], 0
    for name, nnz, nbias in layers:
        segs.append(Segment(name, "weight_csr", off, int(nnz)))
        off += (int(nnz) + 3) // 4 * 4
        if nbias:
            segs.append(Segment(name, "bias", off, int(nbias)))
            off += (int(nbias) + 3) // 4 * 4
    return segs, off


def exchange_gradients(flat, world=None, group=None):
    """on_gradients_ready: sum over ranks, then scale by 1/solver_count. In place; returns flat."""
    world = dist.get_world_size(group) if world is None else world
    if world > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if world != 1:
        if flat.is_cuda:
            from . import capi  # scale on the device through the C ABI (comm = NULL: no second reduction)
            capi.allreduce_grads(flat, 1.0 / world)
        else:
            flat.mul_(1.0 / world)
    return flat


def broadcast_weights(flat, src=0, group=None):
    """NCCL::Broadcast of the flat data buffer from the root solver (parallel.cpp:189-199)."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.broadcast(flat, src=src, group=group)
    return flat
