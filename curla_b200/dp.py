"""Host-side data-parallel plumbing for CurlSacAgent.update (one process per GPU).

The reference is single-GPU; this is the B200 scale-out of the same update:
every rank draws the IDENTICAL global index block from the numpy global stream (same seed on
every rank, utils.py:147 / augmentations.py:66-67 order), takes the contiguous slice
[rank*B, (rank+1)*B) of it, and the engine (csrc/engine.cu) all-reduces the three gradient
buckets and all-gathers the CURL keys so that each rank's (B x B_global) logits block sees
the full-batch negatives with labels rank*B + arange(B) (curl_sac.py:411-413).

Only pure host logic lives here so that it can be exercised on CPU with the gloo backend
(tests/test_dp_gloo.py); the collectives of the product path are NCCL calls inside the
engine.
"""
import ctypes as C

import torch


def world_info():
    """(rank, world) of the default process group, (0, 1) when not initialised."""
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        return torch.distributed.get_rank(), torch.distributed.get_world_size()
    return 0, 1


def local_batch(global_batch, world):
    """Per-rank batch; every loss is a mean over the batch (F.mse_loss, .mean(),
    CrossEntropyLoss: curl_sac.py:359,379,399,413), so equal shards make the mean of means
    equal to the global mean."""
    if global_batch % world != 0:
        raise ValueError('global batch %d must divide evenly over %d ranks' % (global_batch, world))
    return global_batch // world


def shard_slice(rank, world, global_batch):
    b = local_batch(global_batch, world)
    return slice(rank * b, (rank + 1) * b)


def label_offset(rank, world, global_batch):
    """Column of this rank's first positive key in the all-gathered key matrix."""
    return rank * local_batch(global_batch, world)


def grad_scale(global_batch):
    """Every rank scales its local gradient sums by 1/B_global so that a SUM all-reduce yields
    the gradient of the global-batch mean loss."""
    return 1.0 / float(global_batch)


def broadcast_bytes(payload, nbytes, device=None, src=0):
    """Broadcast a small byte string (the NCCL unique id) from `src` over the default group;
    works for gloo (CPU tensors) and nccl (tensors on `device`)."""
    import torch.distributed as dist
    t = torch.zeros(nbytes, dtype=torch.uint8)
    if dist.get_rank() == src:
        t = torch.frombuffer(bytearray(payload), dtype=torch.uint8).clone()
    if dist.get_backend() == 'nccl':
        t = t.to(device)
    dist.broadcast(t, src)
    return bytes(t.cpu().numpy().tobytes())


def nccl_unique_id(lib):
    buf = (C.c_char * 128)()
    rc = lib.curla_nccl_unique_id(C.cast(buf, C.c_void_p))
    if rc != 0:
        raise RuntimeError('curla_nccl_unique_id: %s' % lib.curla_last_error().decode())
    return bytes(buf.raw)
