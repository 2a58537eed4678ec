"""Multi-GPU inference plumbing: one process per GPU, weights broadcast once, batch sharded with no per-step
collective.  Replaces torch.nn.DataParallel in the reference (src/kbnet_model.py:408-415), which
re-broadcasts all weights from GPU 0 on every forward.

Every op on the path is per-sample (no BatchNorm), so rank r of R simply runs samples [r*N/R, (r+1)*N/R).
The backend is whatever the process group was created with: NCCL over NVLink/NVSwitch on the B200 box, gloo in the
CPU tests of this host-side logic.
"""
import torch
import torch.distributed as dist


def shard_range(n, rank, world_size):
    """Contiguous slice [lo, hi) of a batch of n samples owned by `rank` (first n % R ranks get one extra)."""
    base, extra = divmod(n, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(tensors, rank, world_size):
    n = tensors[0].shape[0]
    lo, hi = shard_range(n, rank, world_size)
    return [t[lo:hi] for t in tensors]


def broadcast_parameters(params, src=0, group=None):
    """One flattened broadcast of all weights (27.8 MB fp32 for KBNet) from `src`; called once at start-up."""
    params = [p for p in params]
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return 0
    with torch.no_grad():
        flat = torch.cat([p.detach().reshape(-1) for p in params])
        dist.broadcast(flat, src=src, group=group)
        off = 0
        for p in params:
            k = p.numel()
            p.copy_(flat[off:off + k].view_as(p))
            off += k
    return flat.numel() * flat.element_size()


def gather_outputs(local_out, n_total, group=None):
    """Optional all-gather of per-rank outputs [n_r, ...] back into [n_total, ...] (verification only)."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local_out
    world = dist.get_world_size(group)
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    max_n = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((max_n,) + tuple(local_out.shape[1:]), dtype=local_out.dtype, device=local_out.device)
    pad[:local_out.shape[0]] = local_out
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([b[:hi - lo] for b, (lo, hi) in zip(bufs, sizes)], dim=0)
