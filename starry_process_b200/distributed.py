"""Multi-GPU plumbing: one process per GPU (torch.distributed, NCCL over NVLink on the GPU box,
gloo in the CPU tests).  The hyperparameter-sample x light-curve x inclination batch is a set of
independent likelihood evaluations, so it is split statically across ranks with NO data-path
collective; the only exchange is one all-gather of the per-element log-likelihoods
(<= 8 bytes per element)."""
import torch
import torch.distributed as dist


def shard_range(n_items, rank=None, world_size=None):
    """Contiguous, balanced [begin, end) slice of ``n_items`` for this rank."""
    if rank is None:
        rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
    if world_size is None:
        world_size = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    base, rem = divmod(int(n_items), int(world_size))
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def gather_lnlike(local, n_items=None, equal_shards=False):
    """All-gather the per-element log-likelihoods of every rank's shard into the full vector
    (ordered by rank).  ``local`` is this rank's 1-D tensor; shards may differ in length by one.
    ``equal_shards=True`` (every rank holds the same count, e.g. weak scaling) is a single
    collective with no host synchronisation."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    if equal_shards:
        out = torch.empty(world * local.numel(), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous())
        if n_items is not None:
            assert out.numel() == n_items
        return out
    n_local = torch.tensor([local.numel()], dtype=torch.int64, device=local.device)
    sizes = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(sizes, n_local)
    sizes = [int(s.item()) for s in sizes]
    nmax = max(sizes)
    pad = torch.zeros(nmax, dtype=local.dtype, device=local.device)
    pad[: local.numel()] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad)
    out = torch.cat([p[:s] for p, s in zip(parts, sizes)])
    if n_items is not None:
        assert out.numel() == n_items
    return out
