"""Multi-GPU plumbing: one process per GPU (torch.distributed, NCCL over NVLink on the GPU box,
gloo in the CPU tests).  SURVEY.md section 8(e): the hyperparameter-sample x light-curve x
inclination batch is a set of independent likelihood evaluations, so every split below is a static
contiguous partition with NO data-path collective; the only exchanges are

  * ``log_likelihood_sharded``      one all-gather of the per-sample log-likelihoods (8 B each) --
                                    configs[2] / [3]: ONE sweep of B hyperparameter samples split B/G
                                    per GPU (calibrate/inclination.py:63-74 batches of this shape);
  * ``ensemble_log_likelihood_sharded``  one all-reduce(sum) of a scalar -- configs[1]: the single
                                    factorisation is replicated (0.33 GF, cheaper than broadcasting an
                                    8 MB factor) and the right-hand-side light curves are split;
  * ``design_matrix_sharded``       none -- configs[4]: the inclination (or time) axis of the design
                                    matrix is split and every rank keeps its rows.

Nothing here touches a CUDA kernel directly: the per-rank work goes through ``StarryProcess``; the
``evaluate`` hooks exist so that the host logic can be exercised with gloo on CPU tensors."""
import torch
import torch.distributed as dist

__all__ = ["shard_range", "gather_lnlike", "log_likelihood_sharded",
           "ensemble_log_likelihood_sharded", "design_matrix_sharded"]

_HYPER = ("r", "mu", "sigma", "a", "b", "c", "n", "tau")


def _world(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def shard_range(n_items, rank=None, world_size=None):
    """Contiguous, balanced [begin, end) slice of ``n_items`` for this rank."""
    if rank is None:
        rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
    if world_size is None:
        world_size = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    base, rem = divmod(int(n_items), int(world_size))
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def gather_lnlike(local, n_items=None, equal_shards=False, group=None):
    """All-gather the per-element log-likelihoods of every rank's shard into the full vector
    (ordered by rank).  ``local`` is this rank's 1-D tensor.  ``equal_shards=True`` (every rank
    holds the same count) is a single collective.  When ``n_items`` is given the shard sizes follow
    from ``shard_range`` (they differ by at most one), so ragged shards are still ONE collective of
    the padded vectors with no size exchange and no host synchronisation."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    if equal_shards:
        out = torch.empty(world * local.numel(), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        if n_items is not None:
            assert out.numel() == n_items
        return out
    if n_items is not None:
        sizes = [shard_range(n_items, r, world)[1] - shard_range(n_items, r, world)[0]
                 for r in range(world)]
        assert local.numel() == sizes[dist.get_rank(group)], "local shard does not match shard_range"
    else:
        n_local = torch.tensor([local.numel()], dtype=torch.int64, device=local.device)
        sz = [torch.zeros_like(n_local) for _ in range(world)]
        dist.all_gather(sz, n_local, group=group)
        sizes = [int(s.item()) for s in sz]
    nmax = max(sizes)
    if min(sizes) == nmax:
        return gather_lnlike(local, None, equal_shards=True, group=group)
    pad = torch.zeros(nmax, dtype=local.dtype, device=local.device)
    pad[: local.numel()] = local
    out = torch.empty(world * nmax, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    return torch.cat([out[r * nmax: r * nmax + s] for r, s in enumerate(sizes)])


def _slice_hyper(kwargs, b0, b1, B):
    """Per-rank slice of every batched hyperparameter (length-B tensors / arrays / lists); scalars
    and everything else pass through."""
    out = {}
    for k, v in kwargs.items():
        if k in _HYPER and v is not None and not isinstance(v, (int, float)):
            tv = torch.as_tensor(v)
            if tv.ndim >= 1 and tv.shape[0] == B and B > 1:
                out[k] = tv[b0:b1]
                continue
        out[k] = v
    return out


def _batch_size(kwargs):
    B = 1
    for k in _HYPER:
        v = kwargs.get(k, None)
        if v is None or isinstance(v, (int, float)):
            continue
        tv = torch.as_tensor(v)
        if tv.ndim >= 1:
            B = max(B, int(tv.shape[0]))
    return B


def log_likelihood_sharded(hyper, t, flux, data_cov, group=None, evaluate=None, gather=True,
                           process_kwargs=None, **ll_kwargs):
    """One sweep of ``B`` hyperparameter samples, split ``B / world_size`` per rank.

    ``hyper``: the FULL batch on every rank -- a dict of ``StarryProcess`` hyperparameters
    (``r, mu, sigma | a, b, c, n[, tau]``), each a scalar or a length-``B`` tensor; further
    constructor keywords go in ``process_kwargs``.  ``flux`` may be ``(nt,)``, ``(M, nt)`` (shared)
    or ``(B, M, nt)`` (per sample, sliced with the batch); per-sample ``i`` / ``baseline_mean`` /
    ``baseline_var`` vectors in ``ll_kwargs`` are sliced as well.  Every rank evaluates
    ``StarryProcess(**shard).log_likelihood(t, flux, data_cov, **ll_kwargs)`` on its
    ``shard_range`` and the ``(B,)`` result is all-gathered (the only collective; ``gather=False``
    returns the local shard and its range instead).

    ``evaluate(shard_hyper, t, flux, data_cov, **ll_kwargs) -> (n_local,) tensor`` replaces the
    per-rank evaluation (the gloo tests run the CPU oracle through it)."""
    rank, world = _world(group)
    B = _batch_size(hyper)
    b0, b1 = shard_range(B, rank, world)
    shard = _slice_hyper(hyper, b0, b1, B)
    f = flux
    if hasattr(f, "ndim") and f.ndim == 3 and f.shape[0] == B:
        f = f[b0:b1]
    kw = dict(ll_kwargs)
    for k in ("i", "baseline_mean", "baseline_var"):
        v = kw.get(k, None)
        if v is not None and hasattr(v, "ndim") and v.ndim == 1 and v.shape[0] == B and B > 1:
            kw[k] = v[b0:b1]
    if evaluate is None:
        from .sp import StarryProcess

        gp = StarryProcess(**shard, **(process_kwargs or {}))
        local = gp.log_likelihood(t, f, data_cov, **kw).reshape(-1)
    else:
        local = evaluate(shard, t, f, data_cov, **kw).reshape(-1)
    if not gather:
        return local, (b0, b1)
    return gather_lnlike(local, n_items=B, group=group)


def ensemble_log_likelihood_sharded(hyper, t, flux, data_cov, group=None, evaluate=None,
                                    process_kwargs=None, **ll_kwargs):
    """configs[1] across ranks: ``flux (M, nt)`` light curves sharing ONE hyperparameter set.  The
    factorisation of the single ``K`` is replicated on every rank, the light curves (right-hand-side
    columns) are split, and because

        lnlike = sum_m [ -1/2 r_m^T K^-1 r_m - sum log L_ii - nt/2 log 2 pi ]      (sp.py:1163-1176)

    is a sum over light curves, the joint value is ONE all-reduce(sum) of every rank's joint
    log-likelihood of its own columns (``-inf`` on any rank -- non-PD ``K`` -- stays ``-inf``)."""
    rank, world = _world(group)
    M = flux.shape[0]
    m0, m1 = shard_range(M, rank, world)
    if evaluate is None:
        from .sp import StarryProcess

        gp = StarryProcess(**hyper, **(process_kwargs or {}))
        if m1 > m0:
            local = gp.log_likelihood(t, flux[m0:m1], data_cov, **ll_kwargs).reshape(1)
        else:
            local = torch.zeros(1, dtype=torch.float64, device=gp.device)
    else:
        local = evaluate(hyper, t, flux[m0:m1], data_cov, **ll_kwargs).reshape(1) if m1 > m0 \
            else torch.zeros(1, dtype=torch.float64)
    local = local.clone()
    if world > 1:
        dist.all_reduce(local, op=dist.ReduceOp.SUM, group=group)
    return local[0]


def design_matrix_sharded(gp, t, i, p=1.0, u=None, group=None, axis=None, evaluate=None):
    """configs[4] across ranks: the design matrix ``A (I, nt, 256)`` of ``I`` inclinations, split
    along the inclination axis (``axis="i"``, default when ``I >= world_size``) or the time axis
    (``axis="t"``).  No collective: every rank returns its own block and the ``(begin, end)`` range
    it covers along the split axis (13 GB of rows at nt = 1e5, I = 64 stay where they are made)."""
    rank, world = _world(group)
    inc = torch.as_tensor(i, dtype=torch.float64).reshape(-1)
    tt = torch.as_tensor(t, dtype=torch.float64).reshape(-1)
    if axis is None:
        axis = "i" if inc.numel() >= world else "t"
    if axis == "i":
        b0, b1 = shard_range(inc.numel(), rank, world)
        inc_l, t_l = inc[b0:b1], tt
    elif axis == "t":
        b0, b1 = shard_range(tt.numel(), rank, world)
        inc_l, t_l = inc, tt[b0:b1]
    else:
        raise ValueError("axis must be 'i' or 't'")
    if b1 == b0:
        return None, (b0, b1), axis
    fn = gp.design_matrix if evaluate is None else evaluate
    A = fn(t_l, inc_l, p, u)
    return A, (b0, b1), axis
