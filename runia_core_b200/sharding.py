"""Multi-GPU partitioning of the scoring path (one process per GPU, torch.distributed).

* Row scorers (entropy, PCA, LaREM, Mahalanobis, ViM, DDU, logit scores, ReAct/DICE/ASH) shard
  test rows contiguously; there is NO data-path collective -- `gather_rows` only exists for a
  caller that wants the full score vector on every rank.
* kNN / KDE shard the BANK by contiguous row ranges; queries are replicated; each rank emits a
  partial result and one exchange step merges them: all-gather of the per-rank top-k
  (float64 distance, global int64 index) followed by the (distance, index) merge kernel, or a
  MAX / SUM all-reduce of the running (max, sum-exp) pair for the KDE.
The `*_fn` hooks let the host logic run on CPU tensors over gloo in the unit tests."""
from typing import Callable, Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist


def row_shard(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced [lo, hi) slice of n rows for `rank` (first n % world ranks get one more)."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _world(group=None):
    if not dist.is_available() or not dist.is_initialized():
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


def gather_rows(local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """All-gathers row shards produced with `row_shard` into the full [n_total, ...] tensor."""
    rank, world = _world(group)
    if world == 1:
        return local
    sizes = [row_shard(n_total, r, world) for r in range(world)]
    maxn = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((maxn,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([b[: hi - lo] for b, (lo, hi) in zip(bufs, sizes)])


def _all_ok(ok: bool, like: torch.Tensor, group=None) -> bool:
    """True on every rank iff `ok` on every rank: one MAX all-reduce of a flag, issued BEFORE the data collectives so
    that a rank-local failure (bad input, out of memory, unsupported shape) makes every rank raise together instead of
    leaving the others blocked in all_gather / all_reduce."""
    _, world = _world(group)
    if world == 1:
        return ok
    flag = torch.tensor([0 if ok else 1], dtype=torch.int32, device=like.device)
    dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=group)
    return int(flag.item()) == 0


def knn_search_sharded(qn: torch.Tensor, bank_shard, k: int, group=None,
                       search_fn: Optional[Callable] = None, merge_fn: Optional[Callable] = None):
    """qn: replicated normalised queries; bank_shard: this rank's KNNBank (idx_offset = first
    global row of the shard; may hold zero rows).  Returns (dist [Nq,k] f32, idx [Nq,k] i64, kth [Nq] f32), identical
    on every rank and identical to a single-GPU search over the concatenated bank."""
    if search_fn is None or merge_fn is None:
        from . import _ops

        search_fn = search_fn or (lambda q, b, kk: (lambda r: (r["dist64"], r["idx"]))(
            _ops.knn_search(q, b, kk, want_f64=True, want_dist=False, check_status=False)))
        merge_fn = merge_fn or _ops.topk_merge
    rank, world = _world(group)
    err = None
    nq = qn.shape[0]
    try:
        if bank_shard.bank.shape[0] == 0:  # an empty shard contributes an exhausted list
            d64 = torch.full((nq, k), float("inf"), dtype=torch.float64, device=qn.device)
            idx = torch.full((nq, k), -1, dtype=torch.int64, device=qn.device)
        else:
            d64, idx = search_fn(qn, bank_shard, k)
    except Exception as e:  # noqa: BLE001 -- re-raised below on every rank
        err = e
        d64 = torch.zeros((nq, k), dtype=torch.float64, device=qn.device)
        idx = torch.zeros((nq, k), dtype=torch.int64, device=qn.device)
    if not _all_ok(err is None, qn, group):
        raise RuntimeError(f"sharded kNN: rank {rank} failed: {err!r}" if err is not None
                           else "sharded kNN: another rank failed its local search") from err
    if world == 1:
        return merge_fn(d64.unsqueeze(0), idx.unsqueeze(0))
    gd = [torch.empty_like(d64) for _ in range(world)]
    gi = [torch.empty_like(idx) for _ in range(world)]
    dist.all_gather(gd, d64.contiguous(), group=group)
    dist.all_gather(gi, idx.contiguous(), group=group)
    return merge_fn(torch.stack(gd), torch.stack(gi))


def kde_score_sharded(q, kde_shard, group=None, partial_fn: Optional[Callable] = None) -> torch.Tensor:
    """Log-density under the KDE of the whole bank from per-rank partial (max, sum-exp) pairs.
    `kde_shard.n_total` must be the TOTAL bank size.  Returns float64 [Nq] on every rank."""
    if partial_fn is None:
        from . import _ops

        partial_fn = lambda qq, kb: _ops.kde_score(qq, kb, partial=True)  # noqa: E731
    rank, world = _world(group)
    err = None
    try:
        if kde_shard.bank.shape[0] == 0:
            dev = kde_shard.bank.device
            m = torch.full((q.shape[0],), float("-inf"), dtype=torch.float32, device=dev)
            s = torch.zeros((q.shape[0],), dtype=torch.float32, device=dev)
        else:
            m, s = partial_fn(q, kde_shard)
    except Exception as e:  # noqa: BLE001
        err = e
        m = torch.zeros((q.shape[0],), dtype=torch.float32, device=kde_shard.bank.device)
        s = torch.zeros_like(m)
    if not _all_ok(err is None, m, group):
        raise RuntimeError(f"sharded KDE: rank {rank} failed: {err!r}" if err is not None
                           else "sharded KDE: another rank failed its local pass") from err
    M = m.clone()
    if world > 1:
        dist.all_reduce(M, op=dist.ReduceOp.MAX, group=group)
    S = s.to(torch.float64) * torch.exp((m - M).to(torch.float64))
    S = torch.where(torch.isinf(m) & (m < 0), torch.zeros_like(S), S)
    if world > 1:
        dist.all_reduce(S, op=dist.ReduceOp.SUM, group=group)
    d = kde_shard.bank.shape[1]
    h = kde_shard.bandwidth
    log_norm = np.log(kde_shard.n_total) + 0.5 * d * np.log(2.0 * np.pi * h * h)
    return M.to(torch.float64) + torch.log(S) - log_norm


def merge_topk_reference(part_d: torch.Tensor, part_i: torch.Tensor):
    """Host restatement of the merge kernel's total order (distance, index); used by the CPU
    (gloo) tests of the exchange logic."""
    R, nq, k = part_d.shape
    d = part_d.permute(1, 0, 2).reshape(nq, R * k).numpy()
    i = part_i.permute(1, 0, 2).reshape(nq, R * k).numpy()
    dd = np.where(i < 0, np.inf, d)
    out_d = np.full((nq, k), np.finfo(np.float32).max, np.float32)
    out_i = np.full((nq, k), -1, np.int64)
    for r in range(nq):
        order = np.lexsort((i[r], dd[r]))[:k]
        ok = i[r][order] >= 0
        out_d[r, : ok.sum()] = dd[r][order][ok].astype(np.float32)
        out_i[r, : ok.sum()] = i[r][order][ok]
    return torch.from_numpy(out_d), torch.from_numpy(out_i), torch.from_numpy(out_d[:, -1].copy())


def combine_class_means(local_means: torch.Tensor, local_counts: torch.Tensor, group=None):
    """Per-rank class means [C, d] float32 + counts [C] -> the means of the whole bank (count-weighted, combined in
    float64, returned as float32 [C, d]) and the global counts; identical on every rank (all-gather, fixed rank order).
    A class that is empty everywhere keeps NaN, like `x[labels == c].mean(0)`."""
    _, world = _world(group)
    cnt = local_counts.to(device=local_means.device, dtype=torch.float64)
    if world == 1:
        return local_means, local_counts.to(torch.int64)
    means = [torch.empty_like(local_means) for _ in range(world)]
    cnts = [torch.empty_like(cnt) for _ in range(world)]
    dist.all_gather(means, local_means.contiguous(), group=group)
    dist.all_gather(cnts, cnt, group=group)
    total = torch.stack(cnts).sum(0)                                                   # [C]
    acc = torch.zeros_like(local_means, dtype=torch.float64)
    for m, c in zip(means, cnts):                                                      # empty shards contribute 0, not NaN
        acc += torch.where(c[:, None] > 0, m.to(torch.float64) * c[:, None], torch.zeros((), dtype=torch.float64, device=m.device))
    gm = (acc / total[:, None]).to(torch.float32)                                      # 0 / 0 = NaN for an empty class
    return gm, total.to(torch.int64)


def fit_mean_precision_sharded(x_local, labels_local=None, num_classes: int = 1, group=None):
    """setup() statistics of MDLatentSpace / cMD / Mahalanobis over a bank whose rows are sharded across the ranks
    ("setup() statistics"): class means from the per-rank NumPy-ordered means, the float64 Gram
    matrix of the class-centred residuals per rank, one all-reduce of d*d + d doubles, then `pinvh` on every rank
    (replicated, identical).  Returns (means [C, d] float32 ndarray, counts [C], precision [d, d] float64)."""
    from scipy.linalg import pinvh

    from . import _ops

    n_local = int(x_local.shape[0])
    if n_local == 0:  # an empty shard contributes zero counts and a zero Gram matrix
        from ._device import device as _dev

        d = int(x_local.shape[1])
        lm = torch.full((num_classes, d), float("nan"), dtype=torch.float32, device=_dev())
        lc = np.zeros(num_classes, np.int64)
    else:
        lm, lc, xf, lab = _ops.class_means(x_local, labels_local, num_classes)
    gm, total = combine_class_means(lm, torch.from_numpy(lc), group)
    if n_local == 0:
        G = torch.zeros((d, d), dtype=torch.float64, device=lm.device)
        cs = torch.zeros((d,), dtype=torch.float64, device=lm.device)
    else:
        G, cs = _ops.centered_gram(xf, lab, gm)
    _, world = _world(group)
    if world > 1:
        dist.all_reduce(G, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(cs, op=dist.ReduceOp.SUM, group=group)
    n_used = int(total.sum())
    cov = _ops.covariance_from_gram(G, cs, n_used)
    return gm.cpu().numpy(), total.cpu().numpy(), pinvh(cov, check_finite=False)


# ------------------------------------------------------------------------------------------------------------------
# Bank sharding behind the reference-facing classes (KNNLatentSpace / KNN / KDELatentSpace / MDLatentSpace.setup take
# `bank_group=`): what the classes call.
# ------------------------------------------------------------------------------------------------------------------
def shard_of_bank(n_rows: int, group=None, rows_are_local: bool = False):
    """(lo, hi, n_total) of this rank's slice of a bank of `n_rows` rows: either the full bank is handed to every rank
    (replicated input, the drop-in case: each rank keeps rows [lo, hi)) or every rank hands in its own rows
    (`rows_are_local`: offsets from an all-gather of the local counts)."""
    rank, world = _world(group)
    if not rows_are_local:
        lo, hi = row_shard(n_rows, rank, world)
        return lo, hi, n_rows
    if world == 1:
        return 0, n_rows, n_rows
    counts = [None] * world
    dist.all_gather_object(counts, int(n_rows), group=group)
    lo = sum(counts[:rank])
    return lo, lo + n_rows, sum(counts)
