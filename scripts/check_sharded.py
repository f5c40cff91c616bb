"""Multi-GPU parity of the bank-sharded scorers over NCCL (run under torchrun, one rank per GPU):
every rank holds a contiguous slice of the bank, queries are replicated, and the merged result must be
IDENTICAL on every rank to a single-GPU search over the whole bank (kNN: indices and float32 distances bit
for bit; KDE: log-densities to 1e-6 relative -- the per-shard partial sums travel as float32).  Prints one JSON line on rank 0.

    torchrun --nproc-per-node 2 scripts/check_sharded.py
"""
import json
import os
import sys

if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
    os.environ["NCCL_DEBUG"] = "WARN"  # keep stdout to the one JSON line

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from runia_core_b200 import _ops, sharding  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    g = torch.Generator(device=dev).manual_seed(1234)  # same stream on every rank: replicated data
    nb, d, nq, k = 200_003, 256, 3001, 50
    bank = torch.randn(nb, d, generator=g, device=dev)
    bank[1000:1100] = bank[1000]  # ties across shard-local candidate lists
    q = torch.randn(nq, d, generator=g, device=dev) + 0.2 * bank[torch.randperm(nb, generator=g, device=dev)[:nq]]
    bn, qn = _ops.normalize_rows(bank), _ops.normalize_rows(q)
    full = _ops.knn_search(qn, _ops.knn_bank(bn), k)
    lo, hi = sharding.row_shard(nb, rank, world)
    shard = _ops.knn_bank(bn[lo:hi].contiguous(), idx_offset=lo)
    dd, ii, kth = sharding.knn_search_sharded(qn, shard, k)
    ok_knn = bool(torch.equal(ii, full["idx"]) and torch.equal(dd, full["dist"]) and torch.equal(kth, full["kth"]))
    kb_full = _ops.kde_bank(bank)
    ref = _ops.kde_score(q, kb_full)
    kb = _ops.kde_bank(bank[lo:hi].contiguous(), center=kb_full.center, n_total=nb)
    got = sharding.kde_score_sharded(q, kb)
    err = float(((got - ref).abs() / ref.abs().clamp(min=1.0)).max())
    # every rank must hold the same merged answer
    chk = torch.stack([ii.double().sum(), dd.double().sum(), got.sum()])
    lst = [torch.empty_like(chk) for _ in range(world)]
    dist.all_gather(lst, chk)
    same = all(bool(torch.equal(lst[0][:2], t[:2])) and float((lst[0][2] - t[2]).abs()) < 1e-6 for t in lst)
    # setup() statistics over a row-sharded bank vs the single-GPU fit of the whole bank
    from scipy.linalg import pinvh
    C, nf, df = 7, 60_001, 128
    labels = torch.randint(0, C, (nf,), generator=g, device=dev)
    labels[labels == 3] = 4  # an empty class
    mu = torch.randn(C, df, generator=g, device=dev)
    feats = (mu[labels] + torch.randn(nf, df, generator=g, device=dev)).contiguous()
    lab_np = labels.cpu().numpy()
    m1, c1, xf1, l1 = _ops.class_means(feats, lab_np, C)
    cov1 = _ops.centered_covariance(xf1, l1, m1, int(c1.sum()))
    p1 = pinvh(cov1, check_finite=False)
    flo, fhi = sharding.row_shard(nf, rank, world)
    m2, c2, p2 = sharding.fit_mean_precision_sharded(feats[flo:fhi].contiguous(), lab_np[flo:fhi], C)
    m1h = m1.cpu().numpy()
    okc = np.arange(C) != 3
    mean_err = float(np.abs(m2[okc] - m1h[okc]).max())
    prec_err = float(np.abs(p2 - p1).max() / np.abs(p1).max())
    # the single-GPU means are float32 row-by-row sums (NumPy order, ~1e-5 of rounding at 8.5k rows per class); the
    # sharded combination is at least as accurate
    ok_fit = bool(np.array_equal(c2, c1) and np.isnan(m2[3]).all() and mean_err < 5e-5 and prec_err < 1e-6)
    # ---- the same partitioning behind the reference-facing classes: setup(..., bank_group=group) ----
    import runia_core_b200 as R
    import warnings

    grp = dist.group.WORLD
    tr_np, q_np = bank[:60_000].cpu().numpy(), q[:500].cpu().numpy()
    single, sharded = R.inference.KNNLatentSpace(), R.inference.KNNLatentSpace()
    single.setup(tr_np)
    sharded.setup(tr_np, bank_group=grp)                      # replicated input: every rank keeps its slice
    ok_api = np.array_equal(single.postprocess(q_np), sharded.postprocess(q_np))
    lo2, hi2 = sharding.row_shard(tr_np.shape[0], rank, world)
    local = R.inference.KNNLatentSpace()
    local.setup(tr_np[lo2:hi2], bank_group=grp, bank_rows_are_local=True)   # every rank hands in its own rows
    ok_api = ok_api and np.array_equal(single.postprocess(q_np), local.postprocess(q_np))
    knn_b = R.inference.KNN(flip_sign=False, k_neighbors=20)
    knn_s = R.inference.KNN(flip_sign=False, k_neighbors=20)
    knn_b.setup(tr_np, valid_feats=q_np)
    knn_s.setup(tr_np, valid_feats=q_np, bank_group=grp)
    ok_api = ok_api and np.array_equal(knn_b.postprocess(q_np), knn_s.postprocess(q_np)) and knn_b.threshold == knn_s.threshold
    kde_b, kde_s = R.inference.KDELatentSpace(), R.inference.KDELatentSpace()
    kde_b.setup(tr_np[:20_000])
    kde_s.setup(tr_np[:20_000], bank_group=grp)
    e_kde = float(np.max(np.abs(kde_b.postprocess(q_np) - kde_s.postprocess(q_np)) / np.maximum(1.0, np.abs(kde_b.postprocess(q_np)))))
    md_b, md_s = R.inference.MDLatentSpace(), R.inference.MDLatentSpace()
    md_b.setup(tr_np)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        md_s.setup(tr_np, bank_group=grp)
    e_md = float(np.max(np.abs(md_b.postprocess(q_np) - md_s.postprocess(q_np)) / np.maximum(1.0, np.abs(md_b.postprocess(q_np)))))
    ok_api = bool(ok_api and e_kde < 1e-6 and e_md < 1e-6)
    # a rank-local failure must raise on EVERY rank (flag all-reduce before the collectives), not hang the others
    raised = False
    try:
        sharding.knn_search_sharded(qn, shard, k, search_fn=(lambda *a: (_ for _ in ()).throw(ValueError("boom")))
                                    if rank == world - 1 else None)
    except RuntimeError:
        raised = True
    flags = torch.tensor([int(ok_knn), int(err < 1e-6), int(same), int(ok_fit), int(ok_api), int(raised)], device=dev)
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({"world": world, "bank_rows": nb, "queries": nq, "k": k, "knn_bit_exact": bool(flags[0]),
                          "kde_max_rel_err": err, "kde_ok": bool(flags[1]), "identical_on_all_ranks": bool(flags[2]),
                          "sharded_fit_ok": bool(flags[3]), "sharded_fit_mean_abs_err": mean_err,
                          "sharded_fit_precision_rel_err": prec_err,
                          "postprocessors_with_bank_group_ok": bool(flags[4]), "api_kde_rel_err": e_kde, "api_md_rel_err": e_md,
                          "rank_local_failure_raises_everywhere": bool(flags[5])}))
    dist.destroy_process_group()
    sys.exit(0 if bool(flags.min()) else 1)


if __name__ == "__main__":
    main()
