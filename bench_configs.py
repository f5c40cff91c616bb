"""BASELINE.json configs[2], configs[3] and configs[4] at their STATED sizes (SURVEY.md section 8d inputs), run from
bench.py into `extra.config3 / config4 / config5`.  Each entry carries `ms`, a `roofline` for its dominant kernel and
`parity_max_rel_err`: the CUDA path against a brute-force restatement of the definition in float64 torch ops on a
sample of the same rows, computed in the same run (the CPU oracle of `oracle/` checks the same kernels in `tests/`;
bench.py may run the oracle only in its cpu_baseline leg, so the in-run sample is restated here from the reference
lines cited per function).

    config3  Faster R-CNN / YOLOv8 object-level LaRED + LaREM: 1M boxes x 16 MC samples x d = 1024 (65.5 GB of
             samples resident in HBM) -> entropy -> PCA-256 -> KDE (1M queries x 100k-box bank) + LaREM
    config4  ViT-B/16 ImageNet-scale kNN: 50,000 queries x 10M x 768 bank, k = 50, bank FIXED at 10M rows and sharded
             over the ranks (strong scaling), NCCL all-gather of the per-rank top-k + merge kernel
    config5  DeepLabv3+ per-pixel LaREM: 64 x 512 x 1024 = 33,554,432 embeddings x 256 in 4 chunks + EigenScore on
             10 samples x 4096 hidden units
"""
import math
import time

import numpy as np


def _events(torch, fn, reps=1, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def _hbm(alg_bytes, ms, peak):
    a = alg_bytes / (ms * 1e-3) / 1e9
    return {"bound": "hbm", "achieved": round(a, 1), "peak": peak, "unit": "GB/s", "frac": round(a / peak, 4)}


def _tensor(flop, ms, tf32_probe, bf16_peak, products=3):
    """FP32-equivalent TFLOP/s of a contraction issued as `products` TF32 products per FLOP (3: the 3xTF32 scorers;
    1: the kNN candidate filter, whose survivors are re-ranked exactly) against (TF32 probe of this run) / products;
    the figure derived from MEASURED_PEAKS.json's dense bf16 number (/ 2 / products) is kept beside it."""
    a = flop / (ms * 1e-3) / 1e12
    peak = (tf32_probe or bf16_peak / 2.0) / products
    out = {"bound": "tensor", "achieved": round(a, 2), "peak": round(peak, 2), "unit": "TFLOP/s",
           "frac": round(a / peak, 4), "tf32_products_per_flop": products,
           "peak_source": f"runia_tf32_peak_probe of this run / {products}" if tf32_probe
           else f"MEASURED_PEAKS bf16 / 2 / {products}"}
    out[f"frac_of_measured_bf16_over_{2 * products}"] = round(a / (bf16_peak / (2.0 * products)), 4)
    return out


def _rel(torch, got, ref):
    got, ref = got.double().reshape(-1), ref.double().reshape(-1)
    return float(((got - ref).abs() / ref.abs().clamp(min=1.0)).max())


# ------------------------------------------------------------------------------------------------------------------
# brute-force float64 restatements (definitions, not the kernels' algorithms)
# ------------------------------------------------------------------------------------------------------------------
def ref_entropy(torch, z_items):
    """evaluation/entropy.py:56-84 with entropy_estimators.get_h(x, k, norm="max", min_dist=1e-5): the k-th
    neighbour distance of every sample from the full pairwise table (what the KD-tree query returns), no sorting
    network, no windows.  z_items [m, n, D] -> (h_mvn [m], h_z [m, D]) float64."""
    from scipy.special import digamma

    x = z_items.double()
    m, n, D = x.shape
    k = 5 if n > 5 else n - 1
    c = float(-digamma(k) + digamma(n))
    xd = x.permute(0, 2, 1)                                       # [m, D, n]
    pd = (xd.unsqueeze(-1) - xd.unsqueeze(-2)).abs()              # [m, D, n, n], self distance 0 included
    r = pd.kthvalue(k + 1, dim=-1).values.clamp_min(1e-5)
    h_z = c + torch.log(2.0 * r).sum(-1) / n
    rj = pd.amax(dim=1).kthvalue(k + 1, dim=-1).values.clamp_min(1e-5)  # Chebyshev distance between the D-vectors
    h_mvn = c + (D / n) * torch.log(2.0 * rj).sum(-1)
    return h_mvn, h_z


def ref_knn(torch, q, bank, k, idx_offset=0, chunk=500_000):
    """faiss.IndexFlatL2.search restated: exact squared L2 in float64 over the whole bank, k smallest by
    (distance, index).  q [m, d], bank [Nb, d] float32 device -> (dist [m, k] f64, idx [m, k] i64)."""
    q64 = q.double()
    best_d = torch.full((q.shape[0], 0), 0.0, dtype=torch.float64, device=q.device)
    best_i = torch.zeros((q.shape[0], 0), dtype=torch.int64, device=q.device)
    for lo in range(0, bank.shape[0], chunk):
        b = bank[lo:lo + chunk].double()
        d2 = (q64 * q64).sum(1, keepdim=True) + (b * b).sum(1)[None, :] - 2.0 * q64 @ b.t()
        kk = min(k, d2.shape[1])
        dd, ii = d2.topk(kk, dim=1, largest=False)
        best_d = torch.cat([best_d, dd], 1)
        best_i = torch.cat([best_i, ii + lo + idx_offset], 1)
        order = torch.argsort(best_d, dim=1, stable=True)[:, :k]
        best_d, best_i = best_d.gather(1, order), best_i.gather(1, order)
    return best_d, best_i


# ------------------------------------------------------------------------------------------------------------------
# config4: kNN, 50k queries x 10M x 768, k = 50, strong scaling over the ranks
# ------------------------------------------------------------------------------------------------------------------
C4_SHARDS, C4_SHARD_ROWS, C4_D, C4_K = 8, 1_250_000, 768, 50


def config4(torch, dist, _ops, world, rank, barrier, max_over_ranks, tf32_probe, bf16_peak, nq=50_000,
            shard_rows=C4_SHARD_ROWS):
    """The bank is always the same 8 seeded blocks of 1.25M normalised rows (10M x 768 = 30.7 GB + its TF32 planes);
    rank r of `world` holds blocks [8 r / world, 8 (r + 1) / world).  Queries (replicated): normalise(randn + 0.1 x
    the first rows of block 0) (queries near bank rows).  Timed: per-rank search + NCCL all-gather of the
    partial top-k + merge, max over ranks (CUDA events).  Parity (64 queries): the merged neighbours against the
    float64 brute force over the WHOLE bank (every rank scans its own blocks, partial results all-gathered), and the
    merged result compared across ranks."""
    dev = torch.device("cuda", torch.cuda.current_device())
    d, k = C4_D, C4_K
    assert C4_SHARDS % world == 0, "config4 needs 1, 2, 4 or 8 ranks"
    per = C4_SHARDS // world
    rows = per * shard_rows
    bank = torch.empty((rows, d), dtype=torch.float32, device=dev)
    first = None
    for j in range(per):
        b = rank * per + j
        g = torch.Generator(device=dev).manual_seed(100 + b)
        raw = torch.randn(shard_rows, d, generator=g, device=dev)
        if b == 0:
            first = raw[:nq].clone()
        bank[j * shard_rows:(j + 1) * shard_rows] = _ops.normalize_rows(raw)
        del raw
    if first is None:  # ranks that do not hold block 0 regenerate it for the queries
        g = torch.Generator(device=dev).manual_seed(100)
        first = torch.randn(shard_rows, d, generator=g, device=dev)[:nq].clone()
    gq = torch.Generator(device=dev).manual_seed(4)
    q = _ops.normalize_rows(torch.randn(nq, d, generator=gq, device=dev) + 0.1 * first[:nq])
    del first
    kb = _ops.knn_bank(bank, idx_offset=rank * rows)
    out = {}

    def step():
        r = _ops.knn_search(q, kb, k, want_f64=True, want_dist=False, check_status=False)
        if world == 1:
            gd, gi = r["dist64"].unsqueeze(0), r["idx"].unsqueeze(0)
        else:
            gd = torch.empty((world, nq, k), dtype=torch.float64, device=dev)
            gi = torch.empty((world, nq, k), dtype=torch.int64, device=dev)
            dist.all_gather_into_tensor(gd, r["dist64"])
            dist.all_gather_into_tensor(gi, r["idx"])
        out["m"] = _ops.topk_merge(gd, gi)

    step()
    barrier()
    ms = max_over_ranks(_events(torch, step, reps=1, warm=0))
    md, mi, _ = out["m"]
    # ---- parity sample: 64 queries against the float64 brute force over the whole bank ----
    ns = 64
    sel = torch.linspace(0, nq - 1, ns, device=dev).long()
    ld, li = ref_knn(torch, q[sel], bank, k, idx_offset=rank * rows)
    if world > 1:
        gd = [torch.empty_like(ld) for _ in range(world)]
        gi = [torch.empty_like(li) for _ in range(world)]
        dist.all_gather(gd, ld)
        dist.all_gather(gi, li)
        ld, li = torch.cat(gd, 1), torch.cat(gi, 1)
        order = torch.argsort(ld, dim=1, stable=True)[:, :k]
        ld, li = ld.gather(1, order), li.gather(1, order)
    idx_mismatch = int((mi[sel] != li).sum().item())
    rel = _rel(torch, md[sel], ld)
    # ---- the merged answer must be the same on every rank ----
    w = torch.arange(1, k + 1, device=dev, dtype=torch.float64)
    chk = torch.stack([(mi.double() * w).sum(), md.double().sum()])
    same = True
    if world > 1:
        lst = [torch.empty_like(chk) for _ in range(world)]
        dist.all_gather(lst, chk)
        same = all(bool(torch.equal(lst[0], t)) for t in lst)
    flop = 2.0 * nq * (C4_SHARDS * shard_rows) * d
    res = {"ms": ms, "queries_per_s": nq / (ms * 1e-3), "queries": nq, "bank": [C4_SHARDS * shard_rows, d], "k": k,
           "bank_rows_per_rank": rows, "scaling": "strong (bank fixed, sharded over the ranks)", "ranks": world,
           "distance_tflops_total": round(flop / (ms * 1e-3) / 1e12, 1),
           "roofline": _tensor(flop / world, ms, tf32_probe, bf16_peak, products=1),
           "parity_sample": f"{ns} queries vs float64 brute force over all {C4_SHARDS * shard_rows} bank rows",
           "parity_idx_mismatches": idx_mismatch, "parity_max_rel_err": rel,
           "filter_tf32_products": _ops.knn_filter_products(kb, k),
           "merged_identical_on_all_ranks": bool(same),
           "exchange_bytes_per_rank": nq * k * 16}
    del kb, bank, q
    torch.cuda.empty_cache()
    return res


# ------------------------------------------------------------------------------------------------------------------
# config3: object-level LaRED + LaREM on RoI features
# ------------------------------------------------------------------------------------------------------------------
def config3(torch, R, _ops, hbm_peak, tf32_probe, bf16_peak, n_boxes=1_000_000, n_mc=16, D=1024, d_pca=256,
            n_bank=100_000):
    """samples z = base + 0.1 N(0,1) with 40 % of the entries zeroed per MC row (the exact duplicates /
    min_dist clamps real MC-DropBlock samples show), [1M x 16, 1024] f32 = 65.5 GB generated in 8 chunks into one
    resident tensor; ONE entropy launch over all boxes; PCA 1024 -> 256 fitted (sklearn, host) on 20k boxes'
    entropies; KDE bank = the first 100k boxes (the reference caps boxes with subset_boxes, metrics.py:465-509),
    queries = all 1M; LaREM fitted on the bank."""
    from sklearn.decomposition import PCA

    dev = torch.device("cuda", torch.cuda.current_device())
    g = torch.Generator(device=dev).manual_seed(3)
    z = torch.empty((n_boxes * n_mc, D), dtype=torch.float32, device=dev)
    chunks = 8
    per = n_boxes // chunks
    for c in range(chunks):
        base = torch.randn(per, 1, D, generator=g, device=dev)
        blk = base + 0.1 * torch.randn(per, n_mc, D, generator=g, device=dev)
        blk.mul_((torch.rand(per, n_mc, D, generator=g, device=dev) >= 0.4).to(torch.float32))
        z[c * per * n_mc:(c + 1) * per * n_mc] = blk.reshape(per * n_mc, D)
        del base, blk
    res = {"boxes": n_boxes, "n_mc": n_mc, "D": D, "sample_bytes": z.numel() * 4}
    hold = {}

    def ent():
        hold["e"] = _ops.mcd_entropy(z, n_mc)

    ms_e = _events(torch, ent, reps=2, warm=1)
    h_mvn, h_z = hold["e"]
    sel = torch.linspace(0, n_boxes - 1, 64, device=dev, dtype=torch.float64).long()
    zs = torch.stack([z[i * n_mc:(i + 1) * n_mc] for i in sel.tolist()])
    rm, rz = ref_entropy(torch, zs)
    par_e = max(_rel(torch, h_z[sel], rz), _rel(torch, h_mvn[sel], rm))
    alg = n_boxes * (n_mc * D * 4 + D * 8 + 8)
    res["entropy"] = {"ms": ms_e, "boxes_per_s": n_boxes / (ms_e * 1e-3), "roofline": _hbm(alg, ms_e, hbm_peak),
                      "parity_max_rel_err": par_e, "parity_sample": "64 boxes vs float64 pairwise-table k-NN distances"}
    del z, hold["e"]
    torch.cuda.empty_cache()
    # ---- PCA 1024 -> 256 on the entropies (float64 [1M, 1024], as get_dl_h_z returns them) ----
    np.random.seed(1)
    fit_rows = h_z[:20_000].cpu().numpy()
    pca = PCA(n_components=d_pca, svd_solver="randomized", whiten=True).fit(fit_rows)
    st = _ops.pca_prepare(pca.mean_, pca.components_, pca.explained_variance_, True)

    def proj():
        hold["p"] = _ops.pca_transform(h_z, st)

    ms_p = _events(torch, proj, reps=2, warm=1)
    zp = hold["p"]
    mean = torch.from_numpy(pca.mean_).to(dev)
    comp = torch.from_numpy(pca.components_).to(dev)
    sc = torch.from_numpy(np.sqrt(pca.explained_variance_)).to(dev)
    rp = ((h_z[sel] - mean) @ comp.t()) / sc
    res["pca"] = {"ms": ms_p, "boxes_per_s": n_boxes / (ms_p * 1e-3), "input": "float64 [1M, 1024] (centre + cast kernel, then the tcgen05 projection)",
                  "roofline": _hbm(n_boxes * (D * 8 + d_pca * 4), ms_p, hbm_peak),
                  "fp32_equiv_tflops": round(2.0 * n_boxes * D * d_pca / (ms_p * 1e-3) / 1e12, 1),
                  "parity_max_rel_err": _rel(torch, zp[sel], rp)}
    del h_z
    torch.cuda.empty_cache()
    # ---- LaRED: KDE of 1M queries against the 100k-box bank ----
    bank = zp[:n_bank].contiguous()
    kb = _ops.kde_bank(bank)

    def kde():
        hold["k"] = _ops.kde_score(zp, kb)

    ms_k = _events(torch, kde, reps=1, warm=1)
    qs = zp[sel].double()
    d2 = torch.cdist(qs, bank.double()).pow(2)
    rk = torch.logsumexp(-0.5 * d2, dim=1) - math.log(n_bank) - 0.5 * d_pca * math.log(2.0 * math.pi)
    flop = 2.0 * n_boxes * n_bank * d_pca
    res["kde"] = {"ms": ms_k, "queries_per_s": n_boxes / (ms_k * 1e-3), "bank": [n_bank, d_pca],
                  "roofline": _tensor(flop, ms_k, tf32_probe, bf16_peak), "parity_max_rel_err": _rel(torch, hold["k"][sel], rk)}
    # ---- LaREM fitted on the bank, all 1M boxes scored ----
    md = R.inference.MDLatentSpace()
    md.setup(bank.cpu().numpy())

    def larem():
        hold["m"] = _ops.md_score(zp, md._state, torch.float64)

    ms_m = _events(torch, larem, reps=3, warm=1)
    mu = torch.from_numpy(np.asarray(md.feats_mean, np.float64).reshape(-1)).to(dev)
    P = torch.from_numpy(np.asarray(md.precision, np.float64)).to(dev)
    df = zp[sel].double() - mu
    rmd = -torch.einsum("ij,jk,ik->i", df, P, df)
    res["larem"] = {"ms": ms_m, "boxes_per_s": n_boxes / (ms_m * 1e-3),
                    "roofline": _tensor(n_boxes * (2.0 * d_pca * d_pca + 3 * d_pca), ms_m, tf32_probe, bf16_peak),
                    "parity_max_rel_err": _rel(torch, hold["m"][sel], rmd)}
    res["ms"] = ms_e + ms_p + ms_k + ms_m
    res["boxes_per_s"] = n_boxes / (res["ms"] * 1e-3)
    res["roofline"] = res["kde"]["roofline"]  # the dominant kernel of the chain
    res["parity_max_rel_err"] = max(res[s]["parity_max_rel_err"] for s in ("entropy", "pca", "kde", "larem"))
    hold.clear()
    del zp, bank, kb
    torch.cuda.empty_cache()
    return res


# ------------------------------------------------------------------------------------------------------------------
# config5: per-pixel LaREM at 64 x 512 x 1024 pixels + EigenScore
# ------------------------------------------------------------------------------------------------------------------
def config5(torch, R, _ops, md, hbm_peak, tf32_probe, bf16_peak, n_pix=64 * 512 * 1024, d=256, chunks=4):
    """33,554,432 pixel embeddings x 256 (34.4 GB, resident) through the LaREM kernel fitted on 50k rows, in 4
    launches of 8.4M rows; EigenScore of 10 sampled generations x 4096 hidden units (seed 42)."""
    dev = torch.device("cuda", torch.cuda.current_device())
    g = torch.Generator(device=dev).manual_seed(5)
    per = n_pix // chunks
    X = torch.empty((n_pix, d), dtype=torch.float32, device=dev)
    for c in range(chunks):
        X[c * per:(c + 1) * per] = torch.randn(per, d, generator=g, device=dev)
    out = torch.empty((n_pix,), dtype=torch.float64, device=dev)
    st = md._state

    def run():
        for c in range(chunks):
            _ops.md_score(X[c * per:(c + 1) * per], st, torch.float64, out=out[c * per:(c + 1) * per])

    ms = _events(torch, run, reps=3, warm=1)
    sel = torch.linspace(0, n_pix - 1, 4096, device=dev, dtype=torch.float64).long()
    mu = torch.from_numpy(np.asarray(md.feats_mean, np.float64).reshape(-1)).to(dev)
    P = torch.from_numpy(np.asarray(md.precision, np.float64)).to(dev)
    df = X[sel].double() - mu
    ref = -torch.einsum("ij,jk,ik->i", df, P, df)
    res = {"pixels": n_pix, "d": d, "chunks": chunks, "ms": ms, "embeddings_per_s": n_pix / (ms * 1e-3),
           "roofline": _tensor(n_pix * (2.0 * d * st.r + 3 * d), ms, tf32_probe, bf16_peak),
           "hbm": _hbm(n_pix * (d * 4 + 8), ms, hbm_peak),
           "parity_max_rel_err": _rel(torch, out[sel], ref), "parity_sample": "4096 pixels vs float64 einsum",
           "launches": chunks}
    del X, out
    torch.cuda.empty_cache()
    # EigenScore (llm_uncertainty/scores.py:49-66): 10 x 4096
    ge = torch.Generator(device="cpu").manual_seed(42)
    E = torch.randn(10, 4096, generator=ge)
    hs = ((None,) * 15 + (E.unsqueeze(0),),)  # hidden_states[-1][15].squeeze() -> [10, 4096]
    t0 = time.perf_counter()
    for _ in range(20):
        got = R.llm_uncertainty.eigen_score(hs, alpha=1e-3)
    us = (time.perf_counter() - t0) / 20 * 1e6
    cov = torch.cov(E.double().t()).numpy()
    sv = np.linalg.svd(cov + 1e-3 * np.eye(cov.shape[0]), compute_uv=False)
    ref_e = float(np.mean(np.log(sv)))
    res["eigen_score_10x4096"] = {"us_per_call_host_to_host": us, "value": got, "reference_svd_value": ref_e,
                                  "parity_max_rel_err": abs(got - ref_e) / max(1.0, abs(ref_e))}
    return res


# ------------------------------------------------------------------------------------------------------------------
# KDE with the bank sharded over the ranks: MAX / SUM all-reduce of the running (max, sum-exp) pair
# ------------------------------------------------------------------------------------------------------------------
def kde_sharded(torch, dist, _ops, sharding, world, rank, barrier, max_over_ranks, tf32_probe, bf16_peak,
                nq=200_000, nb=400_000, d=256):
    dev = torch.device("cuda", torch.cuda.current_device())
    g = torch.Generator(device=dev).manual_seed(77)  # replicated stream: every rank builds the same bank
    bank = 0.5 + torch.randn(nb, d, generator=g, device=dev)
    q = torch.randn(nq, d, generator=g, device=dev)
    center = bank.double().mean(0)
    lo, hi = sharding.row_shard(nb, rank, world)
    kb = _ops.kde_bank(bank[lo:hi].contiguous(), center=center, n_total=nb)
    hold = {}

    def step():
        hold["s"] = sharding.kde_score_sharded(q, kb)

    step()
    barrier()
    ms = max_over_ranks(_events(torch, step, reps=2, warm=0))
    sel = torch.linspace(0, nq - 1, 128, device=dev, dtype=torch.float64).long()
    d2 = torch.cdist(q[sel].double(), bank.double()).pow(2)
    ref = torch.logsumexp(-0.5 * d2, dim=1) - math.log(nb) - 0.5 * d * math.log(2.0 * math.pi)
    flop = 2.0 * nq * nb * d
    res = {"ms": ms, "queries_per_s": nq / (ms * 1e-3), "queries": nq, "bank": [nb, d], "ranks": world,
           "scaling": "strong (bank fixed, sharded over the ranks)",
           "roofline": _tensor(flop / world, ms, tf32_probe, bf16_peak),
           "parity_max_rel_err": _rel(torch, hold["s"][sel], ref), "collectives": "all_reduce(MAX) + all_reduce(SUM), 2 x 4 B / 8 B per query"}
    del bank, q, kb
    torch.cuda.empty_cache()
    return res
