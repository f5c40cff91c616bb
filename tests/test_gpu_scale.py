"""GPU parity at BASELINE.json's full sizes (-m gpu).  The oracle cannot run these sizes in seconds,
so each case checks (1) a random subset of rows against the oracle and (2) size-independent
properties of the domain: row independence (scoring a slice alone gives bit-identical scores),
invariance of the entropy under permutation of the MC samples of an item, tensor-core vs FP32-SIMT
candidate passes giving the same exact neighbours, and a sharded bank merging to the single-bank
result."""
import numpy as np
import pytest
import torch

from oracle import oracle_np as O
from tests.conftest import rel_err

pytestmark = pytest.mark.gpu
RTOL = 1e-4


@pytest.fixture(scope="module")
def ops():
    from runia_core_b200 import _ops

    return _ops


def _gen(seed):
    return torch.Generator(device="cuda").manual_seed(seed)


def test_larem_4m_rows(ops):
    """LaREM at the bench size: 4,194,304 x 256 (4.3 GB), fit on 50k latents (configs[0]/[1])."""
    rng = np.random.RandomState(1)
    train = (rng.randn(50_000, 256) + 0.05 * rng.randn(1, 256)).astype(np.float32)
    mean, prec = O.md_fit(train)
    st = ops.md_prepare(mean, prec)
    n = 4 * 1024 * 1024
    X = torch.randn(n, 256, generator=_gen(11), device="cuda")
    X[n // 2:] -= 0.5
    s = ops.md_score(X, st)
    assert s.dtype == torch.float64 and s.shape == (n,)
    # row independence: any slice scored alone (different tile / CTA-pair assignment) is bit-identical
    for lo, hi in ((0, 1000), (n // 2 - 77, n // 2 + 1234), (n - 4099, n)):
        assert torch.equal(ops.md_score(X[lo:hi].contiguous(), st), s[lo:hi])
    idx = torch.from_numpy(np.random.RandomState(2).choice(n, 4096, replace=False)).cuda()
    ref = O.md_score(X[idx].cpu().numpy(), mean, prec)
    assert rel_err(s[idx].cpu().numpy(), ref) < RTOL
    # AUROC / FPR@95 of InD half vs shifted half from GPU scores equal the oracle's on the subset
    half = idx < n // 2
    a = O.auroc_fpr95(s[idx][half].cpu().numpy(), s[idx][~half].cpu().numpy())
    b = O.auroc_fpr95(ref[half.cpu().numpy()], ref[(~half).cpu().numpy()])
    assert abs(a[0] - b[0]) < 1e-6 and abs(a[1] - b[1]) < 1e-6


def test_pca_2m_rows(ops):
    rng = np.random.RandomState(3)
    mean = rng.randn(512)
    comp = np.linalg.qr(rng.randn(512, 256))[0].T.copy()
    ev = 1.0 + rng.rand(256)
    st = ops.pca_prepare(mean, comp, ev, True)
    n = 2_000_000
    X = torch.randn(n, 512, generator=_gen(12), device="cuda") + torch.from_numpy(mean).float().cuda()
    Z = ops.pca_transform(X, st)
    for lo, hi in ((0, 513), (n - 3001, n)):
        assert torch.equal(ops.pca_transform(X[lo:hi].contiguous(), st), Z[lo:hi])
    idx = torch.from_numpy(np.random.RandomState(4).choice(n, 2048, replace=False)).cuda()
    ref = O.pca_transform(X[idx].cpu().numpy(), mean, comp, ev)
    assert np.allclose(Z[idx].cpu().numpy(), ref, rtol=RTOL, atol=RTOL)


def test_entropy_config1_size(ops):
    """configs[0]: 60k items x 16 MC samples x 512 dims (1.97 GB)."""
    n_items, n_mc, D = 60_000, 16, 512
    g = _gen(13)
    z = torch.randn(n_items, 1, D, generator=g, device="cuda") + 0.1 * torch.randn(n_items, n_mc, D, generator=g, device="cuda")
    z = z * (torch.rand(n_items, n_mc, D, generator=g, device="cuda") >= 0.4)  # MC-dropout zeros -> duplicates, min_dist clamp
    z2 = z.reshape(n_items * n_mc, D).contiguous()
    hm, hz = ops.mcd_entropy(z2, n_mc)
    assert torch.isfinite(hm).all() and torch.isfinite(hz).all()
    # subset vs oracle
    pick = np.random.RandomState(5).choice(n_items, 48, replace=False)
    zs = z[torch.from_numpy(pick).cuda()].reshape(-1, D).cpu().numpy()
    rm, rz = O.get_dl_h_z(zs, n_mc, chunk=16)
    assert rel_err(hz[pick].cpu().numpy(), rz) < RTOL and rel_err(hm[pick].cpu().numpy(), rm[:, 0]) < RTOL
    # item independence: a slice of items alone gives bit-identical entropies
    lo, hi = 31_111, 31_999
    hm_s, hz_s = ops.mcd_entropy(z2[lo * n_mc:hi * n_mc].contiguous(), n_mc)
    assert torch.equal(hz_s, hz[lo:hi]) and torch.equal(hm_s, hm[lo:hi])
    # permuting the MC samples of every item: per-dimension entropies are bit-identical (sorted
    # statistics), the joint entropy changes only by float summation order
    perm = torch.randperm(n_mc, generator=g, device="cuda")
    zp = z[:4096, perm].reshape(-1, D).contiguous()
    hm_p, hz_p = ops.mcd_entropy(zp, n_mc)
    assert torch.equal(hz_p, hz[:4096])
    assert torch.allclose(hm_p, hm[:4096], rtol=1e-6, atol=1e-4)


def test_knn_config2_size(ops):
    """configs[1]: 50k x 512 bank, 10k queries, k = 50."""
    g = _gen(14)
    centers = torch.randn(10, 512, generator=g, device="cuda")
    lab = torch.randint(0, 10, (50_000,), generator=g, device="cuda")
    bank = ops.normalize_rows(centers[lab] + torch.randn(50_000, 512, generator=g, device="cuda"))
    labq = torch.randint(0, 10, (10_000,), generator=g, device="cuda")
    q = ops.normalize_rows(centers[labq] + torch.randn(10_000, 512, generator=g, device="cuda"))
    k = 50
    full = ops.knn_search(q, ops.knn_bank(bank), k, want_f64=True)
    # tensor-core and FP32-SIMT candidate passes must agree on the exact result
    ops.set_engine("simt")
    try:
        simt = ops.knn_search(q, ops.knn_bank(bank, planes=False), k)
    finally:
        ops.set_engine("tc")
    assert torch.equal(simt["idx"], full["idx"]) and torch.equal(simt["dist"], full["dist"])
    # 32 queries against the oracle's exhaustive search: indices and float32 distances bit-exact
    pick = np.random.RandomState(6).choice(10_000, 32, replace=False)
    D, I = O.flat_l2_search_tree(bank.cpu().numpy(), q[pick].cpu().numpy(), k)
    assert np.array_equal(full["idx"][pick].cpu().numpy(), I)
    assert np.array_equal(full["dist"][pick].cpu().numpy(), D)
    # distances ascending, indices valid and distinct per row
    d = full["dist"]
    assert bool((d[:, 1:] >= d[:, :-1]).all())
    srt = torch.sort(full["idx"], dim=1).values
    assert bool((srt[:, 1:] != srt[:, :-1]).all()) and int(full["idx"].min()) >= 0 and int(full["idx"].max()) < 50_000
    # bank split into 4 shards (what 4 ranks would hold) merges to the same answer
    pd, pi = [], []
    for r in range(4):
        lo, hi = r * 12_500, (r + 1) * 12_500
        rr = ops.knn_search(q, ops.knn_bank(bank[lo:hi].contiguous(), idx_offset=lo), k, want_f64=True)
        pd.append(rr["dist64"])
        pi.append(rr["idx"])
    md, mi, mk = ops.topk_merge(torch.stack(pd), torch.stack(pi))
    assert torch.equal(mi, full["idx"]) and torch.equal(md, full["dist"]) and torch.equal(mk, full["kth"])


def test_kde_50k_bank(ops):
    g = _gen(15)
    bank = 0.5 + torch.randn(50_000, 256, generator=g, device="cuda")
    q = torch.cat([0.5 + torch.randn(1000, 256, generator=g, device="cuda"),
                   -0.5 + torch.randn(1000, 256, generator=g, device="cuda")])
    kb = ops.kde_bank(bank)
    s = ops.kde_score(q, kb)
    pick = np.random.RandomState(7).choice(2000, 24, replace=False)
    ref = O.kde_score(q[pick].cpu().numpy(), bank.cpu().numpy())
    assert rel_err(s[pick].cpu().numpy(), ref) < 1e-5
    # two bank shards, partial (max, sum-exp) pairs combined like sharding.kde_score_sharded
    parts = [ops.kde_score(q, ops.kde_bank(bank[lo:hi].contiguous(), center=kb.center, n_total=50_000), partial=True)
             for lo, hi in ((0, 20_000), (20_000, 50_000))]
    M = torch.maximum(parts[0][0], parts[1][0])
    S = sum(p[1].double() * torch.exp((p[0] - M).double()) for p in parts)
    comb = M.double() + torch.log(S) - (np.log(50_000) + 0.5 * 256 * np.log(2 * np.pi))
    assert torch.allclose(comb, s, rtol=1e-6, atol=1e-5)


def test_more_than_2_31_elements(ops):
    """Inputs with more than 2^31 elements (toward configs[2] / configs[4] sizes: 1M boxes x 16 x 1024,
    33.5M pixel embeddings): the last rows / items must score exactly like the same rows scored alone,
    which catches any 32-bit index arithmetic in the kernels or the TMA coordinates."""
    g = _gen(21)
    # LaREM: 9,000,000 x 256 = 2.3e9 elements (9.2 GB)
    rng = np.random.RandomState(1)
    train = rng.randn(20_000, 256).astype(np.float32)
    mean, prec = O.md_fit(train)
    st = ops.md_prepare(mean, prec)
    n = 9_000_000
    X = torch.empty(n, 256, device="cuda")
    for lo in range(0, n, 1_000_000):
        X[lo:lo + 1_000_000].normal_(generator=g)
    s = ops.md_score(X, st)
    tail = slice(n - 5000, n)
    assert torch.equal(ops.md_score(X[tail].contiguous(), st), s[tail])
    assert rel_err(s[tail].cpu().numpy()[:512], O.md_score(X[tail][:512].cpu().numpy(), mean, prec)) < RTOL
    del X, s
    torch.cuda.empty_cache()
    # entropy: 140,000 items x 16 x 1024 = 2.3e9 elements (9.2 GB in, 1.1 GB out)
    n_items, n_mc, D = 140_000, 16, 1024
    z = torch.empty(n_items * n_mc, D, device="cuda")
    for lo in range(0, n_items * n_mc, 160_000):
        z[lo:lo + 160_000].normal_(generator=g)
    hm, hz = ops.mcd_entropy(z, n_mc)
    lo = n_items - 300
    hm_t, hz_t = ops.mcd_entropy(z[lo * n_mc:].contiguous(), n_mc)
    assert torch.equal(hz_t, hz[lo:]) and torch.equal(hm_t, hm[lo:])
    rm, rz = O.get_dl_h_z(z[(n_items - 16) * n_mc:].cpu().numpy(), n_mc, chunk=16)
    assert rel_err(hz[-16:].cpu().numpy(), rz) < RTOL and rel_err(hm[-16:].cpu().numpy(), rm[:, 0]) < RTOL
