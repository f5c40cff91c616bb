"""CPU (-m "not gpu"): the C-ABI library loads and exports every symbol the header declares; the
reference-facing boundary behaves like the reference where no compute is involved; the
multi-GPU exchange logic runs over gloo with world_size 2."""
import inspect
import os
import sys
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_exported():
    import ctypes

    from runia_core_b200 import _lib

    hdr = open(os.path.join(ROOT, "include", "runia_b200.h")).read()
    declared = set(re.findall(r"\b(runia_\w+)\s*\(", hdr))
    assert len(declared) >= 18
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/runia_b200.h but not exported"
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    assert lib.runia_b200_abi_version() == 1
    assert _lib.launch_count() >= 0


def test_no_oracle_import_in_product():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "runia_core_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), f


def test_registry_and_signatures():
    from runia_core_b200 import inference as I

    keys = {"KDE", "MD", "cMD", "KNN", "GMM", "energy", "msp", "gen", "ddu", "knn", "mahalanobis", "vim",
            "ash", "dice", "react", "dice_react"}
    assert set(I.postprocessors_dict) == keys and set(I.postprocessor_input_dict) == keys
    assert I.postprocessor_input_dict["vim"] == ["features", "logits"]
    assert I.postprocessor_input_dict["MD"] == ["latent_space_means"]
    assert I.LaREMPostprocessor is I.MDLatentSpace and I.LaREDPostprocessor is I.KDELatentSpace
    for cls in ("MDLatentSpace", "KDELatentSpace", "cMDLatentSpace", "KNNLatentSpace", "GMMLatentSpace"):
        assert cls in I.__all__
        assert list(inspect.signature(getattr(I, cls).__init__).parameters)[1:] == ["cfg"]
    sig = lambda c: list(inspect.signature(c.__init__).parameters)[1:]  # noqa: E731
    assert sig(I.GEN) == ["flip_sign", "gamma", "num_classes", "cfg"]
    assert sig(I.KNN) == ["flip_sign", "k_neighbors", "cfg"]
    assert sig(I.DICEReAct) == ["flip_sign", "dice_percentile", "react_percentile", "num_classes", "cfg"]
    assert inspect.signature(I.ASH.__init__).parameters["ash_percentile"].default == 85

    class Cfg(dict):
        def __getattr__(self, k):
            try:
                return self[k]
            except KeyError:
                raise AttributeError(k)

    assert I.KNNLatentSpace().K == 50 and I.KNNLatentSpace(Cfg(k_neighbors=20)).K == 20
    assert I.cMDLatentSpace().num_classes == 10 and I.cMDLatentSpace(Cfg(num_classes=5)).num_classes == 5
    with pytest.raises(AssertionError):
        I.register_postprocessor("x", ["bogus"])(object)


def test_ood_postprocessor_boundary():
    from runia_core_b200.inference import Energy, get_baselines_thresholds

    e = Energy(flip_sign=True)
    assert e.flip_sign and e.threshold is None and not e._setup_flag
    s = np.array([1.0, 2.0, 4.0])
    assert np.array_equal(e.flip_sign_fn(s), -s)
    d = e.flip_sign_fn({"a": s.copy()})
    assert np.array_equal(d["a"], -s)
    with pytest.raises(ValueError, match="scores must be a dict or ndarray"):
        e.flip_sign_fn([1.0])
    e.set_threshold(s)
    assert e._setup_flag and abs(e.threshold - (s.mean() - 1.645 * s.std())) < 1e-12
    th = get_baselines_thresholds(["raw", "m"], {"m": s})
    assert th["raw"] == 0.0 and abs(th["m"] - e.threshold) < 1e-12
    with pytest.raises(AssertionError, match="setup\\(\\) must be called before postprocess\\(\\)"):
        Energy(flip_sign=False).postprocess(np.zeros((2, 3), np.float32))


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_fails_loudly_without_cuda():
    from runia_core_b200.inference import MDLatentSpace

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        MDLatentSpace().setup(np.random.rand(10, 4))


def test_install_as_runia_core():
    import sys

    import runia_core_b200 as R

    saved = {k: v for k, v in sys.modules.items() if k == "runia_core" or k.startswith("runia_core.")}
    try:
        R.install_as_runia_core()
        from runia_core.evaluation import get_dl_h_z  # noqa: F401
        from runia_core.inference.postprocessors import postprocessors_dict  # noqa: F401
        import runia_core

        assert runia_core.apply_pca_ds_split is R.apply_pca_ds_split
    finally:
        for k in [k for k in sys.modules if k == "runia_core" or k.startswith("runia_core.")]:
            del sys.modules[k]
        sys.modules.update(saved)


def test_row_shard_partition():
    from runia_core_b200.sharding import row_shard

    for n in (0, 1, 7, 10_000, 33_554_432):
        for world in (1, 2, 3, 8):
            spans = [row_shard(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _gloo_worker(rank, world, port, ret):
    import torch.distributed as dist

    from oracle import oracle_np as O
    from runia_core_b200 import sharding as S

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        rng = np.random.RandomState(0)
        bank = O.normalize_rows_exact(rng.randn(301, 16).astype(np.float32))
        bank[50:60] = bank[50]  # ties across the shard boundary are broken by global index
        q = O.normalize_rows_exact(rng.randn(17, 16).astype(np.float32))
        k = 12
        lo, hi = S.row_shard(len(bank), rank, world)

        class Shard:
            pass

        sh = Shard()
        sh.bank, sh.idx_offset = bank[lo:hi], lo

        def search_fn(qn, b, kk):
            B = b.bank.astype(np.float64)
            d = np.stack([O.seq32_tree_sum((B - x.astype(np.float64)) ** 2) for x in qn.numpy()])
            order = np.stack([np.lexsort((np.arange(d.shape[1]), row))[:kk] for row in d])
            return (torch.from_numpy(np.take_along_axis(d, order, 1)), torch.from_numpy(order + b.idx_offset))

        d, i, kth = S.knn_search_sharded(torch.from_numpy(q), sh, k, search_fn=search_fn,
                                         merge_fn=S.merge_topk_reference)
        D, I = O.flat_l2_search_tree(bank, q, k)
        ok = np.array_equal(i.numpy(), I) and np.array_equal(d.numpy(), D) and np.array_equal(kth.numpy(), D[:, -1])

        # KDE: partial (max, sum) per shard -> MAX / SUM all-reduce
        class KShard:
            pass

        ks = KShard()
        x = rng.randn(200, 8)
        qq = rng.randn(9, 8)
        ks.bank, ks.bandwidth, ks.n_total = torch.from_numpy(x[slice(*S.row_shard(200, rank, world))]), 1.0, 200

        def partial_fn(qv, kb):
            t = -0.5 * ((qv[:, None, :] - kb.bank[None]) ** 2).sum(-1)
            m = t.max(1).values
            return m, torch.exp(t - m[:, None]).sum(1)

        got = S.kde_score_sharded(torch.from_numpy(qq), ks, partial_fn=partial_fn).numpy()
        ok = ok and np.allclose(got, O.kde_score(qq, x), rtol=1e-12, atol=1e-12)

        loc = torch.arange(*S.row_shard(11, rank, world), dtype=torch.float64)
        ok = ok and torch.equal(S.gather_rows(loc, 11), torch.arange(11, dtype=torch.float64))

        # setup() statistics over row shards: count-weighted combination of the per-rank class means
        feats = rng.randn(101, 5).astype(np.float32)
        labels = rng.randint(0, 4, 101)
        labels[labels == 2] = 1                      # class 2 empty everywhere
        labels[:51][labels[:51] == 3] = 0            # class 3 lives on rank 1 only (world 2)
        flo, fhi = S.row_shard(101, rank, world)
        xs, ls = feats[flo:fhi], labels[flo:fhi]
        with np.errstate(all="ignore"):
            lm = np.stack([xs[ls == c].mean(0) if (ls == c).any() else np.full(5, np.nan, np.float32) for c in range(4)])
        lc = np.array([(ls == c).sum() for c in range(4)])
        gm, total = S.combine_class_means(torch.from_numpy(lm.astype(np.float32)), torch.from_numpy(lc))
        ref = np.stack([feats[labels == c].astype(np.float64).mean(0) if (labels == c).any() else np.full(5, np.nan)
                        for c in range(4)])
        ok = ok and np.array_equal(total.numpy(), [(labels == c).sum() for c in range(4)])
        ok = ok and np.allclose(gm.numpy(), ref, rtol=0, atol=1e-6, equal_nan=True) and np.isnan(gm.numpy()[2]).all()
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_sharded_exchange_gloo_world2():
    import socket

    import torch.multiprocessing as mp

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ret = mp.Manager().dict()
    mp.spawn(_gloo_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret[0] and ret[1]


def test_entropy_sorting_network_sorts_every_input():
    """The 16-key comparator network of csrc/entropy.cu (sort16) is checked with the 0-1 principle:
    a comparator network sorts every input iff it sorts all 2^16 binary inputs."""
    src = open(os.path.join(ROOT, "runia_core_b200", "csrc", "entropy.cu")).read()
    body = src[src.index("void sort16("):]
    body = body[:body.index("#undef RUNIA_CE")]
    ces = [(int(a), int(b)) for a, b in re.findall(r"RUNIA_CE\((\d+),\s*(\d+)\)", body)]
    assert len(ces) == 60 and all(0 <= a < b < 16 for a, b in ces)
    x = ((np.arange(1 << 16)[:, None] >> np.arange(16)) & 1).astype(np.int8)
    for a, b in ces:
        lo, hi = np.minimum(x[:, a], x[:, b]), np.maximum(x[:, a], x[:, b])
        x[:, a], x[:, b] = lo, hi
    assert (np.diff(x, axis=1) >= 0).all()


def test_cabi_argument_errors_without_gpu():
    """Argument validation happens before any CUDA call, so the error contract of
    include/runia_b200.h (0 / RUNIA_E_BADARG / RUNIA_E_UNSUPPORTED + last_error text) and the pure
    host-side workspace queries can be checked on a box without a GPU."""
    import ctypes

    from runia_core_b200 import _lib

    raw = _lib.raw
    fake = ctypes.c_void_p(0x1000)  # never dereferenced: every call below returns before a launch
    err = lambda: raw("runia_b200_last_error")().decode()  # noqa: E731
    # empty inputs are a no-op
    assert raw("runia_mcd_entropy_f32")(None, 0, 16, 64, 5, 1e-5, 0.0, None, None, None) == 0
    assert raw("runia_rownorm_score_f32")(None, 0, 8, None, None, 8, None, 0, None, 0, 0.0, None, None, None) == 0
    assert raw("runia_logit_scores_f32")(None, 0, 10, 0.1, 10, None, None, None, None) == 0
    # entropy: n_mc outside [2, 128] is unsupported, k >= n_mc is a bad argument
    assert raw("runia_mcd_entropy_f32")(fake, 4, 129, 64, 5, 1e-5, 0.0, fake, None, None) == -2
    assert "n_mc=129" in err()
    assert raw("runia_mcd_entropy_f32")(fake, 4, 4, 64, 4, 1e-5, 0.0, fake, None, None) == -1
    assert raw("runia_mcd_entropy_f32")(None, 4, 16, 64, 5, 1e-5, 0.0, None, None, None) == -1
    # kNN: k outside [1, 1016] unsupported; empty bank bad argument; workspace query is host-only
    assert raw("runia_knn_search_f32")(fake, 4, fake, fake, None, None, 100, 8, 0, 0, None, None, None, None, fake,
                                       fake, 1 << 20, None) == -2
    assert raw("runia_knn_search_f32")(fake, 4, fake, fake, None, None, 0, 8, 5, 0, None, None, None, None, fake,
                                       fake, 1 << 20, None) == -1
    ws = raw("runia_knn_workspace_bytes")(10_000, 50_000, 512, 50)
    assert ws > 10_000 * 4 and raw("runia_knn_workspace_bytes")(10, 10, 8, 1017) == 0
    assert raw("runia_knn_workspace_bytes")(10, 10, 8, 500) > 0
    assert raw("runia_knn_search_f32")(fake, 4, fake, fake, None, None, 100, 8, 1017, 0, None, None, None, None, fake,
                                       fake, 1 << 20, None) == -2
    assert raw("runia_knn_search_f32")(fake, 10_000, fake, fake, None, None, 50_000, 512, 50, 0, None, None, None, None,
                                       fake, fake, 1024, None) == -3
    assert "workspace" in err()
    assert raw("runia_kde_workspace_bytes")(1000, 5000) > 0
    # tensor-core entry points refuse shapes they are not built for (K % 4 != 0) instead of mis-computing
    assert raw("runia_rownorm_score_tc")(fake, 4, 7, None, fake, fake, 7, None, 0, None, 0, 0.0, fake, None, None) == -2
    assert raw("runia_pca_transform_tc")(fake, 4, 10, None, fake, fake, 4, None, fake, None) == -2
    # ViM mode needs logits; unknown mode is a bad argument
    assert raw("runia_rownorm_score_f32")(fake, 4, 8, None, fake, 8, None, 1, None, 0, 1.0, None, fake, None) == -1
    assert raw("runia_rownorm_score_f32")(fake, 4, 8, None, fake, 8, None, 7, None, 0, 1.0, None, fake, None) == -1
    # linear heads: the FUSED ASH kernel needs the head in shared memory (wider heads prune with runia_ash_prune_f32
    # and use the general head); ASH keep count outside [1, d]
    assert raw("runia_ash_linear_lse_f32")(fake, 4, 512, fake, fake, 65, 77, fake, None) == -2
    assert "runia_ash_prune_f32" in err()
    assert raw("runia_ash_linear_lse_f32")(fake, 4, 16, fake, fake, 4, 17, fake, None) == -1
    assert raw("runia_ash_prune_f32")(fake, 4, 16, 0, fake, None) == -1
    assert raw("runia_clip_linear_lse_tc")(fake, 4, 510, fake, fake, fake, 1000, 1.0, fake, None) == -2  # d % 4 != 0
    assert raw("runia_topk_merge")(fake, fake, 65, 4, 5, None, None, None, None) == -1


def test_widening_boundary_signatures_and_argument_checks():
    """The functions added along the path keep the reference's names, parameters and assertion texts
    (metrics.py:37-42, llm_uncertainty/scores.py:49, inference/funcs.py:430-446, feature_extraction/utils.py:70);
    their argument checks fire before any device work."""
    from runia_core_b200.evaluation import get_auroc_results
    from runia_core_b200.feature_extraction import get_mean_or_fullmean_ls_sample
    from runia_core_b200.inference.funcs import get_predictive_uncertainty_score
    from runia_core_b200.llm_uncertainty import eigen_score

    assert list(inspect.signature(get_auroc_results).parameters) == [
        "detect_exp_name", "ind_samples_scores", "ood_samples_scores", "return_results_for_mlflow"]
    assert list(inspect.signature(eigen_score).parameters) == ["hidden_states", "alpha"]
    assert inspect.signature(eigen_score).parameters["alpha"].default == 1e-3
    assert list(inspect.signature(get_predictive_uncertainty_score).parameters) == ["input_samples", "mcd_nro_samples"]
    assert list(inspect.signature(get_mean_or_fullmean_ls_sample).parameters) == ["latent_sample", "method"]
    with pytest.raises(AssertionError, match="divisible by the mcd_nro_samples"):
        get_predictive_uncertainty_score(torch.zeros(7, 3), 2)
    with pytest.raises(AssertionError):
        get_mean_or_fullmean_ls_sample(torch.zeros(2, 3, 4, 5), method="median")


def test_mc_sampler_module_seed_stream():
    """MCSamplerModule (feature_extraction/abstract_classes.py:32-101): the seeds are the n_mc sequential
    `torch.rand(B, H, W) < drop_prob / block_size**2` draws of the DropBlock2D layers (drawn here with ONE
    `torch.rand(n_mc, B, H, W)`: torch's CPU uniform_ fills serially from the generator, so the stream is the same);
    eval mode is the identity (host-only paths, no CUDA needed)."""
    import torch

    from runia_core_b200.feature_extraction import MCSamplerModule

    for shape in ((1, 6, 5, 4), (3, 2, 7, 7), (2, 4, 1, 1), (1, 3, 14, 3)):
        for n_mc, bs, p in ((5, 2, 0.4), (16, 3, 0.3), (32, 1, 0.9)):
            smp = MCSamplerModule(mc_samples=n_mc, block_size=bs, drop_prob=p, layer_type="FC").train()
            x = torch.randn(*shape)
            torch.manual_seed(3)
            seeds = smp.draw_seeds(x)
            torch.manual_seed(3)
            ref = torch.stack([(torch.rand(shape[0], *shape[2:]) < p / bs**2) for _ in range(n_mc)]).to(torch.uint8)
            assert seeds.shape == (n_mc, shape[0], *shape[2:]) and seeds.dtype == torch.uint8
            assert torch.equal(seeds, ref)
    with pytest.raises(AssertionError):
        MCSamplerModule(mc_samples=2, block_size=1, drop_prob=0.1, layer_type="Linear")
    smp = MCSamplerModule(mc_samples=5, block_size=2, drop_prob=0.4, layer_type="FC").eval()
    x = torch.randn(1, 6, 5, 4)
    assert torch.equal(smp(x), x.reshape(1, -1).repeat(5, 1))


@pytest.mark.skipif(not os.path.exists("/root/reference/runia_core/evaluation/baselines.py"),
                    reason="the reference tree only exists in the build container")
def test_reference_callers_bind_to_the_product_unchanged():
    """the reference's own driver module must keep working on top of the drop-in.  The
    UNMODIFIED source of `runia_core/evaluation/baselines.py` is executed with `runia_core` aliased to this package
    (`install_as_runia_core()`): its `from runia_core.inference.postprocessors import DICE, ReAct, ...` must resolve to
    the product's classes, and every call it makes (constructor keywords, setup / postprocess keywords) must be
    accepted by their signatures.  (Its numerical output on the product is checked on the GPU against the reference's
    own output: tests/test_gpu_shapes.py::test_calculate_all_baselines_matches_reference_driver.)"""
    import ast
    import importlib.util
    import types

    import runia_core_b200 as R

    saved = {k: v for k, v in sys.modules.items() if k == "runia_core" or k.startswith("runia_core.") or k == "omegaconf"}
    try:
        R.install_as_runia_core()
        if "omegaconf" not in sys.modules:
            sys.modules["omegaconf"] = types.SimpleNamespace(DictConfig=dict)
        path = "/root/reference/runia_core/evaluation/baselines.py"
        spec = importlib.util.spec_from_file_location("reference_baselines_on_product", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        P = R.inference.postprocessors
        for name in ("DICE", "ReAct", "ASH", "GEN", "ViM", "MSP", "Energy", "Mahalanobis", "KNN", "DDU", "DICEReAct"):
            assert getattr(mod, name) is getattr(P, name), name
        # every keyword the reference passes to a constructor / setup / postprocess exists in the product's signature
        tree = ast.parse(open(path).read())
        ctor_kw = {}
        for node in ast.walk(tree):
            if isinstance(node, ast.Call) and isinstance(node.func, ast.Name) and hasattr(P, node.func.id) and \
                    node.func.id[0].isupper():
                ctor_kw.setdefault(node.func.id, set()).update(k.arg for k in node.keywords)
        assert len(ctor_kw) == 11
        for cls, kws in ctor_kw.items():
            params = set(inspect.signature(getattr(P, cls).__init__).parameters)
            assert kws <= params, (cls, kws - params)
        for cls in ctor_kw:
            sig = inspect.signature(getattr(P, cls).setup).parameters
            assert "ind_train_data" in sig and any(p.kind == p.VAR_KEYWORD for p in sig.values()), cls
            sig = inspect.signature(getattr(P, cls).postprocess).parameters
            assert "test_data" in sig and any(p.kind == p.VAR_KEYWORD for p in sig.values()), cls
        # and the product ships the same driver entry points
        from runia_core_b200.evaluation import baselines as mine

        for fn in mod.__all__:
            if fn != "baseline_name_dict":
                assert list(inspect.signature(getattr(mine, fn)).parameters) == \
                    list(inspect.signature(getattr(mod, fn)).parameters), fn
    finally:
        for k in [k for k in sys.modules if k == "runia_core" or k.startswith("runia_core.")]:
            del sys.modules[k]
        sys.modules.pop("omegaconf", None)
        sys.modules.update(saved)
