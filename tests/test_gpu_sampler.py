"""(f3) MC-DropBlock sampler + spatial mean (csrc/sampler.cu) and the online LaREx chain.

The torch reference below is the published forward of dropblock==0.3.0's DropBlock2D (the layer
MCSamplerModule instantiates, feature_extraction/abstract_classes.py:72-79) followed by the reference's
"fullmean" reducer (feature_extraction/utils.py:70-92), run on the same Bernoulli seeds."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _dropblock_fullmean(x, seed, bs):
    """x [B, C, H, W], seed [n_mc, B, H, W] -> [B * n_mc, C]: n_mc DropBlock2D passes, each per image (B = 1 calls)."""
    rows = []
    for b in range(x.shape[0]):
        for m in range(seed.shape[0]):
            mask = seed[m, b:b + 1].float()
            bm = F.max_pool2d(mask[:, None], kernel_size=(bs, bs), stride=(1, 1), padding=bs // 2)
            if bs % 2 == 0:
                bm = bm[:, :, :-1, :-1]
            bm = 1 - bm.squeeze(1)
            out = x[b:b + 1] * bm[:, None, :, :]
            out = out * bm.numel() / bm.sum()
            out = torch.mean(torch.mean(out, dim=3, keepdim=True), dim=2, keepdim=True)
            rows.append(out.reshape(1, -1))
    return torch.cat(rows)


@pytest.mark.parametrize("B,C,H,W,n_mc,bs,p", [(1, 512, 7, 7, 16, 3, 0.3), (5, 100, 16, 12, 16, 4, 0.4), (3, 33, 1, 1, 5, 1, 0.5),
                                               (2, 64, 24, 24, 32, 7, 0.5), (2, 40, 9, 9, 8, 6, 0.9), (64, 512, 7, 7, 16, 3, 0.3), (3, 48, 7, 7, 40, 3, 0.3)])
def test_mc_dropblock_mean_matches_dropblock2d(B, C, H, W, n_mc, bs, p):
    from runia_core_b200 import _ops

    g = torch.Generator().manual_seed(B * 1000 + C)
    x = (torch.randn(B, C, H, W, generator=g) + 0.5).cuda()
    seed = (torch.rand(n_mc, B, H, W, generator=g) < p / bs**2).to(torch.uint8).cuda()
    got = _ops.mc_dropblock_mean(x, seed, bs)
    ref = _dropblock_fullmean(x, seed, bs)
    assert got.shape == (B * n_mc, C)
    fin = torch.isfinite(ref)
    assert torch.equal(fin, torch.isfinite(got))  # every cell dropped -> 0 / 0 in both
    scale = ref[fin].abs().mean()
    assert (got[fin] - ref[fin]).abs().max() <= 2e-6 * scale


def test_sampler_module_reproduces_rng_stream_and_layout():
    from runia_core_b200.feature_extraction import MCSamplerModule

    n_mc, bs, p = 16, 3, 0.3
    smp = MCSamplerModule(mc_samples=n_mc, block_size=bs, drop_prob=p, layer_type="Conv").train()
    x = torch.randn(1, 512, 7, 7).cuda()
    torch.manual_seed(11)
    got = smp(x)
    torch.manual_seed(11)  # what n_mc DropBlock2D layers draw, in order
    seed = torch.stack([(torch.rand(1, 7, 7) < p / bs**2) for _ in range(n_mc)]).to(torch.uint8).cuda()
    ref = _dropblock_fullmean(x, seed, bs)
    assert got.shape == (n_mc, 512) and got.is_cuda
    assert (got - ref).abs().max() <= 2e-6 * ref.abs().mean()
    smp.eval()  # DropBlock2D is the identity outside training
    ev = smp(x)
    assert torch.allclose(ev, x.mean((2, 3)).expand(n_mc, -1), atol=1e-6)


def test_larex_online_chain_matches_staged_api():
    """LaRExInference.get_score (device chain) == get_dl_h_z -> apply_pca_transform -> MD.postprocess on the same samples."""
    import runia_core_b200 as R
    from runia_core_b200.feature_extraction import MCSamplerModule

    class Hook:
        output = None

    class Net(torch.nn.Module):
        def __init__(self, hook):
            super().__init__()
            self.conv = torch.nn.Conv2d(3, 64, 3, padding=1)
            self.head = torch.nn.Linear(64, 10)
            self.hook = hook

        def forward(self, x):
            z = torch.relu(self.conv(x))
            self.hook.output = z
            return self.head(z.mean((2, 3)))

    torch.manual_seed(0)
    hook, n_mc = Hook(), 16
    net = Net(hook).cuda().eval()
    smp = MCSamplerModule(mc_samples=n_mc, block_size=3, drop_prob=0.3).train()
    imgs = torch.randn(96, 3, 16, 16)
    with torch.no_grad():
        net(imgs.cuda())
    train_rows = smp.sample_batch(hook.output)                       # [96 * 16, 64]
    _, h_z = R.evaluation.get_dl_h_z(train_rows, n_mc)
    np.random.seed(1)
    z_train, pca = R.apply_pca_ds_split(h_z, nro_components=8)
    md = R.inference.postprocessors_dict["MD"]()
    md.setup(z_train)
    inf = R.inference.LaRExInference(net, md, drop_block_prob=0.3, drop_block_size=3, mcd_samples_nro=n_mc,
                                     mcd_sampler=MCSamplerModule, pca_transform=pca)
    torch.manual_seed(5)
    out, score = inf.get_score(imgs[:1], hook)
    torch.manual_seed(5)
    rows = inf.mc_sampler(hook.output)
    _, hz1 = R.evaluation.get_dl_h_z(rows, n_mc)
    staged = md.postprocess(R.apply_pca_transform(hz1, pca))
    assert out.shape == (1, 10) and score.shape == (1,) and score.dtype == np.float64
    np.testing.assert_allclose(score, staged, rtol=1e-5)
    # the CUDA-graph replay (default) and the launch-by-launch chain give the same bits on the same seeds, image
    # after image (static buffers are refilled, not re-captured), for a second map shape too
    assert inf.use_cuda_graph and len(inf._graphs) == 1
    for i in (1, 2, 3):
        torch.manual_seed(40 + i)
        _, s_graph = inf.get_score(imgs[i:i + 1], hook)
        inf.use_cuda_graph = False
        torch.manual_seed(40 + i)
        _, s_plain = inf.get_score(imgs[i:i + 1], hook)
        inf.use_cuda_graph = True
        assert np.array_equal(s_graph, s_plain)
    torch.manual_seed(50)
    _, s4 = inf.get_score(imgs[:4], hook)  # a batch of four maps: its own captured graph
    assert s4.shape == (4,) and len(inf._graphs) == 2
    res, secs = inf.test_time_inference(imgs[:1], hook)
    assert res[1].shape == (1,) and secs > 0
    # LaRD (no MC sampling, no entropy): mean map -> PCA -> LaREM on the device == the staged public calls
    lard = R.inference.LaRDInference(net, md, pca_transform=pca, layer_type="Conv")
    out_d, sc_d = lard.get_score(imgs[:1], hook)
    z = hook.output.mean((2, 3)).cpu().numpy().astype(np.float64)
    np.testing.assert_allclose(sc_d, md.postprocess(R.apply_pca_transform(z, pca)), rtol=1e-4)


def test_folded_pca_larem_matches_staged():
    """FoldedLaREM (one contraction over the raw latents) == apply_pca_transform -> MDLatentSpace.postprocess."""
    import runia_core_b200 as R

    rng = np.random.RandomState(3)
    D0, d = 512, 256
    mix = np.eye(D0) + 0.2 * rng.standard_normal((D0, D0)) / np.sqrt(D0)
    train = (0.5 + rng.standard_normal((20_000, D0)) @ mix).astype(np.float32)
    test = (-0.2 + rng.standard_normal((30_000, D0)) @ mix).astype(np.float32)
    np.random.seed(1)
    z_train, pca = R.apply_pca_ds_split(train, nro_components=d)
    md = R.inference.postprocessors_dict["MD"]()
    md.setup(z_train)
    staged = md.postprocess(R.apply_pca_transform(test, pca))
    folded = R.inference.FoldedLaREM(pca, md)
    got = folded.postprocess(test)
    assert got.dtype == np.float64 and got.shape == staged.shape
    np.testing.assert_allclose(got, staged, rtol=2e-5)
    md2 = R.inference.postprocessors_dict["MD"]()
    md2.setup((z_train + 3.0).astype(np.float32))  # a LaREM centre away from the PCA origin
    np.testing.assert_allclose(R.inference.FoldedLaREM(pca, md2).postprocess(test),
                               md2.postprocess(R.apply_pca_transform(test, pca)), rtol=2e-5)


@pytest.mark.parametrize("shape,bs", [((1, 6, 5, 4), 2), ((3, 8, 7, 7), 3), ((2, 5, 9, 6), 4)])
def test_mc_sampler_fc_layer_masked_maps(shape, bs):
    """layer_type "FC" / "RPN" (abstract_classes.py:81-101 without the reduction): the DropBlock2D-masked maps
    themselves, normalised over the whole batch mask, == the published DropBlock2D forward on the same seeds."""
    import torch.nn.functional as F

    from runia_core_b200.feature_extraction import MCSamplerModule

    n_mc, p = 7, 0.4
    smp = MCSamplerModule(mc_samples=n_mc, block_size=bs, drop_prob=p, layer_type="FC").train()
    x = torch.randn(*shape)
    torch.manual_seed(3)
    got = smp(x.cuda()).cpu()
    torch.manual_seed(3)
    rows = []
    for _ in range(n_mc):
        mask = (torch.rand(shape[0], *shape[2:]) < p / bs**2).float()
        bm = F.max_pool2d(mask[:, None], kernel_size=(bs, bs), stride=(1, 1), padding=bs // 2)
        if bs % 2 == 0:
            bm = bm[:, :, :-1, :-1]
        bm = 1 - bm.squeeze(1)
        rows.append((x * bm[:, None] * bm.numel() / bm.sum()).reshape(1, -1))
    ref = torch.cat(rows)
    assert got.shape == ref.shape
    torch.testing.assert_close(got, ref, rtol=1e-6, atol=1e-6, equal_nan=True)


@pytest.mark.parametrize("P,sr,C,H,W", [(7, 2, 64, 25, 34), (14, 2, 16, 50, 68), (7, 0, 8, 13, 17), (5, 4, 33, 10, 10)])
def test_roi_align_and_object_means_vs_torchvision(P, sr, C, H, W):
    """object_level.py:254-309: torchvision.ops.roi_align(aligned=True) + per-object mean / std over the RoI, and the
    RoI maps themselves; boxes that leave the image, degenerate boxes, adaptive sampling (sampling_ratio 0)."""
    from torchvision.ops import roi_align as tv_roi_align

    from runia_core_b200 import _ops
    from runia_core_b200.feature_extraction.object_level import _reduce_features_to_rois

    rng = np.random.RandomState(P + C)
    img_shape = (H * 16, W * 16)
    feat = torch.from_numpy(rng.randn(1, C, H, W).astype(np.float32))
    K = 23
    x1 = rng.rand(K) * img_shape[1] * 0.8
    y1 = rng.rand(K) * img_shape[0] * 0.8
    boxes = np.stack([x1, y1, x1 + 8 + rng.rand(K) * img_shape[1] * 0.5, y1 + 8 + rng.rand(K) * img_shape[0] * 0.5], 1)
    boxes[0] = [-30.0, -20.0, 40.0, 50.0]                       # leaves the image on the top left
    boxes[1] = [img_shape[1] - 10.0, img_shape[0] - 10.0, img_shape[1] + 90.0, img_shape[0] + 60.0]
    boxes[2] = [100.0, 100.0, 100.0, 100.0]                     # zero area
    boxes = torch.from_numpy(boxes.astype(np.float32))
    scale = W / img_shape[1]
    ref = tv_roi_align(feat, [boxes], output_size=P, spatial_scale=scale, sampling_ratio=sr, aligned=True)
    got = _ops.roi_align(feat, boxes, P, scale, sr, aligned=True).cpu()
    # same operation order as torchvision's CPU kernel, products and sums rounded one by one (no FMA contraction)
    torch.testing.assert_close(got, ref, rtol=1e-5, atol=2e-6)
    means, stds = _reduce_features_to_rois([feat.cuda(), (2 * feat).cuda()], (P, P), boxes.cuda(), img_shape, sr, 2, K,
                                           return_stds=True)
    assert len(means) == K and means[0].shape == (1, 2 * C)
    ref_m = torch.cat([ref.mean((2, 3)), 2 * ref.mean((2, 3))], 1)
    ref_s = torch.cat([ref.std((2, 3)), 2 * ref.std((2, 3))], 1)
    torch.testing.assert_close(torch.cat(means).cpu(), ref_m, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(torch.cat(stds).cpu(), ref_s, rtol=1e-4, atol=1e-5)
    ref_na = tv_roi_align(feat, [boxes], output_size=(P, P + 1), spatial_scale=scale, sampling_ratio=sr, aligned=False)
    got_na = _ops.roi_align(feat, boxes, (P, P + 1), scale, sr, aligned=False).cpu()
    torch.testing.assert_close(got_na, ref_na, rtol=1e-5, atol=1e-5)


def test_dropblock_rois_entropy_chain():
    """object_level.py:312-366: RoIAlign -> MC-DropBlock per detection -> get_dl_h_z, against the staged public
    pieces (torchvision RoIs, MCSamplerModule per detection, get_dl_h_z) on the same seeds."""
    from torchvision.ops import roi_align as tv_roi_align

    import runia_core_b200 as R
    from runia_core_b200.feature_extraction import MCSamplerModule
    from runia_core_b200.feature_extraction.object_level import _dropblock_rois_get_entropy

    rng = np.random.RandomState(2)
    C, H, W, P, K, n_mc = 32, 20, 30, 7, 9, 16
    img_shape = (H * 8, W * 8)
    feat = torch.from_numpy(rng.randn(1, C, H, W).astype(np.float32))
    x1, y1 = rng.rand(K) * 100, rng.rand(K) * 60
    boxes = torch.from_numpy(np.stack([x1, y1, x1 + 40 + 60 * rng.rand(K), y1 + 30 + 50 * rng.rand(K)], 1).astype(np.float32))
    smp = MCSamplerModule(mc_samples=n_mc, block_size=3, drop_prob=0.3).train()
    torch.manual_seed(9)
    got = _dropblock_rois_get_entropy([feat.cuda()], (P,), boxes.cuda(), img_shape, 2, 1, n_mc, smp)
    rois = tv_roi_align(feat, [boxes], output_size=P, spatial_scale=W / img_shape[1], sampling_ratio=2, aligned=True)
    torch.manual_seed(9)
    rows = torch.cat([smp(det.unsqueeze(0).cuda()) for det in rois])  # upstream's loop: detection after detection
    _, hz = R.evaluation.get_dl_h_z(rows, n_mc)
    assert got.shape == (K, C) and got.dtype == torch.float32
    np.testing.assert_allclose(got.numpy(), hz, rtol=1e-4, atol=1e-4)
