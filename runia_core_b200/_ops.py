"""Device-level operators: thin Python wrappers that allocate outputs (torch) and call the C ABI.
Inputs and outputs are CUDA tensors; fitted state lives in small dataclasses that hold the
device copies of what `setup()` computed on the host.  The reference-facing classes in
`inference/postprocessors.py` and the free functions in `evaluation/entropy.py` /
`dimensionality_reduction.py` are built on these; `bench.py` times them directly for the
device-resident number."""
import contextlib
import os
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from . import _lib
from ._device import _host_array, as_f32_rows, device, ptr, stream_ptr, stream_rows, to_device

FLT_MAX = float(np.finfo(np.float32).max)

# Contraction engine: "tc" = tcgen05 tensor cores, 3xTF32 operand split, FP32 accumulation in TMEM
# (default); "simt" = FP32 FFMA kernels.  Both are FP32-faithful; the choice is by measured error
# and speed.  The tensor-core kernels need K % 4 == 0; other shapes use "simt".
_ENGINE = os.environ.get("RUNIA_B200_ENGINE", "tc").lower()


def set_engine(name: str):
    global _ENGINE
    assert name in ("tc", "simt")
    _ENGINE = name


def get_engine() -> str:
    return _ENGINE


def split_tf32(w: torch.Tensor):
    """fp32 matrix -> (hi, lo) tf32 planes for the tensor-core kernels."""
    w = w.contiguous()
    hi, lo = torch.empty_like(w), torch.empty_like(w)
    _lib.call("runia_split_tf32", w.data_ptr(), w.numel(), hi.data_ptr(), lo.data_ptr(), stream_ptr())
    return hi, lo


TC_MAX_CLASSES = 16  # tc_kernels.cu kClassMax: per-class accumulators live in registers
TC_MAX_K = 4096  # tc_gemm.cuh kMaxK: the centre of the streamed operand is staged in shared memory


def _tc_ok(k: int) -> bool:
    return _ENGINE == "tc" and k % 4 == 0 and k <= TC_MAX_K


def _empty(shape, dtype):
    return torch.empty(shape, dtype=dtype, device=device())


# ------------------------------------------------------------------------------------------
# (a1) entropy
# ------------------------------------------------------------------------------------------
def entropy_k(n_mc: int) -> int:
    """evaluation/entropy.py:66"""
    return 5 if n_mc > 5 else n_mc - 1


def mcd_entropy(z: torch.Tensor, n_mc: int, k: Optional[int] = None, want_joint: bool = True,
                min_dist: float = 1e-5):
    """z: [n_items * n_mc, D] float32, CUDA tensor or host array / tensor.  Returns (h_mvn [n_items] f64 or None,
    h_z [n_items, D] f64) on the device."""
    from scipy.special import digamma

    if k is None:
        k = entropy_k(n_mc)
    n_items = z.shape[0] // n_mc
    D = z.shape[1]
    h_z = _empty((n_items, D), torch.float64)
    h_mvn = _empty((n_items,), torch.float64) if want_joint else None
    c_term = float(-digamma(k) + digamma(n_mc))

    def run(zc, lo, hi):  # items [lo, hi): zc is the device block of their n_mc * (hi - lo) sample rows
        _lib.call("runia_mcd_entropy_f32", zc.data_ptr(), hi - lo, n_mc, D, k, float(min_dist), c_term,
                  h_z.data_ptr() + lo * D * 8, None if h_mvn is None else h_mvn.data_ptr() + lo * 8, stream_ptr())

    if not (isinstance(z, torch.Tensor) and z.is_cuda):
        # host samples (what get_dl_h_z receives): whole items stream through the pinned ring, the kernel of one
        # chunk of items runs while the next chunk crosses PCIe
        pinned_ok = isinstance(z, torch.Tensor) and z.is_pinned() and z.dtype == torch.float32 and z.is_contiguous()
        zh = z if pinned_ok else np.ascontiguousarray(_host_array(z), np.float32)
        if stream_rows(zh[: n_items * n_mc].reshape(n_items, n_mc * D), run, min_rows=64):
            return h_mvn, h_z
        z = to_device(zh, torch.float32)
    z = z.to(torch.float32).contiguous()
    assert z.is_cuda and z.dim() == 2
    run(z, 0, n_items)
    return h_mvn, h_z


# ------------------------------------------------------------------------------------------
# (a2) PCA projection
# ------------------------------------------------------------------------------------------
@dataclass
class PCAState:
    mean_f32: Optional[torch.Tensor]   # [D0]
    mean_f64: Optional[torch.Tensor]   # [D0]
    components: torch.Tensor           # [d, D0] f32
    inv_scale: Optional[torch.Tensor]  # [d] f32 or None
    d: int
    D0: int
    planes: Optional[tuple] = None     # (hi, lo) tf32 planes of `components`


def pca_prepare(mean, components, explained_variance, whiten) -> PCAState:
    comp = np.ascontiguousarray(components, np.float64)
    inv = None
    if whiten:
        scale = np.sqrt(np.asarray(explained_variance, np.float64))
        eps = np.finfo(np.asarray(explained_variance).dtype).eps
        scale = np.where(scale < eps, eps, scale)
        inv = to_device((1.0 / scale).astype(np.float32))
    m64 = None if mean is None else to_device(np.asarray(mean, np.float64).reshape(-1))
    m32 = None if mean is None else m64.to(torch.float32)
    c32 = to_device(comp.astype(np.float32))
    return PCAState(m32, m64, c32, inv, comp.shape[0], comp.shape[1],
                    split_tf32(c32) if comp.shape[1] % 4 == 0 else None)


def pca_transform(x, st: PCAState) -> torch.Tensor:
    if x.ndim != 2 or x.shape[1] != st.D0:
        raise ValueError(f"expected rows of width {st.D0}, got an array of shape {tuple(x.shape)}")
    z = _empty((x.shape[0], st.d), torch.float32)

    def run(xc, lo, hi):
        xf, centered = as_f32_rows(xc, st.mean_f64)
        mean = None if centered else ptr(st.mean_f32)
        zp = z.data_ptr() + lo * st.d * 4
        if _tc_ok(st.D0) and st.planes is not None and st.d % 4 == 0:
            _lib.call("runia_pca_transform_tc", xf.data_ptr(), hi - lo, st.D0, mean, st.planes[0].data_ptr(),
                      st.planes[1].data_ptr(), st.d, ptr(st.inv_scale), zp, stream_ptr())
        else:
            _lib.call("runia_pca_transform_f32", xf.data_ptr(), hi - lo, st.D0, mean, st.components.data_ptr(), st.d,
                      ptr(st.inv_scale), zp, stream_ptr())

    if not stream_rows(x, run):
        run(x, 0, x.shape[0])
    return z


# ------------------------------------------------------------------------------------------
# (a3) LaREM Mahalanobis / (a6) class-conditional Mahalanobis: factor the precision once
# ------------------------------------------------------------------------------------------
_blas_ctl = None
_PINVH_FACTORS = {}  # id(precision ndarray) -> (that ndarray, eigenvalues, eigenvectors in columns, checksum); see pinvh()


@contextlib.contextmanager
def host_blas_single_thread():
    """Keeps the SMALL host linear algebra that remains in setup() (a [C, d] pseudo-inverse, a least-squares solve of
    the fold) on the calling thread.  A multi-threaded BLAS call leaves its worker threads busy-waiting for ~100 ms
    (OpenBLAS pthreads, measured on the B200 boxes: scripts/sweep_diag2.py); the postprocess() calls that follow a
    setup() then find the cores the staging engine's copy threads need taken, and a 0.6 ms call takes 1.3-6 ms."""
    global _blas_ctl
    try:
        from threadpoolctl import ThreadpoolController
    except ImportError:  # pragma: no cover
        yield
        return
    if _blas_ctl is None:
        _blas_ctl = ThreadpoolController()
    with _blas_ctl.limit(limits=1, user_api="blas"):
        yield


def mm64(a, b) -> np.ndarray:
    """a @ b in float64 on the device, NumPy in / NumPy out: the matrix products of the setup() fits (a plain library
    GEMM through torch -- not on the scoring path).  Products with a dimension below 32 stay on the calling thread."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    if min(a.shape + b.shape) < 32:
        with host_blas_single_thread():
            return a @ b
    return (to_device(np.ascontiguousarray(a), torch.float64) @ to_device(np.ascontiguousarray(b), torch.float64)).cpu().numpy()


def factor_precision(precision):
    """P (symmetric, float64) -> (Wt [r, d] float64, sign [r]) with P = sum_j sign_j w_j w_j^T.
    Eigenvalues below 1e-14 * max|lambda| are rounding residue of pinvh's rank cut and get
    sign 0."""
    cached = _PINVH_FACTORS.get(id(precision))
    P = np.asarray(precision, np.float64)
    P = 0.5 * (P + P.T)
    if cached is not None and cached[0] is precision and cached[3] == float(P.sum()):  # same object, not edited in place
        lam, V = cached[1], cached[2]  # this precision came out of pinvh(): its eigenpairs are already known
    elif P.shape[0] >= 64 and np.isfinite(P).all():
        lam, V = eigh(P)
    else:
        with host_blas_single_thread():
            lam, V = np.linalg.eigh(P)
    amax = np.abs(lam).max() if lam.size else 0.0
    sign = np.sign(lam)
    sign[np.abs(lam) <= 1e-14 * amax] = 0.0
    keep = sign != 0
    if not keep.any():
        keep[:1] = True
    Wt = (V[:, keep] * np.sqrt(np.abs(lam[keep]))).T
    return np.ascontiguousarray(Wt), sign[keep]


@dataclass
class MDState:
    mu_f32: torch.Tensor
    mu_f64: torch.Tensor
    Wt: torch.Tensor
    sign: Optional[torch.Tensor]
    d: int
    r: int
    planes: Optional[tuple] = None


def md_prepare(mean, precision) -> MDState:
    Wt, sign = factor_precision(precision)
    mu64 = to_device(np.asarray(mean, np.float64).reshape(-1))
    sg = None if np.all(sign == 1.0) else to_device(sign.astype(np.float32))
    w32 = to_device(Wt.astype(np.float32))
    return MDState(mu64.to(torch.float32), mu64, w32, sg, Wt.shape[1], Wt.shape[0],
                   split_tf32(w32) if Wt.shape[1] % 4 == 0 else None)


def md_fold_pca(pca_mean, components, explained_variance, whiten, md_mean, precision) -> Optional[MDState]:
    """LaREM on PCA-reduced latents as ONE contraction over the raw latents (2,056 B and 262,656 FLOP
    per embedding at 512 -> 256, no intermediate [N, d] array): with z = (x - m) A^T (A = components / scale) and
    P = sum_j s_j w_j w_j^T,  (z - mu)^T P (z - mu) = sum_j s_j (w_j^T A (x - m'))^2  where  W A (m' - m) = W mu.
    Returns None when that centre does not exist (W A without full row rank)."""
    comp = np.ascontiguousarray(components, np.float64)
    A = comp
    if whiten:
        scale = np.sqrt(np.asarray(explained_variance, np.float64))
        eps = np.finfo(np.asarray(explained_variance).dtype).eps
        A = comp / np.where(scale < eps, eps, scale)[:, None]
    Wt, sign = factor_precision(precision)
    Wf = mm64(Wt, A)                                              # [r, D0]
    with host_blas_single_thread():
        v = Wt @ np.asarray(md_mean, np.float64).reshape(-1)      # [r]
        delta = np.linalg.lstsq(Wf, v, rcond=None)[0]
        bad = np.abs(Wf @ delta - v).max() > 1e-9 * (1.0 + np.abs(v).max())
    if bad:
        return None
    m = (0.0 if pca_mean is None else np.asarray(pca_mean, np.float64).reshape(-1)) + delta
    mu64 = to_device(m)
    sg = None if np.all(sign == 1.0) else to_device(sign.astype(np.float32))
    w32 = to_device(Wf.astype(np.float32))
    return MDState(mu64.to(torch.float32), mu64, w32, sg, Wf.shape[1], Wf.shape[0],
                   split_tf32(w32) if Wf.shape[1] % 4 == 0 else None)


def _rownorm(xf, n, d, mu, W, planes, r, sign, mode, logits, C, alpha, o64, o32):
    if _tc_ok(d) and planes is not None:
        _lib.call("runia_rownorm_score_tc", xf.data_ptr(), n, d, mu, planes[0].data_ptr(), planes[1].data_ptr(), r,
                  sign, mode, logits, C, alpha, o64, o32, stream_ptr())
    else:
        _lib.call("runia_rownorm_score_f32", xf.data_ptr(), n, d, mu, W.data_ptr(), r, sign, mode, logits, C, alpha,
                  o64, o32, stream_ptr())


def md_score(x, st: MDState, out_dtype=torch.float64, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """LaREM scores of the rows of x (into `out` when given: a contiguous [N] device tensor of `out_dtype`).  Host
    matrices (what `MDLatentSpace.postprocess` receives, evaluation/metrics.py:331-340) stream through the pinned
    staging ring: the kernel scores chunk k while chunk k + 1 crosses PCIe and chunk k + 2 is copied into its slot."""
    if x.ndim != 2 or x.shape[1] != st.d:
        raise ValueError(f"expected rows of width {st.d}, got an array of shape {tuple(x.shape)}")
    if out is None:
        out = _empty((x.shape[0],), out_dtype)
    assert out.dtype == out_dtype and out.is_contiguous() and out.shape[0] == x.shape[0]
    esz = out.element_size()

    def run(xc, lo, hi):
        xf, centered = as_f32_rows(xc, st.mu_f64)
        o = out.data_ptr() + lo * esz
        _rownorm(xf, hi - lo, st.d, None if centered else st.mu_f32.data_ptr(), st.Wt, st.planes, st.r, ptr(st.sign),
                 _lib.ROWNORM_MD, None, 0, 0.0, o if out_dtype == torch.float64 else None,
                 o if out_dtype == torch.float32 else None)

    if not stream_rows(x, run):
        run(x, 0, x.shape[0])
    return out


@dataclass
class VimState:
    u_f32: torch.Tensor
    u_f64: torch.Tensor
    NSt: torch.Tensor  # [r, d]
    alpha: float
    d: int
    r: int
    planes: Optional[tuple] = None


def vim_prepare(u, NS, alpha) -> VimState:
    u64 = to_device(np.asarray(u, np.float64).reshape(-1))
    NSt = np.ascontiguousarray(np.asarray(NS, np.float64).T.astype(np.float32))
    n32 = to_device(NSt)
    return VimState(u64.to(torch.float32), u64, n32, float(alpha), NSt.shape[1], NSt.shape[0],
                    split_tf32(n32) if NSt.shape[1] % 4 == 0 else None)


def vim_score(x, logits, st: VimState) -> torch.Tensor:
    if x.ndim != 2 or x.shape[1] != st.d:
        raise ValueError(f"expected rows of width {st.d}, got an array of shape {tuple(x.shape)}")
    xf, centered = as_f32_rows(x, st.u_f64)
    lg = to_device(logits, torch.float32)
    n = xf.shape[0]
    out = _empty((n,), torch.float32)
    _rownorm(xf, n, st.d, None if centered else st.u_f32.data_ptr(), st.NSt, st.planes, st.r, None,
             _lib.ROWNORM_VIM, lg.data_ptr(), lg.shape[1], st.alpha, None, out.data_ptr())
    return out


def residual_norm(x, st: VimState) -> torch.Tensor:
    """||(x - u) NS||_2 per row (used by ViM.setup for alpha, postprocessors.py:1072)."""
    xf, centered = as_f32_rows(x, st.u_f64)
    n = xf.shape[0]
    out = _empty((n,), torch.float32)
    zero_logit = torch.zeros((n, 1), dtype=torch.float32, device=xf.device)  # logsumexp of one zero logit = 0
    _rownorm(xf, n, st.d, None if centered else st.u_f32.data_ptr(), st.NSt, st.planes, st.r, None,
             _lib.ROWNORM_VIM, zero_logit.data_ptr(), 1, -1.0, None, out.data_ptr())  # -alpha sqrt(.) with alpha = -1
    return out


@dataclass
class ClassCondState:
    g_f32: torch.Tensor
    g_f64: torch.Tensor
    Wt: torch.Tensor
    sign: Optional[torch.Tensor]
    Mc: torch.Tensor
    valid: torch.Tensor
    C: int
    d: int
    r: int
    planes: Optional[tuple] = None


def classcond_prepare(class_mean, precision) -> ClassCondState:
    cm = np.asarray(class_mean, np.float64)
    valid = ~np.isnan(cm).any(axis=1)
    Wt, sign = factor_precision(precision)
    g = cm[valid].mean(0) if valid.any() else np.zeros(cm.shape[1])
    Mc = np.zeros((cm.shape[0], Wt.shape[0]))
    Mc[valid] = mm64(cm[valid] - g, Wt.T)
    g64 = to_device(g)
    sg = None if np.all(sign == 1.0) else to_device(sign.astype(np.float32))
    w32 = to_device(Wt.astype(np.float32))
    return ClassCondState(g64.to(torch.float32), g64, w32, sg,
                          to_device(Mc.astype(np.float32)), to_device(valid.astype(np.int32)),
                          cm.shape[0], Wt.shape[1], Wt.shape[0],
                          split_tf32(w32) if (Wt.shape[1] % 4 == 0 and Wt.shape[0] % 4 == 0) else None)


def classcond_score(x, st: ClassCondState, out_dtype=torch.float64) -> torch.Tensor:
    if x.ndim != 2 or x.shape[1] != st.d:
        raise ValueError(f"expected rows of width {st.d}, got an array of shape {tuple(x.shape)}")
    xf, centered = as_f32_rows(x, st.g_f64)
    n = xf.shape[0]
    out = _empty((n,), out_dtype)
    o64 = out.data_ptr() if out_dtype == torch.float64 else None
    o32 = out.data_ptr() if out_dtype == torch.float32 else None
    g = None if centered else st.g_f32.data_ptr()
    if _tc_ok(st.d) and st.planes is not None and st.C <= TC_MAX_CLASSES:
        _lib.call("runia_classcond_mahalanobis_tc", xf.data_ptr(), n, st.d, g, st.planes[0].data_ptr(),
                  st.planes[1].data_ptr(), st.r, ptr(st.sign), st.Mc.data_ptr(), st.valid.data_ptr(), st.C, o64, o32,
                  stream_ptr())
    else:
        _lib.call("runia_classcond_mahalanobis_f32", xf.data_ptr(), n, st.d, g, st.Wt.data_ptr(), st.r, ptr(st.sign),
                  st.Mc.data_ptr(), st.valid.data_ptr(), st.C, o64, o32, stream_ptr())
    return out


# ------------------------------------------------------------------------------------------
# (a9) GMM / DDU
# ------------------------------------------------------------------------------------------
@dataclass
class GMMState:
    At: torch.Tensor        # [C * dpad, d]
    off: torch.Tensor       # [C * dpad]
    logconst: torch.Tensor  # [C]
    C: int
    d: int
    dpad: int
    planes: Optional[tuple] = None


def _dev64(a) -> torch.Tensor:
    if isinstance(a, torch.Tensor):
        return a.detach().to(device=device(), dtype=torch.float64).contiguous()
    return to_device(np.ascontiguousarray(a, np.float64), torch.float64)


def tril_inverse(L) -> torch.Tensor:
    """L^{-1} for a batch [B, n, n] of lower-triangular float64 factors, on the device (`runia_tril_inverse_f64`)."""
    l = _dev64(L)
    B, n, _ = l.shape
    X = _empty((B, n, n), torch.float64)
    _lib.call("runia_tril_inverse_f64", l.data_ptr(), B, n, X.data_ptr(), stream_ptr())
    return X


def gmm_prepare(means, scale_tril) -> GMMState:
    """means [C, d], scale_tril [C, d, d] (lower Cholesky factors; ndarray or tensor) -> whitening blocks
    A_c = L_c^{-1} (float64 forward substitution on the device), offsets A_c mu_c and the log-normalisers."""
    mu = _dev64(means)
    L = _dev64(scale_tril)
    C, d = mu.shape
    dpad = (d + 127) // 128 * 128
    Linv = tril_inverse(L)
    At = torch.zeros((C, dpad, d), dtype=torch.float64, device=mu.device)
    At[:, :d] = Linv
    off = torch.zeros((C, dpad), dtype=torch.float64, device=mu.device)
    off[:, :d] = torch.bmm(Linv, mu[:, :, None])[:, :, 0]
    logconst = -torch.log(torch.diagonal(L, dim1=1, dim2=2)).sum(1) - 0.5 * d * float(np.log(2 * np.pi))
    a32 = At.reshape(C * dpad, d).to(torch.float32)
    return GMMState(a32, off.reshape(-1).to(torch.float32), logconst.to(torch.float32), C, d, dpad,
                    split_tf32(a32) if d % 4 == 0 else None)


def gmm_lse(x, st: GMMState) -> torch.Tensor:
    if x.ndim != 2 or x.shape[1] != st.d:
        raise ValueError(f"expected rows of width {st.d}, got an array of shape {tuple(x.shape)}")
    xf, _ = as_f32_rows(x, None)
    n = xf.shape[0]
    out = _empty((n,), torch.float32)
    if _tc_ok(st.d) and st.planes is not None:
        _lib.call("runia_gmm_lse_tc", xf.data_ptr(), n, st.d, st.planes[0].data_ptr(), st.planes[1].data_ptr(),
                  st.off.data_ptr(), st.dpad, st.logconst.data_ptr(), st.C, out.data_ptr(), stream_ptr())
    else:
        _lib.call("runia_gmm_lse_f32", xf.data_ptr(), n, st.d, st.At.data_ptr(), st.off.data_ptr(), st.dpad,
                  st.logconst.data_ptr(), st.C, out.data_ptr(), stream_ptr())
    return out


# ------------------------------------------------------------------------------------------
# (a5) kNN
# ------------------------------------------------------------------------------------------
def normalize_rows(x) -> torch.Tensor:
    t = to_device(x)
    if t.dtype not in (torch.float32, torch.float64):
        t = t.to(torch.float32)
    if t.dim() == 1:
        t = t.reshape(1, -1)
    n, d = t.shape
    out = _empty((n, d), torch.float32)
    _lib.call("runia_normalize_rows", t.data_ptr(), 1 if t.dtype == torch.float64 else 0, n, d,
              out.data_ptr(), stream_ptr())
    return out


def row_sqnorm(x: torch.Tensor) -> torch.Tensor:
    out = _empty((x.shape[0],), torch.float32)
    _lib.call("runia_row_sqnorm_f32", x.data_ptr(), x.shape[0], x.shape[1], out.data_ptr(), stream_ptr())
    return out


@dataclass
class KNNBank:
    bank: torch.Tensor    # [Nb, d] float32, already normalised
    sqnorm: torch.Tensor  # [Nb]
    idx_offset: int = 0
    planes: Optional[tuple] = None  # tf32 (hi, lo) planes for the tensor-core candidate pass
    filter_products: Optional[dict] = None  # k -> 1 | 3: TF32 products of the candidate filter (knn_filter_products)


def knn_bank(bank_normed: torch.Tensor, idx_offset: int = 0, planes: Optional[bool] = None) -> KNNBank:
    if planes is None:
        planes = _ENGINE == "tc"
    pl = split_tf32(bank_normed) if (planes and bank_normed.shape[1] % 4 == 0) else None
    return KNNBank(bank_normed, row_sqnorm(bank_normed), idx_offset, pl)


class KNNOverflow(RuntimeError):
    """Kept for API compatibility: the exhaustive pass no longer has a tie-buffer limit, nothing raises this."""


KNN_DENSITY_ROWS = 256   # bank rows sampled as queries by the density probe
KNN_DENSITY_BAND = 200   # band population above which a row is likely to leave the single-product filter's scratch


def knn_filter_products(bank: KNNBank, k: int) -> int:
    """1 or 3: how many TF32 products the candidate filter of this bank uses for `k` neighbours.
    The single-product filter certifies its result through a rounding bound of ~3e-3 per unit of squared norm; every
    bank row within twice that of the k-th distance gets an exact evaluation, and a row with more than 4 kseed such
    neighbours (256 at k = 50) goes to the exhaustive pass.  Synthetic high-dimensional data never gets there; a large
    bank of real embeddings can (neighbour spacing ~1e-6 at a million rows).  So the first search of a bank at a given k
    samples 256 bank rows as queries, searches them with the FP32-faithful 3xTF32 filter for 256 neighbours and counts
    how many lie within the band of the k-th (the row itself excluded): if more than a tenth of the sampled rows have
    more than 200, the bank is searched with three products (the round-1 path: 3x the tensor work, a 50x tighter bound).
    One device -> host read per (bank, k), then cached on the bank."""
    if bank.filter_products is None:
        bank.filter_products = {}
    if k in bank.filter_products:
        return bank.filter_products[k]
    nb, d = bank.bank.shape
    mode = 1
    if bank.planes is not None and _tc_ok(d) and nb > 4 * KNN_DENSITY_ROWS and k < 200:
        m = KNN_DENSITY_ROWS
        rows = torch.linspace(0, nb - 1, m, device=bank.bank.device, dtype=torch.float64).long()
        q = bank.bank[rows].contiguous()
        kk = 256
        res = {"dist": None, "dist64": _empty((m, kk), torch.float64), "idx": None, "kth": _empty((m,), torch.float32)}
        _knn_search_chunk(q, bank, kk, res, 0, m, False, products=3)
        d64 = res["dist64"]
        eps1 = 2.0 * (2.0 ** -10 + 2.0 ** -11 + 2.0 ** -21) * 1.0005 + 2.5 * (2.0 * d + 8.0) * 2.0 ** -24
        s_row = eps1 * 0.5 * (bank.sqnorm[rows].double() + bank.sqnorm.max().double())
        band = (d64 <= (d64[:, k] + 2.0 * s_row)[:, None]).sum(1) - 1  # entry 0 is the row itself
        dense_rows = int((band > KNN_DENSITY_BAND).sum().item())
        mode = 3 if dense_rows * 10 > m else 1
    bank.filter_products[k] = mode
    return mode


KNN_MAX_K = 1016               # distance.cu kKnnMaxK
KNN_WS_CHUNK_BYTES = 12 << 30  # candidate-list workspace above this is avoided by searching the queries in chunks


def _knn_search_chunk(qn, bank, k, res, lo, hi, want_status, products=1):
    nq, d = hi - lo, qn.shape[1]
    nb = bank.bank.shape[0]
    ws_bytes = int(_lib.raw("runia_knn_workspace_bytes")(nq, nb, d, k))
    ws = _empty((ws_bytes,), torch.uint8)
    status = _empty((4,), torch.int32)
    use_tc = _tc_ok(d) and bank.planes is not None
    off = lambda t, w: None if t is None else t.data_ptr() + lo * w * t.element_size()  # noqa: E731
    _lib.call("runia_knn_search_ex_f32", qn.data_ptr() + lo * d * 4, nq, bank.bank.data_ptr(), bank.sqnorm.data_ptr(),
              bank.planes[0].data_ptr() if use_tc else None, bank.planes[1].data_ptr() if use_tc else None,
              nb, d, k, bank.idx_offset, off(res["dist"], k), off(res["dist64"], k), off(res["idx"], k),
              off(res["kth"], 1), status.data_ptr(), ws.data_ptr(), ws_bytes, int(products), stream_ptr())
    return status if want_status else None


def knn_search(qn: torch.Tensor, bank: KNNBank, k: int, want_idx=True, want_dist=True, want_f64=False,
               check_status=True, products=None):
    """qn: [Nq, d] float32 CUDA (normalised for the postprocessors; any rows for FlatL2Index).  Returns
    dict(dist [Nq,k] f32, dist64, idx [Nq,k] i64, kth [Nq] f32, exhaustive_rows int).  `check_status` only fills
    `exhaustive_rows` (one device -> host read); the search itself cannot fail on ties.  `products`: TF32 products
    of the candidate filter (1 or 3); None = what `knn_filter_products` measured for this bank."""
    nq, d = qn.shape
    nb = bank.bank.shape[0]
    res = {"dist": None, "dist64": None, "idx": None, "kth": _empty((nq,), torch.float32), "exhaustive_rows": 0}
    if want_dist:
        res["dist"] = _empty((nq, k), torch.float32)
    if want_f64:
        res["dist64"] = _empty((nq, k), torch.float64)
    if want_idx:
        res["idx"] = _empty((nq, k), torch.int64)
    if nq == 0:
        return res
    if not 1 <= k <= KNN_MAX_K:
        raise NotImplementedError(f"kNN: k={k} outside [1, {KNN_MAX_K}]")
    qn = qn.contiguous()
    if products is None:
        products = knn_filter_products(bank, k)
    chunk = nq
    while chunk > 256 and int(_lib.raw("runia_knn_workspace_bytes")(chunk, nb, d, k)) > KNN_WS_CHUNK_BYTES:
        chunk = (chunk // 2 + 255) // 256 * 256
    stats = [_knn_search_chunk(qn, bank, k, res, lo, min(nq, lo + chunk), check_status, products)
             for lo in range(0, nq, chunk)]
    if check_status:
        res["exhaustive_rows"] = int(torch.stack(stats)[:, 0].sum().item())
    return res


def topk_merge(part_dist64: torch.Tensor, part_idx: torch.Tensor):
    """[R, Nq, k] partial lists -> (dist [Nq,k] f32, idx [Nq,k] i64, kth [Nq] f32)."""
    R, nq, k = part_dist64.shape
    dist = _empty((nq, k), torch.float32)
    idx = _empty((nq, k), torch.int64)
    kth = _empty((nq,), torch.float32)
    _lib.call("runia_topk_merge", part_dist64.data_ptr(), part_idx.data_ptr(), R, nq, k, dist.data_ptr(),
              idx.data_ptr(), kth.data_ptr(), stream_ptr())
    return dist, idx, kth


# ------------------------------------------------------------------------------------------
# (a4) KDE
# ------------------------------------------------------------------------------------------
@dataclass
class KDEBank:
    bank: torch.Tensor      # [Nb, d] float32, centred by `center`
    center: torch.Tensor    # [d] float64
    bandwidth: float
    n_total: int
    planes: Optional[tuple] = None


def kde_bank(train, bandwidth=1.0, center=None, n_total=None) -> KDEBank:
    t = to_device(train)
    if center is None:
        center = t.to(torch.float64).mean(0)  # fit-time statistic (setup), any centre is valid
    c = to_device(center, torch.float64)
    n, d = t.shape
    out = _empty((n, d), torch.float32)
    _lib.call("runia_center_cast", t.data_ptr(), 1 if t.dtype == torch.float64 else 0, n, d, c.data_ptr(),
              out.data_ptr(), stream_ptr())
    pl = split_tf32(out) if (_ENGINE == "tc" and d % 4 == 0) else None
    return KDEBank(out, c, float(bandwidth), int(n_total if n_total is not None else n), pl)


def _kde_stage_queries(q, kb: KDEBank):
    t = to_device(q)
    if t.dtype not in (torch.float32, torch.float64):
        t = t.to(torch.float32)
    n, d = t.shape
    out = _empty((n, d), torch.float32)
    _lib.call("runia_center_cast", t.data_ptr(), 1 if t.dtype == torch.float64 else 0, n, d,
              kb.center.data_ptr(), out.data_ptr(), stream_ptr())
    return out


def kde_score(q, kb: KDEBank, partial=False):
    """log-density of each query under the Gaussian KDE of the bank (float64), or the partial
    (max, sum) pair for a bank shard."""
    qc = _kde_stage_queries(q, kb)
    nq, d = qc.shape
    nb = kb.bank.shape[0]
    ws_bytes = int(_lib.raw("runia_kde_workspace_bytes")(nq, nb)) if nq else 0
    ws = _empty((max(ws_bytes, 1),), torch.uint8)
    use_tc = _tc_ok(d) and kb.planes is not None
    hi = kb.planes[0].data_ptr() if use_tc else None
    lo = kb.planes[1].data_ptr() if use_tc else None
    if partial:
        m = _empty((nq,), torch.float32)
        s = _empty((nq,), torch.float32)
        _lib.call("runia_kde_lse_f32", qc.data_ptr(), nq, kb.bank.data_ptr(), hi, lo, nb, d, kb.bandwidth, kb.n_total,
                  None, m.data_ptr(), s.data_ptr(), ws.data_ptr(), ws_bytes, stream_ptr())
        return m, s
    out = _empty((nq,), torch.float64)
    _lib.call("runia_kde_lse_f32", qc.data_ptr(), nq, kb.bank.data_ptr(), hi, lo, nb, d, kb.bandwidth, kb.n_total,
              out.data_ptr(), None, None, ws.data_ptr(), ws_bytes, stream_ptr())
    return out


# ------------------------------------------------------------------------------------------
# (a8) logit scores, (a10) clip-linear-LSE / ASH
# ------------------------------------------------------------------------------------------
def logit_scores(logits, gamma=0.1, M=None, energy=True, msp=True, gen=True):
    lg = to_device(logits)
    in_dtype = lg.dtype
    lg = lg.to(torch.float32) if lg.dtype != torch.float32 else lg
    n, C = lg.shape
    if M is None:
        M = C
    e = _empty((n,), torch.float32) if energy else None
    m = _empty((n,), torch.float32) if msp else None
    g = _empty((n,), torch.float32) if gen else None
    _lib.call("runia_logit_scores_f32", lg.data_ptr(), n, C, float(gamma), int(M), ptr(e), ptr(m), ptr(g),
              stream_ptr())
    return e, m, g, in_dtype


LINEAR_TC_NARROW_CLASSES = 32  # UMMA N of the narrow-panel kernel; wider heads use 256-column panels
LINEAR_TC_MIN_ROWS = 16384     # narrow heads: below this the one-warp-per-row kernel is as fast and needs no planes
LINEAR_SMEM_MAX_CLASSES = 64   # logits.cu CL_MAXC: heads with C <= 64 and C*d*4 <= 200 KiB stay resident in shared memory


def linear_planes(W: torch.Tensor):
    """TF32 hi / lo planes of the head's weight matrix for runia_clip_linear_lse_tc: zero-padded to [32, d] for
    C <= 32 (narrow panel), [C, d] as it is for wider heads."""
    if W.shape[0] <= LINEAR_TC_NARROW_CLASSES:
        Wp = torch.zeros((LINEAR_TC_NARROW_CLASSES, W.shape[1]), dtype=torch.float32, device=device())
        Wp[: W.shape[0]] = W
        return split_tf32(Wp)
    return split_tf32(W.contiguous())


def head_fits_smem(C: int, d: int) -> bool:
    return C <= LINEAR_SMEM_MAX_CLASSES and C * d * 4 <= 200 * 1024


def clip_linear_lse(x, W: torch.Tensor, b: torch.Tensor, clip=float("inf"), planes=None) -> torch.Tensor:
    """logsumexp(min(x, clip) @ W.T + b): the ReAct / DICE / DICE+ReAct head, any C and d.  Rows with d % 4 == 0
    stream through the tcgen05 kernels (narrow panel for C <= 32 and large batches, 256-column panels with an
    online log-sum-exp for any wider head); `planes` = linear_planes(W) (built here when absent).  Other shapes use
    the FP32 SIMT kernels (resident head when it fits shared memory, streamed otherwise)."""
    xf, _ = as_f32_rows(x, None)
    n, d = xf.shape
    C = W.shape[0]
    out = _empty((n,), torch.float32)
    aligned = xf.data_ptr() % 16 == 0 and b.data_ptr() % 16 == 0
    wide = C > LINEAR_TC_NARROW_CLASSES
    if _tc_ok(d) and aligned and (wide or n >= LINEAR_TC_MIN_ROWS):
        hi, lo = planes if planes is not None else linear_planes(W)
        _lib.call("runia_clip_linear_lse_tc", xf.data_ptr(), n, d, hi.data_ptr(), lo.data_ptr(), b.data_ptr(), C,
                  float(clip), out.data_ptr(), stream_ptr())
        return out
    _lib.call("runia_clip_linear_lse_f32", xf.data_ptr(), n, d, W.data_ptr(), b.data_ptr(), C,
              float(clip), out.data_ptr(), stream_ptr())
    return out


def ash_prune(x, k_keep: int) -> torch.Tensor:
    """ASH-S pruned and rescaled rows [N, d] float32 (funcs.py:230-261) for any row width."""
    xf, _ = as_f32_rows(x, None)
    n, d = xf.shape
    out = _empty((n, d), torch.float32)
    _lib.call("runia_ash_prune_f32", xf.data_ptr(), n, d, int(k_keep), out.data_ptr(), stream_ptr())
    return out


def ash_linear_lse(x, W: torch.Tensor, b: torch.Tensor, k_keep: int, planes=None) -> torch.Tensor:
    """ASH-S score.  Heads that fit shared memory run the fused prune + head kernel; any other head prunes into a
    scratch [N, d] and goes through clip_linear_lse (tensor cores for d % 4 == 0)."""
    xf, _ = as_f32_rows(x, None)
    n, d = xf.shape
    if head_fits_smem(W.shape[0], d):
        out = _empty((n,), torch.float32)
        _lib.call("runia_ash_linear_lse_f32", xf.data_ptr(), n, d, W.data_ptr(), b.data_ptr(), W.shape[0],
                  int(k_keep), out.data_ptr(), stream_ptr())
        return out
    return clip_linear_lse(ash_prune(xf, k_keep), W, b, planes=planes)


def gen_entropy_from_probs(probs, gamma: float, M: int) -> torch.Tensor:
    """generalized_entropy on rows that already are probabilities (no softmax; funcs.py:347-375)."""
    p = to_device(probs, torch.float32)
    n, C = p.shape
    out = _empty((n,), torch.float32)
    _lib.call("runia_gen_entropy_f32", p.data_ptr(), n, C, float(gamma), int(M), out.data_ptr(), stream_ptr())
    return out


def linear(x, W: torch.Tensor, b: Optional[torch.Tensor]) -> torch.Tensor:
    """x @ W.T + b as float32 [N, C] (FP32 SIMT contraction): RouteDICE.forward."""
    xf, _ = as_f32_rows(x, None)
    n, d = xf.shape
    out = _empty((n, W.shape[0]), torch.float32)
    _lib.call("runia_linear_f32", xf.data_ptr(), n, d, W.data_ptr(), ptr(b), W.shape[0], out.data_ptr(), stream_ptr())
    return out


# ------------------------------------------------------------------------------------------
# (f1) OoD detection metrics (AUROC, FPR@95, AUPR, ROC curve)
# ------------------------------------------------------------------------------------------
def ood_metrics(ind_scores, ood_scores, want_curve: bool = True):
    """InD (positive) vs OoD scores -> dict(auroc, fpr95, aupr, fpr, tpr).  Scores may live on the
    host or the device; float64 stays float64 (the sigmoid torchmetrics applies to scores outside
    [0, 1] runs in the score dtype), everything else is scored as float32.  `fpr` / `tpr` are float32
    CUDA tensors (one point per distinct score after the origin) or None."""
    a, b = to_device(ind_scores).reshape(-1), to_device(ood_scores).reshape(-1)
    f64 = a.dtype == torch.float64 or b.dtype == torch.float64
    dt = torch.float64 if f64 else torch.float32
    a, b = a.to(dt).contiguous(), b.to(dt).contiguous()
    n_ind, n_ood = a.numel(), b.numel()
    if n_ind == 0 or n_ood == 0:
        raise ValueError("ood_metrics: both score arrays must be non-empty")
    ws_bytes = int(_lib.raw("runia_ood_metrics_workspace_bytes")(n_ind, n_ood))
    ws = _empty((ws_bytes,), torch.uint8)
    out4 = _empty((4,), torch.float64)
    fpr = _empty((n_ind + n_ood + 1,), torch.float32) if want_curve else None
    tpr = _empty((n_ind + n_ood + 1,), torch.float32) if want_curve else None
    _lib.call("runia_ood_metrics", a.data_ptr(), n_ind, b.data_ptr(), n_ood, 1 if f64 else 0, out4.data_ptr(),
              ptr(fpr), ptr(tpr), ws.data_ptr(), ws_bytes, stream_ptr())
    o = out4.cpu()
    npts = int(o[3])
    return {"auroc": float(o[0]), "fpr95": float(o[1]), "aupr": float(o[2]), "n_points": npts,
            "fpr": fpr[:npts] if want_curve else None, "tpr": tpr[:npts] if want_curve else None}


# ------------------------------------------------------------------------------------------
# (f2) order statistics for the ReAct threshold
# ------------------------------------------------------------------------------------------
def sort_f32(x) -> torch.Tensor:
    """Ascending device sort of all elements of x (float32)."""
    t = to_device(x, torch.float32).reshape(-1).contiguous()
    n = t.numel()
    out = _empty((n,), torch.float32)
    if n == 0:
        return out
    ws_bytes = int(_lib.raw("runia_sort_f32_workspace_bytes")(n))
    ws = _empty((ws_bytes,), torch.uint8)
    _lib.call("runia_sort_f32", t.data_ptr(), n, out.data_ptr(), ws.data_ptr(), ws_bytes, stream_ptr())
    return out


def percentile_f32(x, q: float):
    """np.percentile(x.flatten(), q) (method "linear") for float32 data.  The two neighbouring order
    statistics come from the device sort; the virtual index, gamma and the lerp follow NumPy's own
    `_quantile` / `_lerp` arithmetic (NumPy >= 2 casts q to the array dtype, so everything is float32 and
    the index saturates for huge arrays; older NumPy keeps the index in float64)."""
    s = sort_f32(x)
    n = s.numel()
    f32 = int(np.__version__.split(".")[0]) >= 2
    ft = np.float32 if f32 else np.float64
    vi = ft(n - 1) * np.true_divide(np.asanyarray(q, dtype=ft), ft(100))
    if vi >= n - 1:
        return np.float32(s[n - 1].item())
    if vi < 0:
        return np.float32(s[0].item())
    lo = int(np.floor(vi))
    a, b = (np.float32(v) for v in s[[lo, lo + 1]].cpu().numpy())
    g = ft(vi - ft(lo))
    diff = b - a
    out = a + diff * g if g < 0.5 else b - diff * (ft(1) - g)
    return np.float32(out) if f32 else out


# ------------------------------------------------------------------------------------------
# (f2) setup() statistics: class means and the pooled class-centred covariance
# ------------------------------------------------------------------------------------------
def _labels_i32(labels, n):
    lab = np.asarray(labels.detach().cpu() if isinstance(labels, torch.Tensor) else labels).reshape(-1)
    if lab.shape[0] != n:
        raise ValueError(f"{lab.shape[0]} labels for {n} rows")
    if lab.dtype.kind == "f":  # float labels compare like `labels == c`: only integral values match a class
        lab = np.where(lab == np.floor(lab), lab, -1)
    big = np.abs(lab) >= 2 ** 31 if lab.dtype.kind in "iuf" else False
    lab = np.where(big, -1, lab).astype(np.int32)
    return torch.from_numpy(lab).to(device())


def class_means(x, labels=None, num_classes: int = 1):
    """(means [C, d] float32 device, counts [C] int64 host): `x[labels == c].mean(0)` with NumPy's own
    row-by-row float32 accumulation (bit-identical); labels None -> the column mean of all rows, C = 1."""
    xf = to_device(x, torch.float32)
    n, d = xf.shape
    lab = None if labels is None else _labels_i32(labels, n)
    means = _empty((num_classes, d), torch.float32)
    counts = _empty((num_classes,), torch.int64)
    _lib.call("runia_class_mean_f32", xf.data_ptr(), ptr(lab), n, d, num_classes, means.data_ptr(), counts.data_ptr(),
              stream_ptr())
    return means, counts.cpu().numpy(), xf, lab


def centered_gram(xf: torch.Tensor, lab, centers: torch.Tensor):
    """(G [d, d], colsum [d]) float64 device tensors: sum_i r_i r_i^T and sum_i r_i over the rows with a label in
    [0, C), r_i = f32(x_i - centers[label_i]).  Deterministic; additive over row shards (all-reduce both)."""
    n, d = xf.shape
    G = _empty((d, d), torch.float64)
    cs = _empty((d,), torch.float64)
    ws_bytes = int(_lib.raw("runia_centered_gram_workspace_bytes")(n, d))
    ws = _empty((ws_bytes,), torch.uint8)
    _lib.call("runia_centered_gram_f64", xf.data_ptr(), ptr(lab), centers.data_ptr(), n, d, centers.shape[0], G.data_ptr(),
              cs.data_ptr(), ws.data_ptr(), ws_bytes, stream_ptr())
    return G, cs


def shifted_covariance(xf: torch.Tensor, center) -> np.ndarray:
    """(x - u)^T (x - u) / N in float64 for float32 device rows x and ONE float64 centre u: what sklearn's
    EmpiricalCovariance(assume_centered=True).fit(x - u).covariance_ holds (ViM.setup)."""
    n, d = xf.shape
    u = to_device(np.ascontiguousarray(center, np.float64).reshape(-1), torch.float64)
    if u.numel() != d:
        raise ValueError(f"centre of {u.numel()} entries for rows of width {d}")
    G = _empty((d, d), torch.float64)
    ws_bytes = int(_lib.raw("runia_centered_gram_workspace_bytes")(n, d))
    ws = _empty((ws_bytes,), torch.uint8)
    _lib.call("runia_shifted_gram_f64", xf.data_ptr(), u.data_ptr(), n, d, G.data_ptr(), None, ws.data_ptr(), ws_bytes,
              stream_ptr())
    return (G / n).cpu().numpy()


def covariance_from_gram(G, cs, n_used: int) -> np.ndarray:
    """np.cov(R.T, bias=1) from the Gram matrix and column sums of the residual rows: np.cov's own re-centring
    and 1/n scaling on the d x d result."""
    G = G.cpu().numpy() if isinstance(G, torch.Tensor) else np.asarray(G)
    cs = cs.cpu().numpy() if isinstance(cs, torch.Tensor) else np.asarray(cs)
    avg = cs / n_used
    cov = G - n_used * np.outer(avg, avg)
    cov *= np.true_divide(1, n_used)
    return cov


def centered_covariance(xf: torch.Tensor, lab, centers: torch.Tensor, n_used: int) -> np.ndarray:
    """np.cov(R.T, bias=1) in float64 for the residual rows R = f32(x - centers[label]) (what
    sklearn's EmpiricalCovariance(assume_centered=False).fit(R).covariance_ holds)."""
    G, cs = centered_gram(xf, lab, centers)
    return covariance_from_gram(G, cs, n_used)


# ------------------------------------------------------------------------------------------
# (f3) MC-DropBlock sampler + spatial mean
# ------------------------------------------------------------------------------------------
def mc_dropblock_mean(x, seed, block_size: int) -> torch.Tensor:
    """x [B, C, H, W] float32, seed [n_mc, B, H, W] (non-zero = Bernoulli hit) -> [B * n_mc, C] float32 device rows
    (item-major): every DropBlock2D mask applied and reduced over H x W in one pass over x."""
    xf = to_device(x, torch.float32)
    sd = seed.to(device=device(), dtype=torch.uint8).contiguous() if isinstance(seed, torch.Tensor) else \
        torch.from_numpy(np.ascontiguousarray(np.asarray(seed) != 0).astype(np.uint8)).to(device())
    B, C, H, W = xf.shape
    n_mc = sd.shape[0]
    if tuple(sd.shape) != (n_mc, B, H, W):
        raise ValueError(f"seed shape {tuple(sd.shape)} does not match (n_mc, {B}, {H}, {W})")
    if n_mc > MC_SAMPLER_MAX:  # the kernel keeps one accumulator per sample in registers (<= 32): more samples in passes
        parts = [mc_dropblock_mean(xf, sd[m0:m0 + MC_SAMPLER_MAX], block_size).reshape(B, -1, C)
                 for m0 in range(0, n_mc, MC_SAMPLER_MAX)]
        return torch.cat(parts, dim=1).reshape(B * n_mc, C)
    out = _empty((B * n_mc, C), torch.float32)
    ws_bytes = int(_lib.raw("runia_mc_dropblock_workspace_bytes")(B, H, W, n_mc))
    ws = _empty((ws_bytes,), torch.uint8)
    _lib.call("runia_mc_dropblock_mean_f32", xf.data_ptr(), sd.data_ptr(), B, C, H, W, n_mc, int(block_size), out.data_ptr(),
              ws.data_ptr(), ws_bytes, stream_ptr())
    return out


MC_SAMPLER_MAX = 32  # sampler.cu: MC samples per launch


def mc_dropblock_apply(x, seed, block_size: int) -> torch.Tensor:
    """x [B, C, H, W] float32, seed [n_mc, B, H, W] -> [n_mc, B * C * H * W]: the DropBlock2D-masked, renormalised maps
    themselves (MCSamplerModule layer types "FC" / "RPN")."""
    xf = to_device(x, torch.float32)
    sd = seed.to(device=device(), dtype=torch.uint8).contiguous() if isinstance(seed, torch.Tensor) else \
        torch.from_numpy(np.ascontiguousarray(np.asarray(seed) != 0).astype(np.uint8)).to(device())
    B, C, H, W = xf.shape
    n_mc = sd.shape[0]
    if n_mc > MC_SAMPLER_MAX:
        return torch.cat([mc_dropblock_apply(xf, sd[m0:m0 + MC_SAMPLER_MAX], block_size)
                          for m0 in range(0, n_mc, MC_SAMPLER_MAX)], dim=0)
    out = _empty((n_mc, B * C * H * W), torch.float32)
    ws_bytes = int(_lib.raw("runia_mc_dropblock_workspace_bytes")(B, H, W, n_mc))
    ws = _empty((ws_bytes,), torch.uint8)
    _lib.call("runia_mc_dropblock_apply_f32", xf.data_ptr(), sd.data_ptr(), B, C, H, W, n_mc, int(block_size),
              out.data_ptr(), ws.data_ptr(), ws_bytes, stream_ptr())
    return out


# ------------------------------------------------------------------------------------------
# (f3) object-level reducers: RoIAlign maps and their per-channel means
# ------------------------------------------------------------------------------------------
def _roi_args(feat, boxes, output_size):
    x = to_device(feat, torch.float32)
    if x.dim() == 3:
        x = x.unsqueeze(0)
    bx = to_device(boxes, torch.float32).reshape(-1, 4).contiguous()
    ph, pw = (output_size, output_size) if isinstance(output_size, int) else tuple(output_size)
    return x, bx, int(ph), int(pw)


def roi_align(feat, boxes, output_size, spatial_scale=1.0, sampling_ratio=-1, aligned=False) -> torch.Tensor:
    """torchvision.ops.roi_align(feat, [boxes], ...) for the boxes of image 0: [K, C, P, P] float32."""
    x, bx, ph, pw = _roi_args(feat, boxes, output_size)
    B, C, H, W = x.shape
    out = _empty((bx.shape[0], C, ph, pw), torch.float32)
    _lib.call("runia_roi_align_f32", x.data_ptr(), B, C, H, W, bx.data_ptr(), None, bx.shape[0], ph, pw,
              float(spatial_scale), int(sampling_ratio), 1 if aligned else 0, out.data_ptr(), stream_ptr())
    return out


def roi_align_mean(feat, boxes, output_size, spatial_scale=1.0, sampling_ratio=-1, aligned=False, want_std=False):
    """mean (and unbiased std) over the pooled bins of every (box, channel): ([K, C], [K, C] or None) float32."""
    x, bx, ph, pw = _roi_args(feat, boxes, output_size)
    B, C, H, W = x.shape
    mean = _empty((bx.shape[0], C), torch.float32)
    std = _empty((bx.shape[0], C), torch.float32) if want_std else None
    _lib.call("runia_roi_align_mean_f32", x.data_ptr(), B, C, H, W, bx.data_ptr(), None, bx.shape[0], ph, pw,
              float(spatial_scale), int(sampling_ratio), 1 if aligned else 0, mean.data_ptr(), ptr(std), stream_ptr())
    return mean, std


# ------------------------------------------------------------------------------------------
# (f2) float64 eigendecomposition / pseudo-inverse / Cholesky for the setup() fits
# ------------------------------------------------------------------------------------------
def _eigh_device(a_np):
    a = to_device(np.ascontiguousarray(a_np, np.float64), torch.float64)
    n = a.shape[0]
    evals = _empty((n,), torch.float64)
    evecs = _empty((n, n), torch.float64)
    ws_bytes = int(_lib.raw("runia_eigh_workspace_bytes")(n))
    ws = _empty((ws_bytes,), torch.uint8)
    _lib.call("runia_eigh_f64", a.data_ptr(), n, evals.data_ptr(), evecs.data_ptr(), ws.data_ptr(), ws_bytes, 0, None,
              stream_ptr())
    resid = float((a @ evecs.T - evecs.T * evals).abs().max()) if n else 0.0  # |A V - V diag(lambda)|, V = evecs^T
    return evals.cpu().numpy(), evecs.cpu().numpy().T, resid


def eigh(A):
    """Symmetric eigendecomposition on the device: (evals ascending [n], V [n, n] with eigenvectors in the COLUMNS),
    NumPy float64 -- the contract of numpy.linalg.eigh / scipy.linalg.eigh (eigenvector signs are arbitrary there too).
    The one-sided Jacobi kernel diagonalises A^T A; for the positive semi-definite matrices of the fits (covariances,
    precisions) that is the eigenbasis of A itself.  An indefinite A with eigenvalues +x and -x would leave their two
    eigenvectors mixed: the residual |A V - V diag(lambda)| is checked, and such a matrix is decomposed again as
    A + sigma I (sigma = its Frobenius norm: positive definite, same eigenvectors)."""
    A = np.ascontiguousarray(A, np.float64)
    n = A.shape[0]
    assert A.ndim == 2 and A.shape[1] == n
    A = 0.5 * (A + A.T)
    lam, V, resid = _eigh_device(A)
    scale = max(float(np.abs(A).max(initial=0.0)), 1e-300)
    if resid > 1e-9 * scale * max(1.0, np.sqrt(n)):
        sigma = float(np.sqrt((A * A).sum()))
        lam, V, _ = _eigh_device(A + sigma * np.eye(n))
        lam = lam - sigma
    order = np.argsort(lam, kind="stable")
    return lam[order], np.ascontiguousarray(V[:, order])


def pinvh(a):
    """scipy.linalg.pinvh(a) (sklearn's `_set_covariance` -> `precision_`) from the device eigendecomposition, with
    scipy's cut-off: eigenvalues with |lambda| <= max(M, N) * eps * max|lambda| are dropped."""
    a = np.asarray(a, np.float64)
    lam, V = eigh(a)
    rtol = max(a.shape) * np.finfo(np.float64).eps
    keep = np.abs(lam) > rtol * np.abs(lam).max(initial=0.0)
    u = V[:, keep]
    P = mm64(u * (1.0 / lam[keep]), u.T)
    # the scorers factor the precision right after the fit (factor_precision): P = u diag(1 / lambda) u^T is that
    # factorisation -- remembered for the array object returned here (a few entries; a copy of P misses and is redone)
    if len(_PINVH_FACTORS) >= 4:
        _PINVH_FACTORS.pop(next(iter(_PINVH_FACTORS)))
    _PINVH_FACTORS[id(P)] = (P, 1.0 / lam[keep], np.ascontiguousarray(u), float((0.5 * (P + P.T)).sum()))
    return P


def cholesky_batch(A, jitter: float = 0.0, rel_pivot: float = 0.0):
    """(L [B, n, n] float64 lower, fail [B] int32) with A_b + jitter I = L_b L_b^T on the device; A: ndarray or CUDA
    tensor [B, n, n].  rel_pivot > 0 also fails pivots below rel_pivot x their original diagonal entry."""
    a = A.to(torch.float64).contiguous() if isinstance(A, torch.Tensor) and A.is_cuda else \
        to_device(np.ascontiguousarray(A, np.float64), torch.float64)
    B, n, _ = a.shape
    L = _empty((B, n, n), torch.float64)
    fail = _empty((B,), torch.int32)
    _lib.call("runia_cholesky_f64", a.data_ptr(), B, n, float(jitter), float(rel_pivot), L.data_ptr(), fail.data_ptr(),
              stream_ptr())
    return L, fail.cpu().numpy()
