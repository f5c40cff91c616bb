"""Device staging helpers: PyTorch is used for device memory, streams and torch.distributed
only -- all arithmetic happens in libruniab200.so."""
import numpy as np
import torch


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError(
            "runia_core_b200 needs a CUDA device (B200 / sm_100a); there is no CPU fallback")


def device():
    require_cuda()
    return torch.device("cuda", torch.cuda.current_device())


def stream_ptr():
    return torch.cuda.current_stream().cuda_stream


def ptr(t):
    return None if t is None else t.data_ptr()


def to_device(x, dtype=None):
    """ndarray / Tensor (host or device) -> contiguous CUDA tensor (dtype preserved unless given).
    Never mutates or aliases a caller-owned host array."""
    dev = device()
    if isinstance(x, torch.Tensor):
        t = x.detach()
    else:
        a = np.asarray(x)
        if a.dtype == np.float16 or a.dtype.kind in "iub":
            a = a.astype(np.float32)
        t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    if t.dtype not in (torch.float32, torch.float64, torch.int32, torch.int64):
        t = t.to(torch.float32)
    return t.to(dev, non_blocking=False).contiguous()


def as_f32_rows(x, center=None):
    """Stages a [N, d] matrix as float32 on the device.  float64 inputs have `center` (float64
    [d] CUDA tensor or None) subtracted at input precision before the cast
    (runia_center_cast); float32 inputs are returned as they are (the kernels subtract the
    float32 centre in their prologue, like the reference's float32 arithmetic).
    Returns (tensor_f32, centered: bool)."""
    from . import _lib

    t = to_device(x)
    if t.dtype == torch.float32:
        return t, False
    if t.dtype != torch.float64:
        return t.to(torch.float32), False
    out = torch.empty(t.shape, dtype=torch.float32, device=t.device)
    n, d = t.shape
    _lib.call("runia_center_cast", t.data_ptr(), 1, n, d, ptr(center), out.data_ptr(), stream_ptr())
    return out, center is not None


def to_host(t):
    return t.cpu().numpy()
