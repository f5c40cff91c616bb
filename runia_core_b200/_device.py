"""Device staging helpers: PyTorch is used for device memory, streams and torch.distributed
only -- all arithmetic happens in libruniab200.so.

Host -> device staging.  The reference's callers hand NumPy arrays (pageable memory) to `postprocess()`
(`runia_core/evaluation/metrics.py:322-340`).  A plain `tensor.to(device)` of pageable memory is one synchronous
`cudaMemcpy` that the driver stages through its own bounce buffer at 10-16 GB/s; `HostPipe` does the staging
itself: a ring of pinned slots, worker threads that `memcpy` slabs of the source into a slot (NumPy releases the GIL
for plain copies), one `cudaMemcpyAsync` per slot on a copy stream, so that the CPU copy of chunk k+1, the PCIe
transfer of chunk k and -- for the row scorers that ask for it -- the kernel on chunk k-1 overlap.  Pinned sources
skip the CPU copy.  Small arrays (< 1 MiB) take the plain path: the pipeline's fixed cost is not worth it."""
import os
import threading
from concurrent.futures import ThreadPoolExecutor

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


PIPE_MIN_BYTES = 1 << 20
PIPE_SLOT_BYTES = int(os.environ.get("RUNIA_B200_PIPE_SLOT_BYTES", 8 << 20))
PIPE_SLOTS = 4
PIPE_THREADS = max(1, min(8, (os.cpu_count() or 2) // 2))


class HostPipe:
    """Pinned staging ring of one device (see the module docstring).  One instance per device, serialised by a
    lock: calls from several host threads queue up rather than interleave their slots."""

    _instances = {}
    _guard = threading.Lock()

    @classmethod
    def get(cls, dev: torch.device) -> "HostPipe":
        with cls._guard:
            p = cls._instances.get(dev.index)
            if p is None:
                p = cls._instances[dev.index] = HostPipe(dev)
            return p

    def __init__(self, dev: torch.device):
        self.dev = dev
        self.lock = threading.Lock()
        self.slots = [torch.empty(PIPE_SLOT_BYTES, dtype=torch.uint8).pin_memory() for _ in range(PIPE_SLOTS)]
        self.slot_np = [s.numpy() for s in self.slots]
        self.slot_free = [None] * PIPE_SLOTS  # event: the H2D copy out of the slot has completed
        self.copy_stream = torch.cuda.Stream(dev)
        self.pool = ThreadPoolExecutor(max_workers=PIPE_THREADS, thread_name_prefix="runia-h2d")
        self.out_slot = None  # pinned landing buffer for results (grown on demand)

    def _fill(self, slot, src_u8, lo, hi):
        """memcpy src_u8[lo:hi] into the slot, split over the worker threads."""
        n = hi - lo
        dst = self.slot_np[slot]
        if n < (1 << 20) or PIPE_THREADS == 1:
            np.copyto(dst[:n], src_u8[lo:hi])
            return
        cuts = [lo + (n * t) // PIPE_THREADS // 64 * 64 for t in range(PIPE_THREADS)] + [hi]
        list(self.pool.map(lambda t: np.copyto(dst[cuts[t] - lo:cuts[t + 1] - lo], src_u8[cuts[t]:cuts[t + 1]]),
                           range(PIPE_THREADS)))

    def upload(self, src: np.ndarray, dst: torch.Tensor, row_bytes: int, on_rows=None):
        """Copies the C-contiguous host array `src` into the device tensor `dst` (same dtype and shape) chunk by
        chunk.  `on_rows(lo_row, hi_row)`, if given, is called on the caller's current stream after that stream has
        been made to wait for rows [lo_row, hi_row) -- the kernel of a row scorer on that chunk.  Without `on_rows`
        the current stream waits for the whole copy before the call returns (the host does not block)."""
        src_u8 = src.reshape(-1).view(np.uint8)
        dst_u8 = dst.reshape(-1).view(torch.uint8)
        total = src_u8.shape[0]
        rows_per_slot = max(1, PIPE_SLOT_BYTES // row_bytes)
        if rows_per_slot * row_bytes > PIPE_SLOT_BYTES:  # a single row wider than a slot: byte chunks, no callback
            rows_per_slot, row_bytes, on_rows = PIPE_SLOT_BYTES, 1, None
        chunk = rows_per_slot * row_bytes
        cur = torch.cuda.current_stream(self.dev)
        with self.lock:
            self.copy_stream.wait_stream(cur)  # dst may still be in use by earlier work on the caller's stream
            k = 0
            for lo in range(0, total, chunk):
                hi = min(total, lo + chunk)
                s = k % PIPE_SLOTS
                if self.slot_free[s] is not None:
                    self.slot_free[s].synchronize()
                self._fill(s, src_u8, lo, hi)
                with torch.cuda.stream(self.copy_stream):
                    dst_u8[lo:hi].copy_(self.slots[s][: hi - lo], non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(self.copy_stream)
                self.slot_free[s] = ev
                if on_rows is not None:
                    cur.wait_event(ev)
                    on_rows(lo // row_bytes, hi // row_bytes)
                k += 1
            if on_rows is None:
                cur.wait_stream(self.copy_stream)

    def upload_pinned(self, src: torch.Tensor, dst: torch.Tensor, row_bytes: int, on_rows):
        """Pinned host tensor -> device in chunks with the per-chunk callback (no CPU copy needed)."""
        src_u8 = src.reshape(-1).view(torch.uint8)
        dst_u8 = dst.reshape(-1).view(torch.uint8)
        total = src_u8.shape[0]
        chunk = max(1, PIPE_SLOT_BYTES // row_bytes) * row_bytes
        cur = torch.cuda.current_stream(self.dev)
        with self.lock:
            self.copy_stream.wait_stream(cur)
            for lo in range(0, total, chunk):
                hi = min(total, lo + chunk)
                with torch.cuda.stream(self.copy_stream):
                    dst_u8[lo:hi].copy_(src_u8[lo:hi], non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(self.copy_stream)
                cur.wait_event(ev)
                on_rows(lo // row_bytes, hi // row_bytes)

    def download(self, t: torch.Tensor) -> np.ndarray:
        """Device tensor -> fresh NumPy array through the pinned landing buffer (one async copy + one sync)."""
        nbytes = t.numel() * t.element_size()
        with self.lock:
            if self.out_slot is None or self.out_slot.numel() < nbytes:
                self.out_slot = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8).pin_memory()
            land = self.out_slot[:nbytes].view(t.dtype).reshape(t.shape)
            land.copy_(t, non_blocking=True)
            torch.cuda.current_stream(self.dev).synchronize()
            return land.numpy().copy()


_TORCH_DTYPE = {np.float32: torch.float32, np.float64: torch.float64, np.int32: torch.int32, np.int64: torch.int64}


def _host_array(x):
    """ndarray (or CPU tensor) -> C-contiguous ndarray of a dtype the kernels take, or None for device tensors."""
    if isinstance(x, torch.Tensor):
        if x.is_cuda:
            return None
        x = x.detach()
        if x.dtype in (torch.bfloat16, torch.float16):
            x = x.float()
        a = x.numpy()
    else:
        a = np.asarray(x)
    if a.dtype == np.float16 or a.dtype.kind in "iub":
        a = a.astype(np.float32)
    elif a.dtype not in (np.float32, np.float64, np.int32, np.int64):
        a = a.astype(np.float32)
    return np.ascontiguousarray(a)


def to_device(x, dtype=None):
    """ndarray / Tensor (host or device) -> contiguous CUDA tensor (dtype preserved unless given).
    Never mutates or aliases a caller-owned host array.  Large host arrays go through the pinned staging ring."""
    dev = device()
    if isinstance(x, torch.Tensor) and x.is_cuda:
        t = x.detach()
        if dtype is not None and t.dtype != dtype:
            t = t.to(dtype)
        if t.dtype not in (torch.float32, torch.float64, torch.int32, torch.int64):
            t = t.to(torch.float32)
        return t.to(dev).contiguous()
    pinned = isinstance(x, torch.Tensor) and x.is_pinned()
    if pinned:
        t = x.detach()
        if t.dtype not in (torch.float32, torch.float64, torch.int32, torch.int64):
            t = t.to(torch.float32)
        t = t.to(dev, non_blocking=True).contiguous()
    else:
        a = _host_array(x)
        nbytes = a.nbytes
        if nbytes >= PIPE_MIN_BYTES:
            t = torch.empty(a.shape, dtype=_TORCH_DTYPE[a.dtype.type], device=dev)
            HostPipe.get(dev).upload(a, t, row_bytes=max(1, nbytes // max(1, a.shape[0])))
        else:
            t = torch.from_numpy(a).to(dev, non_blocking=False)
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t


def stream_rows(x, fn, min_rows: int = 4096):
    """Row scorer over a HOST matrix with copy / compute overlap: `fn(chunk)` is called with device chunks
    x_dev[lo:hi] (float32 or float64 rows as they are on the host) as they arrive and must enqueue its kernel on
    the current stream.  Returns False (nothing done) when `x` is not a large 2-D float host array -- the caller
    then takes the plain path."""
    if isinstance(x, torch.Tensor) and x.is_cuda:
        return False
    dev = device()
    pinned = isinstance(x, torch.Tensor) and x.is_pinned() and x.is_contiguous() and \
        x.dtype in (torch.float32, torch.float64)
    a = x if pinned else _host_array(x)
    if a.ndim != 2 or a.shape[0] < min_rows or (not pinned and a.dtype not in (np.float32, np.float64)):
        return False
    row_bytes = a.shape[1] * (a.element_size() if pinned else a.itemsize)
    if a.shape[0] * row_bytes < 4 * PIPE_SLOT_BYTES or row_bytes > PIPE_SLOT_BYTES:
        return False
    tdt = a.dtype if pinned else (torch.float32 if a.dtype == np.float32 else torch.float64)
    xd = torch.empty(tuple(a.shape), dtype=tdt, device=dev)
    pipe = HostPipe.get(dev)
    cb = lambda lo, hi: fn(xd[lo:hi], lo, hi)  # noqa: E731
    if pinned:
        pipe.upload_pinned(a, xd, row_bytes, cb)
    else:
        pipe.upload(a, xd, row_bytes, cb)
    return True


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
    """Device tensor -> NumPy array.  Results above 256 KiB land in a pinned buffer (async copy, one sync)."""
    if t.is_cuda and t.numel() * t.element_size() >= (256 << 10):
        return HostPipe.get(t.device).download(t.contiguous())
    return t.cpu().numpy()
