"""Device staging helpers: PyTorch is used for device memory, streams and torch.distributed
only -- all arithmetic happens in libruniab200.so.

Host -> device staging.  The reference's callers hand NumPy arrays (pageable memory) to `postprocess()`
(`runia_core/evaluation/metrics.py:322-340`).  A plain `tensor.to(device)` of pageable memory is one synchronous
`cudaMemcpy` that the driver stages through its own bounce buffer at 10-16 GB/s; `HostPipe` does the staging
itself: a ring of pinned slots, worker threads that `memcpy` slabs of the source into a slot (NumPy releases the GIL
for plain copies), one `cudaMemcpyAsync` per slot on a copy stream, so that the CPU copy of chunk k+1, the PCIe
transfer of chunk k and -- for the row scorers that ask for it -- the kernel on chunk k-1 overlap.  Pinned sources
skip the CPU copy.  Small arrays (< 256 KiB) take the plain path.  The engine itself is native (csrc/stage.cu)."""
import os
import threading

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


def bind_to_gpu_numa_node(dev_index: int = None) -> dict:
    """Pins the calling process (its future threads included: the staging workers are created on first use) to the CPUs
    of the NUMA node the GPU hangs off, so that pinned slots and staging copies stay node-local when several ranks share
    a host.  Returns what it found ({"numa_node": -1} on hosts that expose no topology, e.g. single-node VMs: no-op)."""
    require_cuda()
    dev_index = torch.cuda.current_device() if dev_index is None else dev_index
    info = {"numa_node": -1, "cpus": len(os.sched_getaffinity(0))}
    try:
        pr = torch.cuda.get_device_properties(dev_index)
        addr = f"{getattr(pr, 'pci_domain_id', 0):04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{addr}/numa_node") as f:
            node = int(f.read().strip())
        info["pci"] = addr
        if node < 0:
            return info
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            info.update(numa_node=node, cpus=len(cpus))
    except Exception as e:  # noqa: BLE001 -- topology files are optional
        info["error"] = repr(e)
    return info


PIPE_MIN_BYTES = 256 << 10                                                    # below this a plain copy is as fast
PIPE_CHUNK_BYTES = int(os.environ.get("RUNIA_B200_PIPE_CHUNK_BYTES", 64 << 20))  # rows per kernel launch when streaming


class HostPipe:
    """Per-device host <-> device staging.  Uploads of pageable memory go through the native staging engine
    (`runia_stage_h2d`, csrc/stage.cu: pinned ring, parallel memcpy by a worker pool, one cudaMemcpyAsync per 4 MiB);
    `stream_rows` drives it chunk by chunk on a copy stream so that a row scorer's kernel on chunk k overlaps the
    transfer of chunk k + 1; results come back through a pinned landing buffer."""

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
        self.copy_stream = torch.cuda.Stream(dev)
        self.out_slot = None  # pinned landing buffer for results (grown on demand)

    @staticmethod
    def _h2d(dst_ptr: int, src_ptr: int, nbytes: int, stream: torch.cuda.Stream):
        from . import _lib

        _lib.call("runia_stage_h2d", dst_ptr, src_ptr, nbytes, stream.cuda_stream)

    def upload(self, src: np.ndarray, dst: torch.Tensor):
        """C-contiguous host array -> device tensor of the same dtype and shape, on the caller's current stream
        (ready in stream order; the host returns as soon as the last chunk is enqueued)."""
        self._h2d(dst.data_ptr(), src.ctypes.data, src.nbytes, torch.cuda.current_stream(self.dev))

    def upload_rows(self, src, dst: torch.Tensor, row_bytes: int, on_rows):
        """Row-chunked upload on the copy stream; after each chunk the caller's stream waits for it and
        `on_rows(lo, hi)` enqueues that chunk's kernel.  `src`: C-contiguous ndarray or pinned CPU tensor."""
        pinned = isinstance(src, torch.Tensor)
        n_rows = dst.shape[0]
        rows = max(1, PIPE_CHUNK_BYTES // row_bytes)
        cur = torch.cuda.current_stream(self.dev)
        src_ptr = src.data_ptr() if pinned else src.ctypes.data
        with self.lock:
            self.copy_stream.wait_stream(cur)  # dst was allocated on (and may be reused by) the caller's stream
            for lo in range(0, n_rows, rows):
                hi = min(n_rows, lo + rows)
                if pinned:
                    with torch.cuda.stream(self.copy_stream):
                        dst[lo:hi].copy_(src[lo:hi], non_blocking=True)
                else:
                    self._h2d(dst.data_ptr() + lo * row_bytes, src_ptr + lo * row_bytes, (hi - lo) * row_bytes,
                              self.copy_stream)
                ev = torch.cuda.Event()
                ev.record(self.copy_stream)
                cur.wait_event(ev)
                on_rows(lo, hi)

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
            HostPipe.get(dev).upload(a, t)
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
    if a.shape[0] * row_bytes < 2 * PIPE_CHUNK_BYTES:
        return False
    tdt = a.dtype if pinned else (torch.float32 if a.dtype == np.float32 else torch.float64)
    xd = torch.empty(tuple(a.shape), dtype=tdt, device=dev)
    HostPipe.get(dev).upload_rows(a, xd, row_bytes, lambda lo, hi: fn(xd[lo:hi], lo, hi))
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
