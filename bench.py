#!/usr/bin/env python
"""Benchmark of the OoD scoring hot path (BASELINE.json metric: OoD-scored embeddings/sec).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload larem]

One "step" = one pass of the LaREM (Mahalanobis, d=256 after PCA) scorer over one batch of
synthetic embeddings that is larger than L2.  `value` is measured with inputs resident in HBM
(CUDA events, max over ranks); `e2e` goes through the reference-facing class with HOST buffers
(pinned H2D copy + score + D2H read of the scores inside the timed region).  Secondary numbers
(kNN with k=50 on a 50k bank = BASELINE config 2; sharded-bank kNN at N>1; the entropy kernel)
ride along under "extra", each with its own roofline.  `--impl reference` times the oracle's
faithful port of the reference's CPU path on the host cores (the reference is pure Python on
NumPy/sklearn; it cannot be pip-installed offline because of its unmet dependencies, see
DESIGN.md).
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from bench_configs import _tensor as _tensor_roofline  # noqa: E402  (tensor-bound roofline against this run's TF32 probe)

D_LATENT = 256
N_TRAIN = 50_000
WORKLOAD = ("LaREM (MDLatentSpace) scoring, d=256 after PCA, fit on 50k train latents "
            "(BASELINE configs[1] bank size; configs[0] scorer)")


def _config(rows_per_gpu, world):
    """The `config` object of BOTH arms (the reference arm times bounded samples of this workload)."""
    return {"workload": WORKLOAD, "rows_per_gpu": rows_per_gpu, "d": D_LATENT,
            "input_bytes_per_gpu": rows_per_gpu * D_LATENT * 4,
            "l2_policy": "inputs (4.3 GB) larger than L2 (126 MB)", "parallelism": f"rows x{world}"}


_REAL_STDOUT = None


def _claim_stdout():
    """stdout carries exactly one JSON line: everything libraries print to fd 1 while the bench runs
    (the NCCL version banner, for one) is sent to stderr instead."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def _emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def _peaks():
    """(HBM GB/s, dense bf16 TFLOP/s burst, source).  MEASURED_PEAKS.json is driver-written."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            j = json.load(f)
        return float(j["hbm_gbs"]), float(j["bf16_tflops"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1590.0, "fallback (B200_PROFILING.md)"


def _traffic(kernel):
    """DRAM bytes per launch of `kernel` at the bench shape, from the committed ncu --set full
    capture (profiles/traffic.json: dram__bytes_read.sum + dram__bytes_write.sum), or None."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f).get(kernel)
    return None


class ClockSampler:
    """SM clock and clock-event (throttle) reasons of one GPU while the timed region runs.  Reads NVML -- the
    library behind nvidia-smi's clocks.sm / clocks.max.sm / clocks_event_reasons.* fields -- from a thread every
    2 ms, so that a 40 ms timed region holds ~20 samples taken INSIDE it (an `nvidia-smi -lms` child needs longer than
    that to print its first line); falls back to the nvidia-smi query of the profiling recipe when pynvml is missing."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    BITS = {0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown",
            0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown"}

    def __init__(self, gpu_index, uuid=None):
        self.gpu_index = gpu_index
        self.uuid = uuid
        self.proc = None
        self.lines = []
        self.samples = []  # (host time, sm MHz, reasons bitmask)
        self.nvml = None
        self._stop = False

    def start(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            h = None
            if self.uuid:
                for cand in (f"GPU-{self.uuid}", str(self.uuid)):
                    try:
                        h = pynvml.nvmlDeviceGetHandleByUUID(cand)
                        break
                    except Exception:
                        h = None
            if h is None:
                h = pynvml.nvmlDeviceGetHandleByIndex(self.gpu_index)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            self.nvml = (pynvml, h)
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                 "-i", str(self.gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _poll(self):
        pynvml, h = self.nvml
        while not self._stop:
            try:
                self.samples.append((time.time(), float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)),
                                     int(pynvml.nvmlDeviceGetCurrentClocksEventReasons(h))))
            except Exception:
                pass
            time.sleep(0.002)

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self, window=None):
        """window = (t0, t1) host times of the timed region: only samples inside it count (NVML path)."""
        if self.nvml is not None:
            self._stop = True
            self.thread.join(timeout=1.0)
            inside = [x for x in self.samples if window is None or window[0] <= x[0] <= window[1]]
            if not inside:
                inside = self.samples[-3:]
            mask = 0
            for _, _, m in inside:
                mask |= m
            reasons = sorted(name for bit, name in self.BITS.items() if mask & bit)
            return {"sm_mhz": float(np.median([x[1] for x in inside])) if inside else None, "sm_max_mhz": self.sm_max,
                    "reasons": reasons, "samples": len(inside), "source": "nvml, every 2 ms inside the timed region",
                    "reasons_mask": hex(mask)}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi -lms 20"}


def _fit_larem(seed=1):
    """LaREM fit on 50k x 256 whitened latents (what PCA-256 of config 1/2 features looks like)."""
    rng = np.random.RandomState(seed)
    train = (rng.randn(N_TRAIN, D_LATENT)).astype(np.float32)
    train += (0.05 * rng.randn(1, D_LATENT)).astype(np.float32)
    return train


def _time_events(fn, steps, warmup, dist_barrier):
    import torch

    for _ in range(warmup):
        fn()
    dist_barrier()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    ev[0].record()
    for i in range(steps):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    dist_barrier()
    per = [ev[i].elapsed_time(ev[i + 1]) for i in range(steps)]
    return ev[0].elapsed_time(ev[steps]), per


def _all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1; the CPU arms are supposed to use every host core."""
    n = os.cpu_count() or 1
    try:
        import threadpoolctl

        threadpoolctl.threadpool_limits(limits=n)
        return max([p.get("num_threads", 1) for p in threadpoolctl.threadpool_info()] or [1])
    except Exception:
        return n


def run_b200(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        # the ranks share the host's cores: give each rank's staging engine its share (read once, at first use)
        os.environ.setdefault("RUNIA_B200_STAGE_THREADS", str(max(2, min(8, (os.cpu_count() or 8) // world))))

    def barrier():
        if world > 1:
            dist.barrier()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    import runia_core_b200 as R
    from runia_core_b200 import _device, _lib, _ops

    # several ranks on one host: keep each rank's staging threads and pinned slots on its GPU's NUMA node
    numa = _device.bind_to_gpu_numa_node(local) if world > 1 else {"numa_node": None}
    hbm_peak, bf16_peak, peak_src = _peaks()
    dev = torch.device("cuda", local)
    md = R.inference.MDLatentSpace()
    md.setup(_fit_larem())
    st = md._state

    # ---------------- device-resident LaREM: [N, 256] f32 per rank, 4.3 GB > L2 ----------------
    n_rows = args.rows
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    X = torch.randn(n_rows, D_LATENT, generator=g, device=dev, dtype=torch.float32)
    X[n_rows // 2:] -= 0.5  # OoD-like half
    out_holder = {}

    def step():
        out_holder["s"] = _ops.md_score(X, st, torch.float64)

    # TF32 tensor peak of this GPU: one reading before the timed region (GPU not yet heated), one after it (clocks up)
    tf32_probe = _tf32_probe(torch, _lib) if rank == 0 else None
    try:
        uuid = str(torch.cuda.get_device_properties(local).uuid)
    except Exception:
        uuid = None
    sampler = ClockSampler(local, uuid)
    sampler.start()
    l0 = _lib.launch_count()
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    l1 = _lib.launch_count()
    t_w0 = time.time()
    total_ms, per = _time_events(step, args.steps, 0, barrier)
    clocks = sampler.stop(window=(t_w0, time.time()))
    launches = _lib.launch_count() - l1
    total_ms = max_over_ranks(total_ms)
    ms_per_step = total_ms / args.steps
    value = n_rows * world / (ms_per_step * 1e-3)
    kern_ms = float(np.mean(per))
    # 1,032 algorithmic bytes and 2 d^2 + 3 d = 131,840 FLOP per embedding at d = 256
    # (128 FLOP/B).  FP32-faithful on the tensor cores = three TF32 products per FLOP, so the
    # kernel is tensor-bound: ceiling = TF32 peak / 3, TF32 peak = measured dense bf16 / 2 (no TF32
    # figure is measured; TF32 runs at half the bf16 rate on this part).
    alg_bytes = n_rows * (D_LATENT * 4 + 8)
    hbm_achieved = alg_bytes / (kern_ms * 1e-3) / 1e9
    flops = n_rows * (2.0 * D_LATENT * st.r + 3.0 * D_LATENT)
    tflops = flops / (kern_ms * 1e-3) / 1e12
    tc = _ops.get_engine() == "tc"
    peak_tf = bf16_peak / 2.0 / 3.0
    if rank == 0:  # and once more on the warm GPU: the burst peak is the larger of the two readings
        tf32_probe = max(tf32_probe or 0.0, _tf32_probe(torch, _lib))
    # denominator: the TF32 rate this GPU sustains for the kernel's own MMA instruction, measured in this run
    # (runia_tf32_peak_probe), / 3 products; MEASURED_PEAKS.json has no TF32 entry, and the figure derived from its
    # dense bf16 number (/ 2 / 3) is lower than what the tensor pipe delivers (the kernel exceeds it), so it is
    # reported beside, not used
    derived_tf = peak_tf
    if tf32_probe:
        peak_tf = tf32_probe / 3.0
    if world > 1:  # every rank must divide by the same number
        t = torch.tensor([peak_tf], dtype=torch.float64, device=dev)
        dist.broadcast(t, 0)
        peak_tf = float(t.item())
    roofline = {"bound": "tensor", "achieved": round(tflops, 2), "peak": round(peak_tf, 2), "unit": "TFLOP/s",
                "frac": round(tflops / peak_tf, 4), "traffic": _traffic("tc_kernel<tc::RowNormEpi>") if n_rows == 4 * 1024 * 1024 else None,
                "peak_source": ("measured in this run: runia_tf32_peak_probe (back-to-back tcgen05.mma kind::tf32 256x256x8 from "
                                "resident tiles) / 3 (3xTF32 products per FP32-faithful FLOP)") if tf32_probe else
                               (peak_src + ": dense bf16 burst / 2 (TF32 rate) / 3 (3xTF32 products per FP32-faithful FLOP)"),
                "peak_from_measured_bf16": {"peak": round(derived_tf, 2), "frac": round(tflops / derived_tf, 4),
                                            "source": peak_src + ": dense bf16 burst / 2 / 3"},
                "kernel": ("tc_kernel<RowNormEpi> (tcgen05 cta_group::2 3xTF32 contraction, TMEM row sum of squares)"
                           if tc else "rownorm_kernel (FP32 SIMT contraction)"),
                "algorithmic_flop_per_embedding": 2 * D_LATENT * st.r + 3 * D_LATENT,
                "algorithmic_bytes_per_embedding": D_LATENT * 4 + 8,
                "tensor_tf32_tflops_issued": round(3 * tflops, 2),
                # the same MMA instruction issued back to back from resident tiles, timed in this run: what the
                # tensor pipe of this GPU sustains in TF32 (MEASURED_PEAKS.json has no TF32 entry)
                "tf32_probe": None if not tf32_probe else {"tflops": round(tf32_probe, 1),
                                                           "frac_of_probe": round(3 * tflops / tf32_probe, 4)},
                "hbm": {"achieved": round(hbm_achieved, 1), "peak": hbm_peak, "unit": "GB/s",
                        "frac": round(hbm_achieved / hbm_peak, 4)}}

    # ---------------- the same launch back to back for >= 2 s: the sustained figure beside the 40 ms burst ---------
    sustained = None
    if rank == 0 and not args.no_extra:
        n_launch = int(max(200, math.ceil(2200.0 / max(kern_ms, 1e-3))))
        smp = ClockSampler(local, uuid)
        smp.start()
        t_s0 = time.time()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n_launch):
            step()
        e1.record()
        torch.cuda.synchronize()
        ck = smp.stop(window=(t_s0, time.time()))
        ms_s = e0.elapsed_time(e1) / n_launch
        tf_s = flops / (ms_s * 1e-3) / 1e12
        sustained = {"launches": n_launch, "seconds": round(ms_s * n_launch * 1e-3, 2), "ms_per_step": ms_s,
                     "embeddings_per_s": n_rows / (ms_s * 1e-3), "fp32_equiv_tflops": round(tf_s, 2),
                     "frac_of_probe_over_3": round(tf_s / peak_tf, 4), "clocks": ck}

    # ---------------- end to end through the reference-facing class, host buffers ----------------
    # `md.postprocess(host array)` = what evaluation/metrics.py:331-340 calls: H2D of the rows, kernel, D2H of the
    # float64 scores, all inside the timed region.  Headline: a pageable NumPy array (what the reference API is
    # handed) at --e2e-rows; variants: the same from a pinned tensor, and both at 10,000 rows -- the size of one
    # reference call (the reference arm's step) -- so that the two arms can be compared at the same call size.
    def e2e_measure(src, reps):
        for _ in range(3):
            md.postprocess(src)
        barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            md.postprocess(src)
        torch.cuda.synchronize()
        return max_over_ranks((time.perf_counter() - t0) / reps * 1e3) * 1e-3

    n_e2e = min(args.e2e_rows, n_rows)
    e2e_steps = max(3, min(args.steps, 10))
    variants = {}
    for rows_v, reps in ((n_e2e, e2e_steps), (10_000, 50)):
        host_np = X[:rows_v].cpu().numpy()
        pinned = torch.empty((rows_v, D_LATENT), dtype=torch.float32).pin_memory()
        pinned.copy_(torch.from_numpy(host_np))
        for name, src in (("pageable_ndarray", host_np), ("pinned_tensor", pinned)):
            sec = e2e_measure(src, reps)
            variants[f"{name}_{rows_v}"] = {"embeddings_per_s": rows_v * world / sec, "ms_per_call": sec * 1e3,
                                            "rows_per_call": rows_v,
                                            "h2d_GBps_per_gpu": rows_v * D_LATENT * 4 / sec / 1e9}
        del host_np, pinned
    head = variants[f"pageable_ndarray_{n_e2e}"]
    e2e = {"value": head["embeddings_per_s"], "unit": "embeddings/s", "h2d_bytes_per_step": n_e2e * D_LATENT * 4,
           "d2h_bytes_per_step": n_e2e * 8, "rows_per_step": n_e2e,
           "source": "pageable numpy.ndarray -> MDLatentSpace.postprocess -> numpy.ndarray (pinned staging ring inside)",
           "h2d_GBps_all_gpus": head["h2d_GBps_per_gpu"] * world, "host_cpus": os.cpu_count(), "numa": numa,
           "variants": variants}

    extra = {}
    if sustained is not None:
        extra["larem_sustained"] = sustained
    if rank == 0:
        extra["tf32_probe_tflops"] = tf32_probe
    del X, out_holder
    torch.cuda.empty_cache()
    if rank == 0 and not args.no_extra:
        extra.update(_extra_single_gpu(args, torch, R, _ops, _lib, hbm_peak, tf32_probe, bf16_peak))
        extra.update(_extra_other_kernels(torch, _ops, hbm_peak, bf16_peak, tf32_probe))
        if world == 1 and not args.no_sweep:
            extra["sweep_config2"] = _extra_sweep_config2(R)
    if not args.no_extra and not args.no_configs:
        import bench_configs as BC
        from runia_core_b200 import sharding

        probe_all = peak_tf * 3.0  # every rank divides by the same probe (broadcast above)
        if world in (1, 2, 4, 8):
            extra["config4"] = BC.config4(torch, dist, _ops, world, rank, barrier, max_over_ranks, probe_all, bf16_peak,
                                          nq=args.c4_queries)
        extra["kde_sharded"] = BC.kde_sharded(torch, dist, _ops, sharding, world, rank, barrier, max_over_ranks,
                                              probe_all, bf16_peak)
        if world == 1:
            extra["config3"] = BC.config3(torch, R, _ops, hbm_peak, probe_all, bf16_peak, n_boxes=args.c3_boxes)
            extra["config5"] = BC.config5(torch, R, _ops, md, hbm_peak, probe_all, bf16_peak)

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu_baseline = _cpu_larem(md, seconds=args.cpu_seconds)
        if not args.no_extra:
            extra["cpu_baselines"] = _cpu_other_rows()

    if rank == 0:
        line = {
            "metric": "ood_scored_embeddings_per_sec", "value": value, "unit": "embeddings/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": _config(n_rows, world),
            "roofline": roofline, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "cpu_baseline": cpu_baseline, "extra": extra,
        }
        _emit(line)
    if world > 1:
        dist.destroy_process_group()


def _extra_single_gpu(args, torch, R, _ops, _lib, hbm_peak, tf32_probe=None, bf16_peak=1630.7):
    """Secondary numbers on rank 0: kNN (config 2) and the entropy kernel (config 1 shape)."""
    out = {}
    dev = torch.device("cuda", torch.cuda.current_device())
    g = torch.Generator(device=dev).manual_seed(7)
    # kNN: 50k x 512 bank, 10k queries, k = 50
    centers = torch.randn(10, 512, generator=g, device=dev)
    lab = torch.randint(0, 10, (50_000,), generator=g, device=dev)
    bank = _ops.normalize_rows(centers[lab] + torch.randn(50_000, 512, generator=g, device=dev))
    labq = torch.randint(0, 10, (10_000,), generator=g, device=dev)
    q = _ops.normalize_rows(centers[labq] + torch.randn(10_000, 512, generator=g, device=dev))
    kb = _ops.knn_bank(bank)
    res = {}

    def knn_step():
        res["r"] = _ops.knn_search(q, kb, 50, want_idx=False, want_dist=False, check_status=False)

    for _ in range(2):
        knn_step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record()
    for _ in range(reps):
        knn_step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    fl = 2.0 * 10_000 * 50_000 * 512
    full = _ops.knn_search(q, kb, 50)
    out["knn_config2"] = {"queries_per_s": 10_000 / (ms * 1e-3), "ms": ms, "bank": [50_000, 512], "k": 50,
                          "distance_tflops": fl / (ms * 1e-3) / 1e12,
                          "roofline": _tensor_roofline(fl, ms, tf32_probe, bf16_peak, products=1),
                          "exhaustive_rows": full["exhaustive_rows"],
                          "filter_tf32_products": _ops.knn_filter_products(kb, 50)}
    # entropy: 16 MC samples x 512 dims, 60k items = 1.97 GB
    n_items, n_mc, D = 60_000, 16, 512
    z = torch.randn(n_items, 1, D, generator=g, device=dev) + 0.1 * torch.randn(n_items, n_mc, D, generator=g, device=dev)
    z = z.reshape(n_items * n_mc, D).contiguous()

    def ent_step():
        res["e"] = _ops.mcd_entropy(z, n_mc)

    for _ in range(2):
        ent_step()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        ent_step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    alg = n_items * (n_mc * D * 4 + D * 8 + 8)
    out["entropy_config1"] = {"items_per_s": n_items / (ms * 1e-3), "ms": ms, "n_mc": n_mc, "D": D,
                              "roofline": {"bound": "hbm", "achieved": alg / (ms * 1e-3) / 1e9, "peak": hbm_peak,
                                           "unit": "GB/s", "frac": alg / (ms * 1e-3) / 1e9 / hbm_peak},
                              "alu_pipe_roofline": _alu_pipe_roofline(n_items * (D / 64.0) * 540.0, ms)}
    del z
    # PCA 512 -> 256 projection
    from sklearn.decomposition import PCA
    rng = np.random.RandomState(3)
    pca = PCA(n_components=256, whiten=True)
    pca.mean_ = rng.randn(512)
    pca.components_ = np.linalg.qr(rng.randn(512, 256))[0].T.copy()
    pca.explained_variance_ = 1.0 + rng.rand(256)
    stp = _ops.pca_prepare(pca.mean_, pca.components_, pca.explained_variance_, True)
    Xp = torch.randn(2_000_000, 512, generator=g, device=dev)

    def pca_step():
        res["p"] = _ops.pca_transform(Xp, stp)

    for _ in range(2):
        pca_step()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        pca_step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    alg = 2_000_000 * (512 * 4 + 256 * 4)
    out["pca_512_256"] = {"embeddings_per_s": 2_000_000 / (ms * 1e-3), "ms": ms,
                          "roofline": {"bound": "hbm", "achieved": alg / (ms * 1e-3) / 1e9, "peak": hbm_peak,
                                       "unit": "GB/s", "frac": alg / (ms * 1e-3) / 1e9 / hbm_peak},
                          "fp32_tflops": 2.0 * 2_000_000 * 512 * 256 / (ms * 1e-3) / 1e12,
                          "tensor_roofline": _tensor_roofline(2.0 * 2_000_000 * 512 * 256, ms, tf32_probe, bf16_peak)}
    # entropy at the reference's DEFAULT mcd_samples_nro = 32 (evaluation/entropy.py:41), D = 512
    n_items, n_mc = 30_000, 32
    z = torch.randn(n_items, 1, D, generator=g, device=dev) + 0.1 * torch.randn(n_items, n_mc, D, generator=g, device=dev)
    z = z.reshape(n_items * n_mc, D).contiguous()
    ms = _time_op(torch, lambda: _ops.mcd_entropy(z, n_mc))
    alg = n_items * (n_mc * D * 4 + D * 8 + 8)
    out["entropy_n32"] = {"items_per_s": n_items / (ms * 1e-3), "ms": ms, "n_mc": n_mc, "D": D,
                          "roofline": {"bound": "hbm", "achieved": alg / (ms * 1e-3) / 1e9, "peak": hbm_peak,
                                       "unit": "GB/s", "frac": alg / (ms * 1e-3) / 1e9 / hbm_peak},
                          "alu_pipe_roofline": _alu_pipe_roofline(n_items * (D / 128.0) * 4 * 856.0, ms)}
    del z
    return out


def _alu_pipe_roofline(alu_warp_instructions, ms, sm_mhz=1965.0):
    """ALU-pipe roofline of the entropy kernels, whose arithmetic is min / max: on this GPU every FMNMX / FMNMX3 (2- or
    3-input) holds the ALU pipe of its scheduler for two cycles (scripts/probes/pipe_probe.cu ->
    profiles/r2c_pipe_probe.jsonl: 2 FMNMX 4.0 cycles, 1 FMNMX3 2.0, 2 FMNMX + 2 FADD 4.25), so the peak is
    148 SMs x 4 schedulers x 0.5 instructions per cycle at the maximum SM clock.  `alu_warp_instructions` is counted from
    the shipped SASS (cuobjdump): 540 FMNMX + FMNMX3 per 64-dimension step of a warp at n_mc = 16
    (entropy16_kernel), 856 per 128-dimension tile and warp at n_mc = 32 (entropy32_kernel, four warps per tile)."""
    peak = 148 * 4 * 0.5 * sm_mhz * 1e6
    a = alu_warp_instructions / (ms * 1e-3)
    return {"bound": "alu pipe (min/max)", "achieved": round(a / 1e9, 1), "peak": round(peak / 1e9, 1),
            "unit": "G warp-instructions/s", "frac": round(a / peak, 4)}


def _time_op(torch, fn, reps=5, warm=2):
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


def _tf32_probe(torch, _lib):
    """TF32 TFLOP/s of back-to-back tcgen05.mma (256 x 256 x 8, cta_group::2) on resident tiles.  Burst figure, like
    MEASURED_PEAKS.json's: the best of eight separately timed ~0.5 ms launches (a multi-millisecond burn of pure MMAs
    runs into the power cap and reads 25 % lower than the scorers themselves sustain)."""
    import ctypes

    flop = ctypes.c_double(0.0)
    iters = 2048
    stream = torch.cuda.current_stream().cuda_stream
    _lib.call("runia_tf32_peak_probe", iters, ctypes.byref(flop), stream)
    torch.cuda.synchronize()
    best = 0.0
    for _ in range(8):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.call("runia_tf32_peak_probe", iters, ctypes.byref(flop), stream)
        e1.record()
        torch.cuda.synchronize()
        best = max(best, flop.value / (e0.elapsed_time(e1) * 1e-3) / 1e12)
        time.sleep(0.01)
    return best


def _extra_other_kernels(torch, _ops, hbm_peak, bf16_peak, tf32_probe=None):
    """The remaining rows of SURVEY 8(a), device-resident, each against the roofline SURVEY 8(d)
    names for it: logit scores / ReAct / ASH (HBM), ViM / class-conditional Mahalanobis / DDU / KDE
    (contractions; FP32-equivalent TFLOP/s against bf16/2/3 for the tensor-core ones)."""
    out = {}
    dev = torch.device("cuda", torch.cuda.current_device())
    g = torch.Generator(device=dev).manual_seed(21)
    rng = np.random.RandomState(21)
    hbm = lambda alg, ms: {"bound": "hbm", "achieved": alg / (ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",  # noqa: E731
                           "frac": alg / (ms * 1e-3) / 1e9 / hbm_peak}
    # (a8) Energy + MSP + GEN in one pass: 20M x 10 logits, 52 B/sample
    n = 20_000_000
    L = torch.randn(n, 10, generator=g, device=dev)
    ms = _time_op(torch, lambda: _ops.logit_scores(L, gamma=0.1, M=10))
    out["logit_scores_c10"] = {"samples_per_s": n / (ms * 1e-3), "ms": ms, "roofline": hbm(n * 52, ms)}
    del L
    # (a10) ReAct / DICE (clip -> linear -> LSE) and ASH-S: 2M x 512, 2,052 B/embedding
    n, d, C = 2_000_000, 512, 10
    X = torch.relu(torch.randn(n, d, generator=g, device=dev))
    W = 0.05 * torch.randn(C, d, generator=g, device=dev)
    b = torch.randn(C, generator=g, device=dev)
    planes = _ops.linear_planes(W)  # what ReAct.setup prepares once
    ms = _time_op(torch, lambda: _ops.clip_linear_lse(X, W, b, clip=1.0, planes=planes))
    out["react_512"] = {"embeddings_per_s": n / (ms * 1e-3), "ms": ms, "roofline": hbm(n * (d * 4 + 4), ms)}
    ms = _time_op(torch, lambda: _ops.ash_linear_lse(X, W, b, 77))
    out["ash_512"] = {"embeddings_per_s": n / (ms * 1e-3), "ms": ms, "roofline": hbm(n * (d * 4 + 4), ms)}
    # (a10) the same head at ImageNet size: C = 1000, d = 768 (256-column panels, online log-sum-exp across them)
    nI, dI, CI = 500_000, 768, 1000
    XI = torch.relu(torch.randn(nI, dI, generator=g, device=dev))
    WI = 0.05 * torch.randn(CI, dI, generator=g, device=dev)
    bI = torch.randn(CI, generator=g, device=dev)
    plI = _ops.linear_planes(WI)
    ms = _time_op(torch, lambda: _ops.clip_linear_lse(XI, WI, bI, clip=1.0, planes=plI), reps=3)
    out["react_768_c1000"] = {"embeddings_per_s": nI / (ms * 1e-3), "ms": ms,
                              "roofline": _tensor_roofline(2.0 * nI * dI * CI, ms, tf32_probe, bf16_peak)}
    ms = _time_op(torch, lambda: _ops.ash_linear_lse(XI, WI, bI, 115, planes=plI), reps=3)
    out["ash_768_c1000"] = {"embeddings_per_s": nI / (ms * 1e-3), "ms": ms, "note": "prune kernel + wide head"}
    del XI, WI, bI, plI
    # (a7) ViM: d = 512, residual space 256, C = 10 logits
    NS = np.linalg.qr(rng.randn(d, d))[0][:, :256]
    vst = _ops.vim_prepare(rng.randn(d) * 0.1, NS, 1.7)
    Lg = torch.randn(n, C, generator=g, device=dev)
    ms = _time_op(torch, lambda: _ops.vim_score(X, Lg, vst))
    fl = n * (2.0 * d * 256 + 3 * C)
    out["vim_512"] = {"embeddings_per_s": n / (ms * 1e-3), "ms": ms, "fp32_equiv_tflops": fl / (ms * 1e-3) / 1e12,
                      "tensor_roofline": _tensor_roofline(fl, ms, tf32_probe, bf16_peak), "roofline": hbm(n * (d + C) * 4, ms)}
    del Lg
    # (a6) class-conditional Mahalanobis: d = 512, C = 10 (FP32 SIMT contraction)
    A = rng.randn(d, d)
    prec = A @ A.T / d + np.eye(d)
    cst = _ops.classcond_prepare(rng.randn(C, d), prec)
    n6 = 500_000
    ms = _time_op(torch, lambda: _ops.classcond_score(X[:n6], cst))
    fl = n6 * (2.0 * d * cst.r + 3.0 * cst.r * C)
    out["mahalanobis_512_c10"] = {"embeddings_per_s": n6 / (ms * 1e-3), "ms": ms, "fp32_tflops": fl / (ms * 1e-3) / 1e12,
                                  "roofline": _tensor_roofline(fl, ms, tf32_probe, bf16_peak)}
    # (a9) DDU / GMM: C = 10 whitening contractions of 512 x 512
    mus = rng.randn(C, d)
    Ls = np.stack([np.linalg.cholesky(prec) for _ in range(C)])
    gst = _ops.gmm_prepare(mus, Ls)
    n9 = 200_000
    ms = _time_op(torch, lambda: _ops.gmm_lse(X[:n9], gst))
    fl = n9 * C * 2.0 * d * d
    out["ddu_512_c10"] = {"embeddings_per_s": n9 / (ms * 1e-3), "ms": ms, "fp32_tflops": fl / (ms * 1e-3) / 1e12,
                          "roofline": _tensor_roofline(fl, ms, tf32_probe, bf16_peak)}
    del X
    # (a4) LaRED KDE: 50k x 256 bank, 100k queries
    bank = 0.5 + torch.randn(50_000, 256, generator=g, device=dev)
    q = torch.randn(100_000, 256, generator=g, device=dev)
    kb = _ops.kde_bank(bank)
    ms = _time_op(torch, lambda: _ops.kde_score(q, kb), reps=3)
    fl = 2.0 * 100_000 * 50_000 * 256
    out["kde_50k_256"] = {"queries_per_s": 100_000 / (ms * 1e-3), "ms": ms, "fp32_equiv_tflops": fl / (ms * 1e-3) / 1e12,
                          "roofline": _tensor_roofline(fl, ms, tf32_probe, bf16_peak)}
    # (f1) OoD metrics: AUROC + FPR@95 + AUPR of 1e7 InD vs 1e7 OoD float32 scores (radix sort + fused scan)
    nm = 10_000_000
    si = torch.sigmoid(0.5 + torch.randn(nm, generator=g, device=dev))
    so = torch.sigmoid(-0.5 + torch.randn(nm, generator=g, device=dev))
    ms = _time_op(torch, lambda: _ops.ood_metrics(si, so, want_curve=False), reps=3)
    out["metrics_2e7_f32"] = {"scores_per_s": 2 * nm / (ms * 1e-3), "ms": ms,
                              # bytes the passes move per score: keys out 4+12, four radix passes x (8 + 12 + 12), two scan reads x 12
                              "roofline": hbm(2 * nm * (4 + 12 + 4 * 32 + 2 * 12), ms)}
    del si, so
    # (f3) MC-DropBlock sampler + H x W mean: 1024 ResNet-18 layer-4 maps (512 x 7 x 7), 16 samples each, against
    # the same arithmetic written with torch ops the way DropBlock2D + the reference reducer run it (16 passes)
    import torch.nn.functional as F
    Bm, Cm, Hm, n_mc, bs = 1024, 512, 7, 16, 3
    xm = torch.randn(Bm, Cm, Hm, Hm, generator=g, device=dev)
    seed = (torch.rand(n_mc, Bm, Hm, Hm, generator=g, device=dev) < 0.3 / bs**2).to(torch.uint8)
    ms = _time_op(torch, lambda: _ops.mc_dropblock_mean(xm, seed, bs), reps=5)

    def torch_chain():
        rows = []
        for m in range(n_mc):
            bm = 1 - F.max_pool2d(seed[m].float()[:, None], kernel_size=bs, stride=1, padding=bs // 2).squeeze(1)
            o = xm * bm[:, None] * (Hm * Hm) / bm.sum((1, 2))[:, None, None, None]
            rows.append(o.mean(3).mean(2))
        return torch.stack(rows, 1)

    ms_t = _time_op(torch, torch_chain, reps=3)
    out["mc_sampler_1024x512x7x7"] = {"images_per_s": Bm / (ms * 1e-3), "ms": ms, "torch_ops_same_gpu_ms": ms_t,
                                      "roofline": hbm(xm.numel() * 4 + Bm * n_mc * Cm * 4, ms)}
    del xm, seed
    # (a2 + a3 fused) PCA 512 -> 256 + LaREM as one contraction over the raw latents (2,056 B / embedding)
    rngf = np.random.RandomState(5)
    comps = np.linalg.qr(rngf.randn(512, 256))[0].T.copy()
    var = 1.0 + rngf.rand(256)
    Af = rngf.randn(256, 256)
    fst = _ops.md_fold_pca(rngf.randn(512), comps, var, True, 0.01 * rngf.randn(256), Af @ Af.T / 256 + np.eye(256))
    pst = _ops.pca_prepare(rngf.randn(512), comps, var, True)
    mst = _ops.md_prepare(0.01 * rngf.randn(256), Af @ Af.T / 256 + np.eye(256))
    xr = torch.randn(2_000_000, 512, generator=g, device=dev)
    ms = _time_op(torch, lambda: _ops.md_score(xr, fst))
    ms_st = _time_op(torch, lambda: _ops.md_score(_ops.pca_transform(xr, pst), mst))
    out["pca_larem_fused_512_256"] = {"embeddings_per_s": 2_000_000 / (ms * 1e-3), "ms": ms, "staged_two_kernels_ms": ms_st,
                                      "fp32_equiv_tflops": 2_000_000 * 262_656 / (ms * 1e-3) / 1e12,
                                      "tensor_roofline": _tensor_roofline(2_000_000 * 262_656.0, ms, tf32_probe, bf16_peak),
                                      "roofline": hbm(2_000_000 * 2056, ms)}
    del xr
    # (f3) online LaREx chain on one hooked map (sampler -> entropy -> folded PCA + LaREM -> score on the host): the body
    # of LaRExInference.get_score after the model forward (`score_latent`), wall clock per image: ONE CUDA graph replay
    # per image (default) and the same chain launch by launch; and the chain batched over 1024 maps
    import time as _time
    from sklearn.decomposition import PCA as _PCA

    from runia_core_b200.feature_extraction import MCSamplerModule
    from runia_core_b200.inference import LaRExInference, MDLatentSpace
    pca_o = _PCA(n_components=256, whiten=True)
    pca_o.mean_, pca_o.components_, pca_o.explained_variance_ = rngf.randn(512), comps, var
    md_o = MDLatentSpace()
    md_o.feats_mean, md_o.precision, md_o._setup_flag = 0.01 * rngf.randn(1, 256), Af @ Af.T / 256 + np.eye(256), True
    inf = LaRExInference(torch.nn.Identity(), md_o, drop_block_prob=0.3, drop_block_size=3, mcd_samples_nro=16,
                         mcd_sampler=MCSamplerModule, pca_transform=pca_o)
    lat1 = torch.randn(1, 512, 7, 7, generator=g, device=dev)
    latB = torch.randn(1024, 512, 7, 7, generator=g, device=dev)

    def per_image(reps=300):
        for _ in range(20):
            inf.score_latent(lat1)
        torch.cuda.synchronize()
        t0 = _time.perf_counter()
        for _ in range(reps):
            inf.score_latent(lat1)
        return (_time.perf_counter() - t0) / reps * 1e6

    us_graph = per_image()
    inf.use_cuda_graph = False
    us_plain = per_image()
    inf.score_latent(latB)
    torch.cuda.synchronize()
    t0 = _time.perf_counter()
    for _ in range(20):
        inf.score_latent(latB)
    msB = (_time.perf_counter() - t0) / 20 * 1e3
    out["larex_online_chain"] = {"us_per_image_batch1": us_graph, "us_per_image_batch1_launch_by_launch": us_plain,
                                 "ms_per_1024_images": msB, "images_per_s_batched": 1024 / (msB * 1e-3),
                                 "note": "wall clock from the hooked 512 x 7 x 7 map to the float64 score on the host: CPU-generator "
                                         "seeds (one torch.rand), 16 samples, sampler + entropy + folded PCA/LaREM as one CUDA graph"}
    del lat1, latB
    # (f2) setup() statistics: class means + float64 Gram matrix of a 50k x 512 bank with 10 classes
    xs = torch.randn(50_000, 512, generator=g, device=dev)
    lab = torch.randint(0, 10, (50_000,), generator=g, device=dev).cpu().numpy()

    def fit():
        means, counts, xf, lb = _ops.class_means(xs, lab, 10)
        return _ops.centered_covariance(xf, lb, means, int(counts.sum()))

    ms = _time_op(torch, fit, reps=3)
    out["fit_stats_50k_512_c10"] = {"ms": ms, "fp64_tflops": 50_000 * 512 * 513 / (ms * 1e-3) / 1e12,
                                    "note": "means + Gram + reduction + copy-back of the 512 x 512 result"}
    del xs
    return out


def _extra_sweep_config2(R):
    """BASELINE configs[1]: the full baseline sweep on ResNet-18 / CIFAR-10 shapes (50k x 512 train
    bank, 10 classes, 10k test rows), through the reference-facing classes with NumPy in / NumPy
    out (H2D + kernels + D2H inside the timed call).  setup() (host-side fits, as in the
    reference) is timed separately."""
    rng = np.random.RandomState(11)
    C, d, ntr, nte = 10, 512, 50_000, 10_000
    means = rng.randn(C, d).astype(np.float32)
    ytr = rng.randint(0, C, ntr)
    train = (means[ytr] + rng.randn(ntr, d)).astype(np.float32)
    yv = rng.randint(0, C, nte)
    valid = (means[yv] + rng.randn(nte, d)).astype(np.float32)
    test = np.concatenate([valid[: nte // 2], (1.5 * rng.randn(nte - nte // 2, d)).astype(np.float32)])
    W = (0.05 * rng.randn(C, d)).astype(np.float32)
    b = rng.randn(C).astype(np.float32)
    lg = lambda x: (x @ W.T + b).astype(np.float32)  # noqa: E731
    tr_l, va_l, te_l = lg(train), lg(valid), lg(test)
    fc = {"weight": W, "bias": b}
    I = R.inference
    mk = {
        "msp": lambda: I.MSP(flip_sign=False),
        "energy": lambda: I.Energy(flip_sign=False),
        "mahalanobis": lambda: I.Mahalanobis(flip_sign=False, num_classes=C),
        "knn": lambda: I.KNN(flip_sign=False, k_neighbors=50),
        "vim": lambda: I.ViM(flip_sign=False),
        "ddu": lambda: I.DDU(flip_sign=False, num_classes=C),
        "react": lambda: I.ReAct(flip_sign=False, react_percentile=90),
        "dice": lambda: I.DICE(flip_sign=False, dice_percentile=90, num_classes=C),
    }
    out = {"timing": "median of 20 synchronised calls (ms), mean and max beside it: the host is a shared 16-vCPU VM"}
    import torch

    for name, ctor in mk.items():
        p = ctor()
        t0 = time.perf_counter()
        if name in ("msp", "energy"):
            p.setup(tr_l)
            call = lambda: p.postprocess(te_l)  # noqa: E731
        else:
            p.setup(train, valid_feats=valid, train_labels=ytr, train_logits=tr_l, valid_logits=va_l,
                    final_linear_layer_params=fc)
            call = lambda: p.postprocess(test, logits=te_l)  # noqa: E731
        torch.cuda.synchronize()
        t_setup = time.perf_counter() - t0
        call()
        torch.cuda.synchronize()
        ts = []
        for _ in range(20):  # one synchronised call at a time: NumPy in -> NumPy out latency
            t0 = time.perf_counter()
            sc = call()
            ts.append(time.perf_counter() - t0)
        dt = float(np.median(ts))
        out[name] = {"embeddings_per_s": nte / dt, "ms": dt * 1e3, "ms_mean": float(np.mean(ts)) * 1e3,
                     "ms_max": float(np.max(ts)) * 1e3, "setup_s": round(t_setup, 3),
                     "finite": bool(np.isfinite(sc).all())}
    return out


def _cpu_larem(md, seconds=8.0):
    """The reference's CPU path for the same scorer (postprocessors.py:241-242: N x N product),
    restated in oracle/oracle_np.py, on a bounded sample: 10k-row calls (800 MB temporary each,
    the largest the formulation affords) until `seconds` of work."""
    from oracle import oracle_np as O

    threads = _all_host_threads()
    rng = np.random.RandomState(5)
    x = rng.randn(10_000, D_LATENT).astype(np.float32)
    O.md_score_faithful(x[:2000], md.feats_mean, md.precision)
    t0 = time.perf_counter()
    n = 0
    while time.perf_counter() - t0 < seconds and n < 30:
        O.md_score_faithful(x, md.feats_mean, md.precision)
        n += 1
    dt = time.perf_counter() - t0
    return {"value": n * 10_000 / dt, "unit": "embeddings/s", "cores": int(threads), "kind": "port",
            "sample": f"{n} calls x 10,000 rows x d=256 of MDLatentSpace.postprocess (N x N form), {dt:.1f} s"}


def _cpu_entropy_item(args):
    """One item of get_dl_h_z(parallel_run=True) (evaluation/entropy.py:85-91: process_map over the items)."""
    from oracle import oracle_np as O

    z, n_mc = args
    return O.get_dl_h_z_faithful(z, n_mc)[1]


def _cpu_other_rows():
    """The reference's CPU path for the other measured rows (oracle ports with the reference's loop
    structure), on bounded samples of the same shapes, all host threads: entropy (one estimator call
    per item and per (item, dimension), evaluation/entropy.py:56-84), kNN (one brute-force search per
    query, postprocessors.py:417-421), metrics (torchmetrics / sklearn restatement)."""
    from oracle import oracle_np as O

    threads = _all_host_threads()
    rng = np.random.RandomState(9)
    out = {"cores": int(threads), "kind": "port"}
    n_items, n_mc, D = 12, 16, 512
    z = (rng.randn(n_items, 1, D) + 0.1 * rng.randn(n_items, n_mc, D)).astype(np.float32).reshape(-1, D)
    t0 = time.perf_counter()
    O.get_dl_h_z_faithful(z, n_mc)
    dt = time.perf_counter() - t0
    out["entropy_config1"] = {"items_per_s": n_items / dt, "sample": f"{n_items} items x 16 x 512, {dt:.1f} s (single-threaded Python loops upstream)"}
    # the reference's parallel_run=True: a process pool over the items (chunksize 1), all host cores
    import multiprocessing as mp

    cores = os.cpu_count() or 1
    n_par = 4 * cores
    zp = (rng.randn(n_par, 1, D) + 0.1 * rng.randn(n_par, n_mc, D)).astype(np.float32)
    try:
        with mp.get_context("fork").Pool(cores) as pool:
            pool.map(_cpu_entropy_item, [(zp[i], n_mc) for i in range(cores)], chunksize=1)  # warm the workers
            t0 = time.perf_counter()
            pool.map(_cpu_entropy_item, [(zp[i], n_mc) for i in range(n_par)], chunksize=1)
            dt = time.perf_counter() - t0
        out["entropy_config1_parallel_run"] = {"items_per_s": n_par / dt, "processes": cores,
                                               "sample": f"{n_par} items x 16 x 512 over a {cores}-process pool, {dt:.1f} s"}
    except Exception as e:  # pragma: no cover
        out["entropy_config1_parallel_run"] = {"error": repr(e)}
    bank = O.normalize_rows_exact(rng.randn(50_000, 512).astype(np.float32))
    q = rng.randn(40, 512).astype(np.float32)
    t0 = time.perf_counter()
    O.knn_score_faithful(q, bank, 50)
    dt = time.perf_counter() - t0
    out["knn_config2"] = {"queries_per_s": len(q) / dt, "sample": f"{len(q)} queries x 50k x 512 bank, {dt:.1f} s"}
    n = 1_000_000
    ind, ood = 0.5 + rng.randn(n).astype(np.float32), -0.5 + rng.randn(n).astype(np.float32)
    t0 = time.perf_counter()
    O.ood_metrics(ind, ood)
    dt = time.perf_counter() - t0
    out["metrics_f32"] = {"scores_per_s": 2 * n / dt, "sample": f"2 x {n} scores, {dt:.1f} s"}
    return out


def run_reference(args):
    """Reference arm: the reference's own CPU implementation (faithful oracle port), all BLAS threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle_np as O

    threads = _all_host_threads()
    train = _fit_larem()
    mean, prec = O.md_fit(train)
    rng = np.random.RandomState(5)
    x = rng.randn(10_000, D_LATENT).astype(np.float32)
    for _ in range(max(1, min(args.warmup, 3))):
        O.md_score_faithful(x, mean, prec)
    steps = max(1, min(args.steps, 20))
    t0 = time.perf_counter()
    for _ in range(steps):
        O.md_score_faithful(x, mean, prec)
    dt = (time.perf_counter() - t0) / steps
    v = 10_000 / dt
    line = {"impl": "reference", "metric": "ood_scored_embeddings_per_sec", "value": v, "unit": "embeddings/s",
            "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": steps, "warmup": args.warmup,
            "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": _config(args.rows, int(os.environ.get("WORLD_SIZE", "1"))),
            "cpu_baseline": {"value": v, "unit": "embeddings/s", "cores": int(threads), "kind": "port",
                             "sample": "each step = one 10,000-row call of MDLatentSpace.postprocess in the reference's N x N "
                                       "form (postprocessors.py:241-242; 800 MB temporary, the largest call the formulation "
                                       "affords): a bounded sample of the workload's rows"},
            "e2e": {"value": v, "unit": "embeddings/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    _emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rows", type=int, default=4 * 1024 * 1024, help="embeddings per GPU per step")
    ap.add_argument("--e2e-rows", type=int, default=1024 * 1024)
    ap.add_argument("--knn-shard-rows", type=int, default=1_250_000)
    ap.add_argument("--cpu-seconds", type=float, default=8.0)
    ap.add_argument("--no-extra", action="store_true")
    ap.add_argument("--no-sweep", action="store_true", help="skip the configs[1] baseline sweep (host-side fits take ~20 s)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip BASELINE configs[2..4] at their stated sizes")
    ap.add_argument("--c4-queries", type=int, default=50_000)
    ap.add_argument("--c3-boxes", type=int, default=1_000_000)
    args = ap.parse_args()
    _claim_stdout()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
