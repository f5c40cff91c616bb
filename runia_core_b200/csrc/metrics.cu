// (f1) OoD detection metrics on the device: AUROC, FPR@95 and AUPR of InD-vs-OoD score arrays with the
// semantics of evaluation/metrics.py:37-100 (`get_auroc_results`), i.e. torchmetrics 1.8.2
// `auroc` / `roc` / `precision_recall_curve` (task="binary") + sklearn.metrics.auc:
//   * scores outside [0, 1] anywhere -> every score goes through a sigmoid (in the score dtype);
//   * sort descending, one curve point per DISTINCT score; tps = cumsum(label), fps = rank - tps;
//   * ROC gets (0, 0) prepended, AUROC = trapezoid(tpr, fpr), FPR@95 = fpr at the first tpr >= 0.95
//     (float32 ratios, like the reference's float32 tensors);
//   * PR points reversed with (recall 0, precision 1) appended, AUPR = |trapezoid(precision, recall)|.
// At 1e7-1e8 scores the reference spends its time in a CPU sort; here: order-preserving 64-bit keys,
// an LSD radix sort (8 bits per pass, warp-match ranking, stable), ONE fused scan (label sum, last
// curve point, label sum at the last curve point, curve-point count) and a final pass that turns every
// curve point into its trapezoid terms.  The AUROC numerator is accumulated exactly in integers.
#include <algorithm>

#include "common.cuh"

namespace runia {

// ------------------------------------------------------------------------------------------------
// keys
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) range_flag_kernel(const T *__restrict__ a, int64_t na, const T *__restrict__ b,
                                                         int64_t nb, uint32_t *__restrict__ flag) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  bool out = false;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < na + nb; i += stride) {
    const T v = i < na ? a[i] : b[i - na];
    out |= !(v >= (T)0 && v <= (T)1);
  }
  if (__any_sync(0xffffffffu, out) && (threadIdx.x & 31) == 0) atomicOr(flag, 1u);
}

__device__ __forceinline__ uint64_t ordered_key(double v) {
  const uint64_t u = (uint64_t)__double_as_longlong(v);
  return (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ uint64_t ordered_key(float v) {
  const uint32_t u = __float_as_uint(v);
  return (uint64_t)((u & 0x80000000u) ? ~u : (u | 0x80000000u)) << 32;  // only the high word is sorted
}

// key = ~ordered(score) so that an ascending sort walks the scores in descending order; value = label
template <typename T>
__global__ void __launch_bounds__(256) make_keys_kernel(const T *__restrict__ a, int64_t na, const T *__restrict__ b,
                                                        int64_t nb, const uint32_t *__restrict__ flag,
                                                        uint64_t *__restrict__ keys, uint32_t *__restrict__ labels) {
  const bool sig = *flag != 0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < na + nb; i += stride) {
    T v = i < na ? a[i] : b[i - na];
    if (sig) v = (T)1 / ((T)1 + (sizeof(T) == 8 ? (T)exp(-(double)v) : (T)expf(-(float)v)));
    keys[i] = ~ordered_key(v);
    labels[i] = i < na ? 1u : 0u;
  }
}

// ------------------------------------------------------------------------------------------------
// exclusive scan of uint32 (three launches): used for the radix-sort digit offsets
// ------------------------------------------------------------------------------------------------
constexpr int SC_THREADS = 256, SC_ITEMS = 16, SC_TILE = SC_THREADS * SC_ITEMS;

__device__ __forceinline__ uint32_t block_exclusive_scan_u32(uint32_t v, uint32_t *warp_sums, uint32_t &total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t inc = v;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, inc, off);
    if (lane >= off) inc += t;
  }
  if (lane == 31) warp_sums[warp] = inc;
  __syncthreads();
  uint32_t base = 0, tot = 0;
  for (int w = 0; w < SC_THREADS / 32; ++w) {
    const uint32_t s = warp_sums[w];
    if (w < warp) base += s;
    tot += s;
  }
  __syncthreads();
  total = tot;
  return base + inc - v;
}

__global__ void __launch_bounds__(SC_THREADS) scan_u32_partial_kernel(const uint32_t *__restrict__ in, int64_t n,
                                                                      uint32_t *__restrict__ tile_sums) {
  __shared__ uint32_t ws[SC_THREADS / 32];
  const int64_t base = (int64_t)blockIdx.x * SC_TILE + (int64_t)threadIdx.x * SC_ITEMS;
  uint32_t s = 0;
#pragma unroll
  for (int j = 0; j < SC_ITEMS; ++j) s += (base + j < n) ? in[base + j] : 0u;
  uint32_t total;
  block_exclusive_scan_u32(s, ws, total);
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}
// single block: exclusive scan of the tile sums in place
__global__ void __launch_bounds__(SC_THREADS) scan_u32_tiles_kernel(uint32_t *__restrict__ tile_sums, int64_t nt) {
  __shared__ uint32_t ws[SC_THREADS / 32];
  uint32_t carry = 0;
  for (int64_t t0 = 0; t0 < nt; t0 += SC_THREADS) {
    const int64_t t = t0 + threadIdx.x;
    const uint32_t v = t < nt ? tile_sums[t] : 0u;
    uint32_t total;
    const uint32_t ex = block_exclusive_scan_u32(v, ws, total);
    if (t < nt) tile_sums[t] = carry + ex;
    carry += total;
  }
}
__global__ void __launch_bounds__(SC_THREADS) scan_u32_final_kernel(uint32_t *__restrict__ data, int64_t n,
                                                                    const uint32_t *__restrict__ tile_sums) {
  __shared__ uint32_t ws[SC_THREADS / 32];
  const int64_t base = (int64_t)blockIdx.x * SC_TILE + (int64_t)threadIdx.x * SC_ITEMS;
  uint32_t v[SC_ITEMS], s = 0;
#pragma unroll
  for (int j = 0; j < SC_ITEMS; ++j) {
    v[j] = (base + j < n) ? data[base + j] : 0u;
    s += v[j];
  }
  uint32_t total;
  uint32_t run = tile_sums[blockIdx.x] + block_exclusive_scan_u32(s, ws, total);
#pragma unroll
  for (int j = 0; j < SC_ITEMS; ++j) {
    if (base + j < n) data[base + j] = run;
    run += v[j];
  }
}

// ------------------------------------------------------------------------------------------------
// LSD radix sort of (uint64 key, uint32 value), 8 bits per pass, stable
// ------------------------------------------------------------------------------------------------
constexpr int RS_THREADS = 256, RS_WARPS = 8, RS_ROUNDS = 16, RS_TILE = RS_THREADS * RS_ROUNDS;  // 4096

__global__ void __launch_bounds__(RS_THREADS) rs_hist_kernel(const uint64_t *__restrict__ keys, int64_t n, int shift,
                                                             uint32_t *__restrict__ hist, int64_t nblk) {
  __shared__ uint32_t h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * RS_TILE;
#pragma unroll 4
  for (int r = 0; r < RS_ROUNDS; ++r) {
    const int64_t i = base + r * RS_THREADS + threadIdx.x;
    if (i < n) atomicAdd(&h[(uint32_t)(keys[i] >> shift) & 255u], 1u);
  }
  __syncthreads();
  hist[(int64_t)threadIdx.x * nblk + blockIdx.x] = h[threadIdx.x];  // bin-major: one scan gives global offsets
}

// lanes of `vm` that hold the same 8-bit digit as this lane: eight ballots (VOTE is far cheaper than
// MATCH.ANY, which made the scatter ADU-bound: ncu sm__throughput 75 % at 16 % issue)
__device__ __forceinline__ unsigned same_digit_mask(unsigned vm, uint32_t d) {
  unsigned peers = vm;
#pragma unroll
  for (int b = 0; b < 8; ++b) {
    const bool bit = (d >> b) & 1u;
    const unsigned bal = __ballot_sync(vm, bit);
    peers &= bit ? bal : ~bal;
  }
  return peers;
}

// item order inside a tile: warp w owns items [512 w, 512 w + 512), round r of the warp covers 32
// consecutive items -> ranks by (warp, round, lane) are the original order: the sort is stable.
// The tile is first sorted by digit in shared memory, so that the global writes of a digit are one
// contiguous run per tile (coalesced) instead of 4096 scattered 12-byte stores.
constexpr size_t kScatterSmem = (size_t)RS_TILE * 12 + (size_t)(RS_WARPS + 2) * 256 * 4 + 64;

__global__ void __launch_bounds__(RS_THREADS)
rs_scatter_kernel(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ vals, int64_t n, int shift,
                  const uint32_t *__restrict__ offsets, int64_t nblk, uint64_t *__restrict__ keys_out,
                  uint32_t *__restrict__ vals_out) {
  extern __shared__ __align__(16) unsigned char rs_smem[];
  uint64_t *skey = reinterpret_cast<uint64_t *>(rs_smem);                       // [RS_TILE]
  uint32_t *sval = reinterpret_cast<uint32_t *>(rs_smem + (size_t)RS_TILE * 8); // [RS_TILE]
  uint32_t(*cnt)[256] = reinterpret_cast<uint32_t(*)[256]>(rs_smem + (size_t)RS_TILE * 12);  // [RS_WARPS][256]
  uint32_t *tile_base = &cnt[0][0] + RS_WARPS * 256;  // [256] first slot of each digit in the sorted tile
  uint32_t *gofs = tile_base + 256;                   // [256] global offset of that slot
  uint32_t *ws = gofs + 256;                          // scan scratch
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int e = threadIdx.x; e < RS_WARPS * 256; e += RS_THREADS) (&cnt[0][0])[e] = 0;
  __syncthreads();
  const int64_t tbase = (int64_t)blockIdx.x * RS_TILE;
  const int64_t wbase = tbase + (int64_t)warp * (32 * RS_ROUNDS);
  const int tile_n = (int)((n - tbase < RS_TILE) ? (n - tbase) : RS_TILE);
  uint64_t k[RS_ROUNDS];
  uint32_t v[RS_ROUNDS];
#pragma unroll
  for (int r = 0; r < RS_ROUNDS; ++r) {
    const int64_t i = wbase + r * 32 + lane;
    k[r] = i < n ? keys[i] : 0ull;
    v[r] = i < n ? vals[i] : 0u;
  }
  // phase 1: per-warp digit counts
#pragma unroll
  for (int r = 0; r < RS_ROUNDS; ++r) {
    const bool valid = wbase + r * 32 + lane < n;
    const unsigned vm = __ballot_sync(0xffffffffu, valid);
    if (valid) {
      const uint32_t d = (uint32_t)(k[r] >> shift) & 255u;
      const unsigned m = same_digit_mask(vm, d);
      if ((m & ((1u << lane) - 1u)) == 0) cnt[warp][d] += __popc(m);  // one lane per distinct digit
    }
    __syncwarp();
  }
  __syncthreads();
  // phase 2: thread = digit: offsets of each warp inside the digit's run, the run's first slot in the
  // sorted tile (exclusive scan over the digits) and its global offset
  {
    uint32_t run = 0;
#pragma unroll
    for (int w = 0; w < RS_WARPS; ++w) {
      const uint32_t c = cnt[w][threadIdx.x];
      cnt[w][threadIdx.x] = run;
      run += c;
    }
    uint32_t total;
    tile_base[threadIdx.x] = block_exclusive_scan_u32(run, ws, total);
    gofs[threadIdx.x] = offsets[(int64_t)threadIdx.x * nblk + blockIdx.x];
  }
  __syncthreads();
  // phase 3: rank every item and place it in the sorted tile
#pragma unroll
  for (int r = 0; r < RS_ROUNDS; ++r) {
    const bool valid = wbase + r * 32 + lane < n;
    const unsigned vm = __ballot_sync(0xffffffffu, valid);
    if (valid) {
      const uint32_t d = (uint32_t)(k[r] >> shift) & 255u;
      const unsigned m = same_digit_mask(vm, d);
      const uint32_t pos = tile_base[d] + cnt[warp][d] + __popc(m & ((1u << lane) - 1u));
      skey[pos] = k[r];
      sval[pos] = v[r];
      __syncwarp(vm);
      if ((m & ((1u << lane) - 1u)) == 0) cnt[warp][d] += __popc(m);
    }
    __syncwarp();
  }
  __syncthreads();
  // phase 4: coalesced copy-out: slot i of the sorted tile belongs to digit d and goes to gofs[d] + (i - tile_base[d])
#pragma unroll 4
  for (int r = 0; r < RS_ROUNDS; ++r) {
    const int i = r * RS_THREADS + threadIdx.x;
    if (i < tile_n) {
      const uint64_t kk = skey[i];
      const uint32_t d = (uint32_t)(kk >> shift) & 255u;
      const uint32_t pos = gofs[d] + ((uint32_t)i - tile_base[d]);
      keys_out[pos] = kk;
      vals_out[pos] = sval[i];
    }
  }
}

// ------------------------------------------------------------------------------------------------
// the fused curve scan.  State of a segment of the sorted sequence:
//   S   label sum, nB number of curve points (positions where the score changes or the sequence ends),
//   last index of its last curve point (-1: none), SB label sum from the segment start through that point
// ------------------------------------------------------------------------------------------------
struct CurveState {
  uint32_t S, nB, SB;
  int32_t last;
};
__device__ __forceinline__ CurveState combine(const CurveState &L, const CurveState &R) {
  CurveState o;
  o.S = L.S + R.S;
  o.nB = L.nB + R.nB;
  o.last = R.last >= 0 ? R.last : L.last;
  o.SB = R.last >= 0 ? L.S + R.SB : L.SB;
  return o;
}
__device__ __forceinline__ CurveState curve_identity() { return CurveState{0u, 0u, 0u, -1}; }
__device__ __forceinline__ CurveState shfl_up_state(const CurveState &s, int off) {
  CurveState o;
  o.S = __shfl_up_sync(0xffffffffu, s.S, off);
  o.nB = __shfl_up_sync(0xffffffffu, s.nB, off);
  o.SB = __shfl_up_sync(0xffffffffu, s.SB, off);
  o.last = __shfl_up_sync(0xffffffffu, s.last, off);
  return o;
}

constexpr int CV_THREADS = 256, CV_ITEMS = 8, CV_TILE = CV_THREADS * CV_ITEMS;

__device__ __forceinline__ CurveState element_state(const uint64_t *keys, const uint32_t *labels, int64_t i, int64_t n) {
  if (i >= n) return curve_identity();
  const uint32_t y = labels[i];
  const bool b = (i == n - 1) || (keys[i] != keys[i + 1]);
  return CurveState{y, b ? 1u : 0u, b ? y : 0u, b ? (int32_t)i : -1};
}

// inclusive block scan of per-thread aggregates; returns the exclusive prefix of this thread and the block total
__device__ __forceinline__ CurveState block_exclusive_scan_state(const CurveState &mine, CurveState *warp_tot,
                                                                 CurveState &total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  CurveState inc = mine;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const CurveState t = shfl_up_state(inc, off);
    if (lane >= off) inc = combine(t, inc);
  }
  if (lane == 31) warp_tot[warp] = inc;
  __syncthreads();
  CurveState base = curve_identity(), tot = curve_identity();
  for (int w = 0; w < CV_THREADS / 32; ++w) {
    if (w < warp) base = combine(base, warp_tot[w]);
    tot = combine(tot, warp_tot[w]);
  }
  __syncthreads();
  total = tot;
  CurveState ex = shfl_up_state(inc, 1);
  if (lane == 0) ex = curve_identity();
  return combine(base, ex);
}

__global__ void __launch_bounds__(CV_THREADS) curve_partial_kernel(const uint64_t *__restrict__ keys,
                                                                   const uint32_t *__restrict__ labels, int64_t n,
                                                                   CurveState *__restrict__ tile_state) {
  __shared__ CurveState ws[CV_THREADS / 32];
  const int64_t base = (int64_t)blockIdx.x * CV_TILE + (int64_t)threadIdx.x * CV_ITEMS;
  CurveState s = curve_identity();
#pragma unroll
  for (int j = 0; j < CV_ITEMS; ++j) s = combine(s, element_state(keys, labels, base + j, n));
  CurveState total;
  block_exclusive_scan_state(s, ws, total);
  if (threadIdx.x == 0) tile_state[blockIdx.x] = total;
}
__global__ void __launch_bounds__(CV_THREADS) curve_tiles_kernel(CurveState *__restrict__ tile_state, int64_t nt) {
  __shared__ CurveState ws[CV_THREADS / 32];
  CurveState carry = curve_identity();
  for (int64_t t0 = 0; t0 < nt; t0 += CV_THREADS) {
    const int64_t t = t0 + threadIdx.x;
    const CurveState v = t < nt ? tile_state[t] : curve_identity();
    CurveState total;
    const CurveState ex = block_exclusive_scan_state(v, ws, total);
    if (t < nt) tile_state[t] = combine(carry, ex);  // exclusive prefix of tile t
    carry = combine(carry, total);
  }
}

struct MetricsOut {
  unsigned long long auroc_num;  // sum over curve points of (fps - fps_prev) (tps + tps_prev): exact
  unsigned long long first95;    // min over curve points with float32 tpr >= 0.95 of (index << 32 | fps)
  uint32_t n_points;             // curve points (distinct scores)
};

// every curve point turns into its trapezoid terms; AUPR partial sums go to one slot per block (summed in a
// fixed order by the host-side finish kernel: deterministic)
__global__ void __launch_bounds__(CV_THREADS)
curve_final_kernel(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ labels, int64_t n,
                   const CurveState *__restrict__ tile_state, uint32_t P, uint32_t Nn, MetricsOut *__restrict__ mo,
                   double *__restrict__ aupr_partial, float *__restrict__ fpr_out, float *__restrict__ tpr_out) {
  __shared__ CurveState ws[CV_THREADS / 32];
  __shared__ double red[CV_THREADS / 32];
  __shared__ unsigned long long redn[CV_THREADS / 32];
  const int64_t base = (int64_t)blockIdx.x * CV_TILE + (int64_t)threadIdx.x * CV_ITEMS;
  CurveState e[CV_ITEMS], s = curve_identity();
#pragma unroll
  for (int j = 0; j < CV_ITEMS; ++j) {
    e[j] = element_state(keys, labels, base + j, n);
    s = combine(s, e[j]);
  }
  CurveState total;
  CurveState run = combine(tile_state[blockIdx.x], block_exclusive_scan_state(s, ws, total));
  double aupr = 0.0;
  unsigned long long num = 0;
  unsigned long long first95 = ~0ull;
  const float fP = (float)P, fN = (float)Nn;
#pragma unroll
  for (int j = 0; j < CV_ITEMS; ++j) {
    if (e[j].last >= 0) {  // curve point at index i = base + j; `run` = state of everything before it
      const int64_t i = base + j;
      const uint32_t tps = run.S + e[j].S, fps = (uint32_t)(i + 1) - tps;
      const bool has_prev = run.last >= 0;
      const uint32_t tps_p = has_prev ? run.SB : 0u;
      const uint32_t fps_p = has_prev ? (uint32_t)(run.last + 1) - tps_p : 0u;
      num += (unsigned long long)(fps - fps_p) * (unsigned long long)(tps + tps_p);
      const double prec = (double)tps / (double)(tps + fps);
      const double prec_p = has_prev ? (double)tps_p / (double)(tps_p + fps_p) : 1.0;  // appended (recall 0, precision 1)
      aupr += ((double)(tps - tps_p) / (double)P) * (prec + prec_p) * 0.5;
      const float tpr32 = __fdiv_rn((float)tps, fP);
      if (tpr32 >= 0.95f) first95 = min(first95, ((unsigned long long)i << 32) | fps);
      const uint32_t slot = run.nB + 1;  // slot 0 is the prepended (0, 0)
      if (fpr_out) fpr_out[slot] = Nn ? __fdiv_rn((float)fps, fN) : 0.f;
      if (tpr_out) tpr_out[slot] = P ? tpr32 : 0.f;
    }
    run = combine(run, e[j]);
  }
  // block reduction
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    aupr += __shfl_xor_sync(0xffffffffu, aupr, off);
    num += __shfl_xor_sync(0xffffffffu, num, off);
    first95 = min(first95, (unsigned long long)__shfl_xor_sync(0xffffffffu, first95, off));
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) {
    red[warp] = aupr;
    redn[warp] = num;
    if (first95 != ~0ull) atomicMin(&mo->first95, first95);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0;
    unsigned long long m = 0;
    for (int w = 0; w < CV_THREADS / 32; ++w) {
      a += red[w];
      m += redn[w];
    }
    aupr_partial[blockIdx.x] = a;
    atomicAdd(&mo->auroc_num, m);  // integer: exact, order-independent
    if (blockIdx.x == gridDim.x - 1) mo->n_points = total.nB + tile_state[blockIdx.x].nB;
    if (blockIdx.x == 0) {
      if (fpr_out) fpr_out[0] = 0.f;
      if (tpr_out) tpr_out[0] = 0.f;
    }
  }
}

// out[0] = auroc, out[1] = fpr@95, out[2] = aupr, out[3] = number of ROC points (incl. the prepended origin)
__global__ void __launch_bounds__(256) metrics_finish_kernel(const MetricsOut *__restrict__ mo,
                                                             const double *__restrict__ aupr_partial, int64_t nblk,
                                                             uint32_t P, uint32_t Nn, double *__restrict__ out) {
  __shared__ double red[256];
  double a = 0.0;
  for (int64_t b = threadIdx.x; b < nblk; b += 256) a += aupr_partial[b];  // fixed partition and tree: deterministic
  red[threadIdx.x] = a;
  __syncthreads();
  for (int off = 128; off > 0; off >>= 1) {
    if ((int)threadIdx.x < off) red[threadIdx.x] += red[threadIdx.x + off];
    __syncthreads();
  }
  if (threadIdx.x != 0) return;
  out[0] = (P && Nn) ? (double)mo->auroc_num / (2.0 * (double)P * (double)Nn) : 0.0;
  const uint32_t fps95 = (uint32_t)(mo->first95 & 0xffffffffull);
  out[1] = (mo->first95 != ~0ull && Nn) ? (double)__fdiv_rn((float)fps95, (float)Nn) : 1.0;
  out[2] = red[0];
  out[3] = (double)mo->n_points + 1.0;
}

static inline size_t align256(size_t b) { return (b + 255) / 256 * 256; }

struct MetricsLayout {
  size_t keys_a, keys_b, vals_a, vals_b, hist, hist_tiles, tile_state, aupr_partial, mo, flag, total;
};
static MetricsLayout metrics_layout(int64_t n) {
  MetricsLayout L;
  size_t o = 0;
  auto take = [&](size_t bytes) {
    const size_t at = o;
    o += align256(bytes);
    return at;
  };
  const int64_t nblk = ceil_div(n, RS_TILE);
  L.keys_a = take((size_t)n * 8);
  L.keys_b = take((size_t)n * 8);
  L.vals_a = take((size_t)n * 4);
  L.vals_b = take((size_t)n * 4);
  L.hist = take((size_t)256 * nblk * 4);
  L.hist_tiles = take((size_t)ceil_div(256 * nblk, SC_TILE) * 4 + 4);
  L.tile_state = take((size_t)ceil_div(n, CV_TILE) * sizeof(CurveState));
  L.aupr_partial = take((size_t)ceil_div(n, CV_TILE) * 8);
  L.mo = take(sizeof(MetricsOut));
  L.flag = take(4);
  L.total = o;
  return L;
}

template <typename T>
static int ood_metrics_impl(const T *ind, int64_t n_ind, const T *ood, int64_t n_ood, double *out4, float *fpr_out,
                            float *tpr_out, void *workspace, int64_t workspace_bytes, cudaStream_t st) {
  const int64_t n = n_ind + n_ood;
  RUNIA_REQUIRE(n_ind > 0 && n_ood > 0, RUNIA_E_BADARG, "ood_metrics: both score arrays must be non-empty");
  RUNIA_REQUIRE(n < (int64_t)0x7fffffff, RUNIA_E_UNSUPPORTED, "ood_metrics: more than 2^31 - 1 scores");
  RUNIA_REQUIRE(ind && ood && out4 && workspace, RUNIA_E_BADARG, "ood_metrics: null pointer");
  const MetricsLayout L = metrics_layout(n);
  RUNIA_REQUIRE((size_t)workspace_bytes >= L.total, RUNIA_E_WORKSPACE, "ood_metrics: workspace %lld < %lld bytes",
                (long long)workspace_bytes, (long long)L.total);
  unsigned char *ws = (unsigned char *)workspace;
  uint64_t *ka = (uint64_t *)(ws + L.keys_a), *kb = (uint64_t *)(ws + L.keys_b);
  uint32_t *va = (uint32_t *)(ws + L.vals_a), *vb = (uint32_t *)(ws + L.vals_b);
  uint32_t *hist = (uint32_t *)(ws + L.hist), *hist_tiles = (uint32_t *)(ws + L.hist_tiles);
  CurveState *tile_state = (CurveState *)(ws + L.tile_state);
  double *aupr_partial = (double *)(ws + L.aupr_partial);
  MetricsOut *mo = (MetricsOut *)(ws + L.mo);
  uint32_t *flag = (uint32_t *)(ws + L.flag);

  static PerDeviceFlag attr;
  if (!attr) {
    RUNIA_CUDA(cudaFuncSetAttribute(rs_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kScatterSmem));
    attr = true;
  }
  RUNIA_CUDA(cudaMemsetAsync(flag, 0, 4, st));
  const unsigned g1 = (unsigned)std::min<int64_t>(ceil_div(n, 256), (int64_t)kNumSMs * 8);
  range_flag_kernel<T><<<g1, 256, 0, st>>>(ind, n_ind, ood, n_ood, flag);
  make_keys_kernel<T><<<g1, 256, 0, st>>>(ind, n_ind, ood, n_ood, flag, ka, va);
  int launches = 2;
  // float32 scores only occupy the high word of the key
  const int64_t nblk = ceil_div(n, RS_TILE);
  const int64_t hist_n = 256 * nblk, hist_nt = ceil_div(hist_n, SC_TILE);
  for (int shift = sizeof(T) == 8 ? 0 : 32; shift < 64; shift += 8) {
    rs_hist_kernel<<<(unsigned)nblk, RS_THREADS, 0, st>>>(ka, n, shift, hist, nblk);
    scan_u32_partial_kernel<<<(unsigned)hist_nt, SC_THREADS, 0, st>>>(hist, hist_n, hist_tiles);
    scan_u32_tiles_kernel<<<1, SC_THREADS, 0, st>>>(hist_tiles, hist_nt);
    scan_u32_final_kernel<<<(unsigned)hist_nt, SC_THREADS, 0, st>>>(hist, hist_n, hist_tiles);
    rs_scatter_kernel<<<(unsigned)nblk, RS_THREADS, kScatterSmem, st>>>(ka, va, n, shift, hist, nblk, kb, vb);
    std::swap(ka, kb);
    std::swap(va, vb);
    launches += 5;
  }
  // an even number of passes: the sorted pairs are back in (keys_a, vals_a)
  const int64_t nt = ceil_div(n, CV_TILE);
  MetricsOut init{0ull, ~0ull, 0u};
  RUNIA_CUDA(cudaMemcpyAsync(mo, &init, sizeof(init), cudaMemcpyHostToDevice, st));
  curve_partial_kernel<<<(unsigned)nt, CV_THREADS, 0, st>>>(ka, va, n, tile_state);
  curve_tiles_kernel<<<1, CV_THREADS, 0, st>>>(tile_state, nt);
  curve_final_kernel<<<(unsigned)nt, CV_THREADS, 0, st>>>(ka, va, n, tile_state, (uint32_t)n_ind, (uint32_t)n_ood, mo,
                                                         aupr_partial, fpr_out, tpr_out);
  metrics_finish_kernel<<<1, 256, 0, st>>>(mo, aupr_partial, nt, (uint32_t)n_ind, (uint32_t)n_ood, out4);
  count_launch(launches + 4);
  return finish_launch("ood_metrics");
}

// ------------------------------------------------------------------------------------------------
// (f2) ascending sort of float32 values with the same radix sort: the order statistics behind
// np.percentile(train.flatten(), p) in ReAct / DICE+ReAct setup (postprocessors.py:1433, 1576)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sort_keys_in_kernel(const float *__restrict__ x, int64_t n,
                                                           uint64_t *__restrict__ keys, uint32_t *__restrict__ vals) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    keys[i] = ordered_key(x[i]);
    vals[i] = 0u;
  }
}
__global__ void __launch_bounds__(256) sort_keys_out_kernel(const uint64_t *__restrict__ keys, int64_t n,
                                                            float *__restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint32_t k = (uint32_t)(keys[i] >> 32);
    out[i] = __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
  }
}

}  // namespace runia

using namespace runia;

extern "C" int64_t runia_sort_f32_workspace_bytes(int64_t n) { return n > 0 ? (int64_t)metrics_layout(n).total : 0; }

extern "C" int runia_sort_f32(const float *x, int64_t n, float *out_sorted, void *workspace, int64_t workspace_bytes,
                              void *stream) {
  RUNIA_NVTX();
  RUNIA_REQUIRE(n >= 0 && n < (int64_t)0x7fffffff, RUNIA_E_BADARG, "sort_f32: bad size");
  if (n == 0) return RUNIA_OK;
  RUNIA_REQUIRE(x && out_sorted && workspace, RUNIA_E_BADARG, "sort_f32: null pointer");
  const MetricsLayout L = metrics_layout(n);
  RUNIA_REQUIRE((size_t)workspace_bytes >= L.total, RUNIA_E_WORKSPACE, "sort_f32: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  unsigned char *ws = (unsigned char *)workspace;
  uint64_t *ka = (uint64_t *)(ws + L.keys_a), *kb = (uint64_t *)(ws + L.keys_b);
  uint32_t *va = (uint32_t *)(ws + L.vals_a), *vb = (uint32_t *)(ws + L.vals_b);
  uint32_t *hist = (uint32_t *)(ws + L.hist), *hist_tiles = (uint32_t *)(ws + L.hist_tiles);
  static PerDeviceFlag attr;
  if (!attr) {
    RUNIA_CUDA(cudaFuncSetAttribute(rs_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kScatterSmem));
    attr = true;
  }
  const unsigned g1 = (unsigned)std::min<int64_t>(ceil_div(n, 256), (int64_t)kNumSMs * 8);
  sort_keys_in_kernel<<<g1, 256, 0, st>>>(x, n, ka, va);
  const int64_t nblk = ceil_div(n, RS_TILE);
  const int64_t hist_n = 256 * nblk, hist_nt = ceil_div(hist_n, SC_TILE);
  for (int shift = 32; shift < 64; shift += 8) {
    rs_hist_kernel<<<(unsigned)nblk, RS_THREADS, 0, st>>>(ka, n, shift, hist, nblk);
    scan_u32_partial_kernel<<<(unsigned)hist_nt, SC_THREADS, 0, st>>>(hist, hist_n, hist_tiles);
    scan_u32_tiles_kernel<<<1, SC_THREADS, 0, st>>>(hist_tiles, hist_nt);
    scan_u32_final_kernel<<<(unsigned)hist_nt, SC_THREADS, 0, st>>>(hist, hist_n, hist_tiles);
    rs_scatter_kernel<<<(unsigned)nblk, RS_THREADS, kScatterSmem, st>>>(ka, va, n, shift, hist, nblk, kb, vb);
    std::swap(ka, kb);
    std::swap(va, vb);
  }
  sort_keys_out_kernel<<<g1, 256, 0, st>>>(ka, n, out_sorted);
  count_launch(22);
  return finish_launch("sort_f32");
}

extern "C" int64_t runia_ood_metrics_workspace_bytes(int64_t n_ind, int64_t n_ood) {
  if (n_ind <= 0 || n_ood <= 0) return 0;
  return (int64_t)metrics_layout(n_ind + n_ood).total;
}

extern "C" int runia_ood_metrics(const void *ind, int64_t n_ind, const void *ood, int64_t n_ood, int is_f64,
                                 double *out4, float *fpr_out, float *tpr_out, void *workspace,
                                 int64_t workspace_bytes, void *stream) {
  RUNIA_NVTX();
  if (is_f64)
    return ood_metrics_impl<double>((const double *)ind, n_ind, (const double *)ood, n_ood, out4, fpr_out, tpr_out,
                                    workspace, workspace_bytes, (cudaStream_t)stream);
  return ood_metrics_impl<float>((const float *)ind, n_ind, (const float *)ood, n_ood, out4, fpr_out, tpr_out, workspace,
                                 workspace_bytes, (cudaStream_t)stream);
}
