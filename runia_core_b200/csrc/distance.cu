// kNN (a5) and Gaussian-KDE (a4) scorers: fused FP32 distance-GEMM + per-row streaming
// top-k / online log-sum-exp.  The [Nq, Nb] distance matrix never reaches HBM.
//
// kNN exactness contract (SURVEY section 8a5 / DESIGN.md "kNN"): the fused pass only PROPOSES
// candidates; every reported neighbour is re-ranked with float64 distances accumulated in a
// fixed order that the oracle replays bit for bit, ties broken by index, and a row is accepted
// only when the bound "approximate distance of any rejected bank row - eps > exact k-th
// distance" proves that no rejected row could belong to the top-k.  Rows that fail the proof
// (near-duplicates at rank k) go through an exhaustive exact pass.
#include <algorithm>
#include <mutex>

#include "rowgemm.cuh"

namespace runia {

// --------------------------------------------------------------------------------------------
// fixed-order float64 reductions shared with the oracle (oracle_np.seq32_tree_sum)
// --------------------------------------------------------------------------------------------
__device__ __forceinline__ double exact_sqdist_warp(const float *__restrict__ a, const float *__restrict__ b,
                                                    int d, int lane) {
  double p = 0.0;
  for (int j = lane; j < d; j += 32) {
    const double df = __dsub_rn((double)a[j], (double)b[j]);
    p = __dadd_rn(p, __dmul_rn(df, df));
  }
  return warp_tree_sum_f64(p);
}

template <typename T>
__global__ void __launch_bounds__(256) normalize_rows_kernel(const T *__restrict__ in, int64_t N, int d,
                                                             float *__restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= N) return;
  const T *x = in + row * (int64_t)d;
  double p = 0.0;
  for (int j = lane; j < d; j += 32) {
    const double v = (double)x[j];
    p = __dadd_rn(p, __dmul_rn(v, v));
  }
  const double nrm = __dadd_rn(__dsqrt_rn(warp_tree_sum_f64(p)), 1e-10);
  float *o = out + row * (int64_t)d;
  for (int j = lane; j < d; j += 32) o[j] = (float)__ddiv_rn((double)x[j], nrm);
}

__global__ void __launch_bounds__(256) row_sqnorm_kernel(const float *__restrict__ X, int64_t N, int d,
                                                         float *__restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= N) return;
  const float *x = X + row * (int64_t)d;
  double p = 0.0;
  for (int j = lane; j < d; j += 32) {
    const double v = (double)x[j];
    p = fma(v, v, p);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) p += __shfl_xor_sync(0xffffffffu, p, off);
  if (lane == 0) out[row] = (float)p;
}

template <typename T>
__global__ void __launch_bounds__(256) center_cast_kernel(const T *__restrict__ in, int64_t total, int d,
                                                          const double *__restrict__ center,
                                                          float *__restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    const int j = (int)(e % d);
    if (sizeof(T) == 8) {
      const double v = (double)in[e];
      out[e] = (float)(center ? v - center[j] : v);
    } else {
      // float32 input: the reference subtracts a float32 mean in float32 (postprocessors.py:241)
      const float v = (float)in[e];
      out[e] = center ? v - (float)center[j] : v;
    }
  }
}

// --------------------------------------------------------------------------------------------
// warp-level bitonic sort of (key, idx) pairs in shared memory, ascending lexicographic order
// --------------------------------------------------------------------------------------------
template <typename KT, typename IT>
__device__ __forceinline__ void warp_bitonic_sort(KT *key, IT *idx, int n /* power of two */) {
  const int lane = threadIdx.x & 31;
  for (int k = 2; k <= n; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = lane; t < (n >> 1); t += 32) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const int p = i | j;
        const bool asc = ((i & k) == 0);
        const KT ka = key[i], kb = key[p];
        const IT ia = idx[i], ib = idx[p];
        const bool gt = (ka > kb) || (ka == kb && ia > ib);
        if (gt == asc) {
          key[i] = kb; key[p] = ka;
          idx[i] = ib; idx[p] = ia;
        }
      }
      __syncwarp();
    }
  }
}

// --------------------------------------------------------------------------------------------
// kNN stage 1: fused distance GEMM + streaming top-KCAP candidate filter
// --------------------------------------------------------------------------------------------

struct KnnPlan {
  int kseed;    // power of two >= k + 8: the seed's bound has at least this many bank rows below it
  int kcap;     // candidates that survive the merge (re-rank capacity): kseed for the FP32 pass, a multiple of it for
                // the single-TF32-product pass, whose wider rounding bound asks for more exact evaluations
  int fin_max;  // most entries the candidate pass may leave per (row, split); SIMT pass: exactly kcap (padded)
  int capp;     // buffer entries per (row, split): power of two, >= max(fin_max, 2 kcap + panel width)
  int splits;  // bank splits (grid.y)
  int64_t panels_per_split;
  int counted;  // 1: the pass wrote the list lengths to `counts` (tensor-core pass); 0: lists are padded to kcap
  int groups;   // tensor-core pass: query-tile groups; the re-rank of group g runs beside the candidate pass of g + 1
};

__global__ void __launch_bounds__(GEMM_THREADS, 2)
knn_candidates_kernel(const float *__restrict__ Q, const float *__restrict__ qn, int64_t Nq,
                      const float *__restrict__ B, const float *__restrict__ bn, int64_t Nb, int d,
                      KnnPlan plan, float *__restrict__ buf_d, int32_t *__restrict__ buf_i) {
  __shared__ GemmSmem sm;
  __shared__ float thr[BM];
  __shared__ int cnt[BM];
  extern __shared__ unsigned char dyn[];  // per-warp sort scratch: 8 x capp x (float + int)
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int split = blockIdx.y;
  const int capp = plan.capp, kcap = plan.kcap;
  const int trigger = capp - BN;
  float *skey = reinterpret_cast<float *>(dyn) + (size_t)warp * capp;
  int32_t *sidx = reinterpret_cast<int32_t *>(dyn + (size_t)8 * capp * sizeof(float)) + (size_t)warp * capp;

  if (threadIdx.x < BM) {
    thr[threadIdx.x] = INFINITY;
    cnt[threadIdx.x] = 0;
  }
  const Prologue pro{nullptr, INFINITY};
  float qnr[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t row = m0 + tile_row(ty, i);
    qnr[i] = row < Nq ? __ldg(qn + row) : 0.f;
  }
  const int64_t b_lo = (int64_t)split * plan.panels_per_split * BN;
  int64_t b_hi = b_lo + plan.panels_per_split * BN;
  if (b_hi > Nb) b_hi = Nb;

  auto row_buf = [&](int r) -> size_t { return ((size_t)(m0 + r) * plan.splits + split) * capp; };
  // one warp compacts one row: sort the buffer, keep the kcap smallest, tighten the threshold
  auto compact = [&](int r, bool final_pass) {
    const int n = cnt[r];
    const size_t base = row_buf(r);
    for (int e = lane; e < capp; e += 32) {
      skey[e] = e < n ? buf_d[base + e] : INFINITY;
      sidx[e] = e < n ? buf_i[base + e] : 0x7fffffff;
    }
    __syncwarp();
    warp_bitonic_sort(skey, sidx, capp);
    const int keep = n < kcap ? n : kcap;
    const int wr = final_pass ? kcap : keep;
    for (int e = lane; e < wr; e += 32) {
      buf_d[base + e] = skey[e];
      buf_i[base + e] = e < keep ? sidx[e] : -1;
    }
    if (lane == 0) {
      cnt[r] = keep;
      if (keep == kcap) thr[r] = skey[kcap - 1];
    }
    __syncwarp();
  };

  for (int64_t n0 = b_lo; n0 < b_hi; n0 += BN) {
    float acc[8][8];
    zero_acc(acc);
    gemm_mainloop(Q, Nq, m0, B, b_hi, n0, d, pro, sm, acc);  // rows >= b_hi read as zero
    float bnr[8];
    int64_t colj[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      colj[j] = n0 + tile_col(tx, j);
      bnr[j] = colj[j] < b_hi ? __ldg(bn + colj[j]) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = tile_row(ty, i);
      if (m0 + r >= Nq) continue;
      const float th = thr[r];
      const size_t base = row_buf(r);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float dist = fmaf(-2.f, acc[i][j], qnr[i] + bnr[j]);
        if (colj[j] < b_hi && dist < th) {
          const int pos = atomicAdd(&cnt[r], 1);
          buf_d[base + pos] = dist;
          buf_i[base + pos] = (int32_t)colj[j];
        }
      }
    }
    __syncthreads();
    for (int r = warp; r < BM; r += 8)
      if (cnt[r] > trigger) compact(r, false);
    // the next gemm_mainloop starts with __syncthreads(): thr/cnt updates are visible
  }
  __syncthreads();
  for (int r = warp; r < BM; r += 8)
    if (m0 + r < Nq) compact(r, true);
}

// --------------------------------------------------------------------------------------------
// kNN stage 2: merge the per-split candidate lists, exact float64 re-rank, certification
// --------------------------------------------------------------------------------------------
struct KnnRerankArgs {
  const float *Q, *B;
  int64_t row0;  // this launch covers query rows [row0, Nq)
  int64_t Nq, Nb;
  int d, k;
  KnnPlan plan;
  const float *buf_d;
  const int32_t *buf_i;
  const int32_t *counts;  // [Nq, splits] list lengths (plan.counted) or nullptr
  float eps;                 // rounding bound of the approximate pass per unit of (|q|^2 + max_b |b|^2) / 2
  const float *thr_fin;      // [Nq, splits] final threshold of every list of the tensor-core pass, or nullptr
  const float *qn;           // [Nq] |q|^2
  const uint32_t *bn_max;    // [1] bit pattern of max_b |b|^2 (non-negative floats order like their bits)
  int64_t idx_offset;
  float *out_dist;
  double *out_dist64;
  int64_t *out_idx;
  float *out_kth;
  int32_t *flag_count;   // [1]
  int32_t *flag_rows;    // [Nq]
  double *flag_T;        // [Nq] exact k-th distance among candidates
  int32_t *flag_I;       // [Nq] its index
};

constexpr int RERANK_WARPS = 8;  // at most; fewer for large kcap (the per-warp scratch is 40 kcap bytes)
__host__ __device__ inline size_t knn_rerank_scratch(int kcap) { return (size_t)2 * kcap * 20; }  // multiple of 16

__global__ void __launch_bounds__(RERANK_WARPS * 32) knn_rerank_kernel(KnnRerankArgs a) {
  extern __shared__ unsigned char dyn[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t row = a.row0 + (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
  if (row >= a.Nq) return;
  const int kcap = a.plan.kcap, S = a.plan.splits, capp = a.plan.capp;
  const int nbuf_max = 2 * kcap;
  // per-warp scratch: selection buffer keys / idx [2 kcap], exact keys / idx [2 kcap]
  unsigned char *base = dyn + (size_t)warp * knn_rerank_scratch(kcap);
  float *akey = reinterpret_cast<float *>(base);
  int32_t *aidx = reinterpret_cast<int32_t *>(base + (size_t)nbuf_max * 4);
  double *ekey = reinterpret_cast<double *>(base + (size_t)nbuf_max * 8);
  int32_t *eidx = reinterpret_cast<int32_t *>(base + (size_t)nbuf_max * 16);

  // one (distance, index) entry of list `s` per lane; index -1 = none
  auto load_entry = [&](int s, int c, int n_s, float &kd, int32_t &ki) {
    kd = INFINITY;
    ki = -1;
    if (c < n_s) {
      const size_t p0 = ((size_t)row * S + s) * capp;
      if (a.counts) {  // tensor-core pass: interleaved (distance, index) pairs
        const float2 e = reinterpret_cast<const float2 *>(a.buf_d)[p0 + c];
        kd = e.x;
        ki = __float_as_int(e.y);
      } else {         // SIMT pass: separate arrays, lists padded with index -1
        ki = a.buf_i[p0 + c];
        kd = a.buf_d[p0 + c];
      }
    }
  };

  // Streaming selection of the kcap smallest approximate distances over the row's per-split lists
  // (row-major [row][split][entry]: coalesced loads).  Entries below the running threshold are
  // appended to a 2*kcap buffer in shared memory; when it would overflow it is sorted, cut back to
  // its kcap smallest and the threshold drops to the largest of them.  Entries equal to the
  // threshold are rejected: they tie with the kcap-th kept value, which is all the certification
  // below needs (every rejected entry has approximate distance >= it).
  for (int e = lane; e < nbuf_max; e += 32) {
    akey[e] = INFINITY;
    aidx[e] = 0x7fffffff;
  }
  __syncwarp();
  int n_buf = 0, n_listed = 0;
  float thr = INFINITY;
  auto cut = [&]() {  // sort the buffer, keep its kcap smallest
    warp_bitonic_sort(akey, aidx, nbuf_max);
    if (n_buf >= kcap) {
      n_buf = kcap;
      thr = akey[kcap - 1];
      for (int e = kcap + lane; e < nbuf_max; e += 32) {
        akey[e] = INFINITY;
        aidx[e] = 0x7fffffff;
      }
    }
    __syncwarp();
  };
  for (int s = 0; s < S; ++s) {
    const int n_s = a.counts ? a.counts[(size_t)row * S + s] : kcap;
    for (int c0 = 0; c0 < n_s; c0 += 32) {
      float kd;
      int32_t ki;
      load_entry(s, c0 + lane, n_s, kd, ki);
      n_listed += __popc(__ballot_sync(0xffffffffu, ki >= 0));
      bool take = ki >= 0 && kd < thr;
      unsigned m = __ballot_sync(0xffffffffu, take);
      if (n_buf + __popc(m) > nbuf_max) {
        cut();
        take = ki >= 0 && kd < thr;
        m = __ballot_sync(0xffffffffu, take);
      }
      if (take) {
        const int pos = n_buf + __popc(m & ((1u << lane) - 1u));
        akey[pos] = kd;
        aidx[pos] = ki;
      }
      n_buf += __popc(m);
      __syncwarp();
    }
  }
  cut();  // survivors sorted ascending in [0, min(n_buf, kcap))
  int n_real = 0;
  for (int e = lane; e < kcap; e += 32) n_real += (aidx[e] != 0x7fffffff) ? 1 : 0;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) n_real += __shfl_xor_sync(0xffffffffu, n_real, off);
  const bool list_full = (n_real == kcap);

  // Lower bound on the approximate distance of every bank row that is in NO list.  Tensor-core pass: the smallest of
  // the splits' final thresholds (the seed bound, lowered where a split had to shrink its list; +inf for an unseeded
  // split that never shrank: it listed its whole range).  SIMT pass: every split keeps its kcap smallest, so a full
  // merged list bounds the rest by its last entry, and a merged list that is not full holds the whole bank.
  float L_lists = INFINITY;
  if (a.thr_fin) {
    for (int s = lane; s < S; s += 32) L_lists = fminf(L_lists, a.thr_fin[(size_t)row * S + s]);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) L_lists = fminf(L_lists, __shfl_xor_sync(0xffffffffu, L_lists, off));
  } else if (list_full) {
    L_lists = akey[kcap - 1];
  }

  // Which entries need their exact distance?  With |approx - exact| <= s, every true neighbour has an approximate
  // distance <= a_k + 2 s (a_k = k-th smallest approximate distance: k rows are exactly <= a_k + s, so the exact k-th
  // is, and a row beyond a_k + 2 s is exactly > a_k + s).  The FP32 pass has s ~ 1e-4 and this is k plus the ties; the
  // single-TF32-product pass has s ~ 3e-3 per unit of squared norm and evaluates a few tens more.
  const double s_row = (double)a.eps * 0.5 * ((double)a.qn[row] + (double)__uint_as_float(*a.bn_max));
  int n_eval = n_real;
  float L_miss = INFINITY;  // smallest approximate distance among the LISTED entries that are not evaluated
  bool overflow = false;
  if (n_real > a.k) {
    const double cut_at = (double)akey[a.k - 1] + 2.0 * s_row;
    int c = 0;
    for (int e = lane; e < n_real; e += 32) c += ((double)akey[e] <= cut_at) ? 1 : 0;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) c += __shfl_xor_sync(0xffffffffu, c, off);
    if (c < n_real) {  // akey is sorted: the evaluated entries are [0, c), everything the selection dropped is beyond
      n_eval = c;
      L_miss = akey[c];
    } else if (n_listed > n_real && a.counts) {
      // all kcap selected entries lie within the band and the selection dropped others: a second pass over the lists
      // collects the whole band (unsorted; up to 2 kcap entries) and the smallest distance beyond it
      __syncwarp();
      n_buf = 0;
      float miss = INFINITY;
      for (int s = 0; s < S && !overflow; ++s) {
        const int n_s = a.counts[(size_t)row * S + s];
        for (int c0 = 0; c0 < n_s; c0 += 32) {
          float kd;
          int32_t ki;
          load_entry(s, c0 + lane, n_s, kd, ki);
          const bool in_band = ki >= 0 && (double)kd <= cut_at;
          if (ki >= 0 && !in_band) miss = fminf(miss, kd);
          const unsigned m = __ballot_sync(0xffffffffu, in_band);
          if (n_buf + __popc(m) > nbuf_max) {
            overflow = true;  // a neighbourhood denser than the scratch: the exhaustive pass decides this row
            break;
          }
          if (in_band) {
            const int pos = n_buf + __popc(m & ((1u << lane) - 1u));
            akey[pos] = kd;
            aidx[pos] = ki;
          }
          n_buf += __popc(m);
        }
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) miss = fminf(miss, __shfl_xor_sync(0xffffffffu, miss, off));
      L_miss = miss;
      n_eval = n_buf;
      __syncwarp();
    } else if (n_listed > n_real) {
      L_miss = akey[kcap - 1];  // SIMT pass: what the selection dropped ties with or exceeds the last kept entry
    }
  }
  const float L_rest = overflow ? -INFINITY : fminf(L_miss, L_lists);

  // exact float64 distances of the evaluated entries, four candidates at a time (independent loads in
  // flight); per candidate the summation order is that of exact_sqdist_warp / the oracle
  int npad = 32;
  while (npad < n_eval) npad <<= 1;  // <= 2 kcap
  for (int e = n_eval + lane; e < npad; e += 32) {
    ekey[e] = (double)INFINITY;
    eidx[e] = 0x7fffffff;
  }
  const float *q = a.Q + row * (int64_t)a.d;
  for (int c0 = 0; c0 < n_eval; c0 += 4) {
    const float *bp[4];
    int32_t bi[4];
    double acc[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      bi[u] = c0 + u < n_eval ? aidx[c0 + u] : 0x7fffffff;
      bp[u] = a.B + (int64_t)(bi[u] != 0x7fffffff ? bi[u] : 0) * a.d;
      acc[u] = 0.0;
    }
    for (int j = lane; j < a.d; j += 32) {
      const double qa = (double)q[j];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const double df = __dsub_rn(qa, (double)bp[u][j]);
        acc[u] = __dadd_rn(acc[u], __dmul_rn(df, df));
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const double ex = warp_tree_sum_f64(acc[u]);
      if (lane == 0 && c0 + u < npad) {
        ekey[c0 + u] = bi[u] != 0x7fffffff ? ex : (double)INFINITY;
        eidx[c0 + u] = bi[u];
      }
    }
  }
  __syncwarp();
  warp_bitonic_sort(ekey, eidx, npad);

  const int k = a.k;
  for (int e = lane; e < k; e += 32) {
    const bool real = e < n_eval;
    const double dv = real ? ekey[e] : (double)INFINITY;
    if (a.out_dist) a.out_dist[row * k + e] = real ? (float)dv : FLT_MAX;
    if (a.out_dist64) a.out_dist64[row * k + e] = dv;
    if (a.out_idx) a.out_idx[row * k + e] = real ? (int64_t)eidx[e] + a.idx_offset : -1;
  }
  if (lane == 0) {
    const bool have_k = n_eval >= k;
    if (a.out_kth) a.out_kth[row] = have_k ? (float)ekey[k - 1] : FLT_MAX;
    // |approx - exact| <= s_row = eps * (|q|^2 + |b|^2) / 2 (|q.b| <= (|q|^2 + |b|^2) / 2; the norms themselves are
    // rounded once): unit-norm rows (every built-in postprocessor) give eps itself, un-normalised FlatL2Index banks
    // scale it.  Every row that was not evaluated is exactly >= L_rest - s_row: the result stands if that is beyond the
    // exact k-th (L_rest = inf: the whole bank was evaluated).
    const bool certified = !(L_rest < INFINITY) || (have_k && (double)L_rest - s_row > ekey[k - 1]);
    if (!certified) {
      const int slot = atomicAdd(a.flag_count, 1);
      a.flag_rows[slot] = (int32_t)row;
      a.flag_T[slot] = have_k ? ekey[k - 1] : (double)INFINITY;
      a.flag_I[slot] = have_k ? eidx[k - 1] : 0x7fffffff;
    }
  }
}

// --------------------------------------------------------------------------------------------
// kNN stage 3: exhaustive exact pass for rows whose result could not be certified.
// Every true neighbour satisfies (dist, idx) <= (T, I) lexicographically, where (T, I) is the
// exact k-th candidate; collect those pairs, sort, emit the first k.
// --------------------------------------------------------------------------------------------
constexpr int FB_CAP = 4096;    // hit buffer per block
constexpr int FB_CHUNK = 2048;  // bank rows scanned between two overflow checks (a chunk adds at most that many hits)
constexpr int FB_THREADS = 256;
constexpr int FB_PER_THREAD = FB_CAP / FB_THREADS;

// Keeps the min(n, k) smallest (distance, index) pairs of the buffer, sorted ascending, in its first slots.
// Rank counting: the order is total (indices are distinct), so ranks are a permutation.
__device__ __forceinline__ void fb_select(double *hd, int32_t *hi, int n, int k) {
  double de[FB_PER_THREAD];
  int32_t ie[FB_PER_THREAD];
  int rk[FB_PER_THREAD];
#pragma unroll
  for (int u = 0; u < FB_PER_THREAD; ++u) {
    const int e = threadIdx.x + FB_THREADS * u;
    de[u] = e < n ? hd[e] : (double)INFINITY;
    ie[u] = e < n ? hi[e] : 0x7fffffff;
    rk[u] = 0;
  }
  for (int o = 0; o < n; ++o) {
    const double dd = hd[o];
    const int32_t io = hi[o];
#pragma unroll
    for (int u = 0; u < FB_PER_THREAD; ++u) rk[u] += (dd < de[u] || (dd == de[u] && io < ie[u])) ? 1 : 0;
  }
  __syncthreads();  // every thread has read the whole buffer
#pragma unroll
  for (int u = 0; u < FB_PER_THREAD; ++u) {
    const int e = threadIdx.x + FB_THREADS * u;
    if (e < n && rk[u] < k) {
      hd[rk[u]] = de[u];
      hi[rk[u]] = ie[u];
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(FB_THREADS) knn_fallback_kernel(KnnRerankArgs a, int32_t *status,
                                                                  double *fb_d, int32_t *fb_i) {
  __shared__ int n_hit;
  __shared__ double sT;
  __shared__ int32_t sI;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_flag = *a.flag_count;
  double *hd = fb_d + (size_t)blockIdx.x * FB_CAP;
  int32_t *hi = fb_i + (size_t)blockIdx.x * FB_CAP;
  for (int f = blockIdx.x; f < n_flag; f += gridDim.x) {
    const int64_t row = a.flag_rows[f];
    if (threadIdx.x == 0) {
      n_hit = 0;
      sT = a.flag_T[f];
      sI = a.flag_I[f];
    }
    const float *q = a.Q + row * (int64_t)a.d;
    for (int64_t c0 = 0; c0 < a.Nb; c0 += FB_CHUNK) {
      __syncthreads();
      // Any number of bank rows may tie with the k-th neighbour (all-zero activations normalise to identical
      // vectors): when the next chunk could overflow the buffer, cut it back to its k smallest pairs and tighten
      // the bound (T, I) to the k-th of them -- every true neighbour still satisfies (dist, idx) <= (T, I).
      if (n_hit > FB_CAP - FB_CHUNK) {  // block-uniform (read after the barrier)
        const int n = n_hit;
        fb_select(hd, hi, n, a.k);
        if (threadIdx.x == 0) {
          n_hit = a.k;
          sT = hd[a.k - 1];
          sI = hi[a.k - 1];
        }
        __syncthreads();
      }
      const double T = sT;
      const int32_t I = sI;
      const int64_t c1 = c0 + FB_CHUNK < a.Nb ? c0 + FB_CHUNK : a.Nb;
      for (int64_t b = c0 + warp; b < c1; b += FB_THREADS / 32) {
        const double ex = exact_sqdist_warp(q, a.B + b * a.d, a.d, lane);
        if (lane == 0 && (ex < T || (ex == T && (int32_t)b <= I))) {
          const int pos = atomicAdd(&n_hit, 1);
          hd[pos] = ex;
          hi[pos] = (int32_t)b;
        }
      }
    }
    __syncthreads();
    const int n = n_hit;  // >= k: the k candidates up to (T, I) are hits themselves
    fb_select(hd, hi, n, a.k);
    for (int e = threadIdx.x; e < a.k && e < n; e += blockDim.x) {
      const double de = hd[e];
      if (a.out_dist) a.out_dist[row * a.k + e] = (float)de;
      if (a.out_dist64) a.out_dist64[row * a.k + e] = de;
      if (a.out_idx) a.out_idx[row * a.k + e] = (int64_t)hi[e] + a.idx_offset;
      if (e == a.k - 1 && a.out_kth) a.out_kth[row] = (float)de;
    }
    if (threadIdx.x == 0) atomicAdd(&status[0], 1);
    __syncthreads();
  }
}

// max over the bank of |b|^2 (scales the certification bound); non-negative floats compare like their bit patterns
__global__ void __launch_bounds__(256) max_sqnorm_kernel(const float *__restrict__ bn, int64_t Nb, uint32_t *out) {
  float m = 0.f;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < Nb; e += (int64_t)gridDim.x * blockDim.x)
    m = fmaxf(m, __ldg(bn + e));
  m = warp_max32(m);
  if ((threadIdx.x & 31) == 0) atomicMax(out, __float_as_uint(m));
}

// --------------------------------------------------------------------------------------------
// merge of R sorted partial top-k lists (bank sharded across GPUs)
// --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) topk_merge_kernel(const double *__restrict__ pd,
                                                         const int64_t *__restrict__ pi, int R, int64_t Nq,
                                                         int k, float *out_dist, int64_t *out_idx,
                                                         float *out_kth) {
  // one thread per query row: R-way merge by head pointers (R <= 64)
  const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= Nq) return;
  int head[64];
  for (int r = 0; r < R; ++r) head[r] = 0;
  for (int e = 0; e < k; ++e) {
    int best = -1;
    double bd = INFINITY;
    int64_t bi = -1;
    for (int r = 0; r < R; ++r) {
      if (head[r] >= k) continue;
      const size_t p = ((size_t)r * Nq + row) * k + head[r];
      const int64_t ii = pi[p];
      if (ii < 0) continue;  // padding: this list is exhausted
      const double dd = pd[p];
      if (best < 0 || dd < bd || (dd == bd && ii < bi)) {
        best = r;
        bd = dd;
        bi = ii;
      }
    }
    if (best >= 0) head[best]++;
    if (out_dist) out_dist[row * k + e] = best >= 0 ? (float)bd : FLT_MAX;
    if (out_idx) out_idx[row * k + e] = best >= 0 ? bi : -1;
    if (e == k - 1 && out_kth) out_kth[row] = best >= 0 ? (float)bd : FLT_MAX;
  }
}

// --------------------------------------------------------------------------------------------
// (a4) KDE: fused distance GEMM + online log-sum-exp
// --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(GEMM_THREADS, 2)
kde_partial_kernel(const float *__restrict__ Q, const float *__restrict__ qn, int64_t Nq,
                   const float *__restrict__ B, const float *__restrict__ bn, int64_t Nb, int d,
                   float neg_half_inv_h2_log2e, int splits, int64_t panels_per_split,
                   float *__restrict__ part_m, float *__restrict__ part_s) {
  __shared__ GemmSmem sm;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int split = blockIdx.y;
  const Prologue pro{nullptr, INFINITY};
  float qnr[8], rm[8], rs[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t row = m0 + tile_row(ty, i);
    qnr[i] = row < Nq ? __ldg(qn + row) : 0.f;
    rm[i] = -INFINITY;
    rs[i] = 0.f;
  }
  const int64_t b_lo = (int64_t)split * panels_per_split * BN;
  int64_t b_hi = b_lo + panels_per_split * BN;
  if (b_hi > Nb) b_hi = Nb;
  for (int64_t n0 = b_lo; n0 < b_hi; n0 += BN) {
    float acc[8][8];
    zero_acc(acc);
    gemm_mainloop(Q, Nq, m0, B, b_hi, n0, d, pro, sm, acc);
    float bnr[8];
    bool ok[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int64_t col = n0 + tile_col(tx, j);
      ok[j] = col < b_hi;
      bnr[j] = ok[j] ? __ldg(bn + col) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float t[8];
      float tmax = -INFINITY;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        // squared distance, clamped at 0 (the expansion can go slightly negative)
        const float dist = fmaxf(fmaf(-2.f, acc[i][j], qnr[i] + bnr[j]), 0.f);
        t[j] = ok[j] ? dist * neg_half_inv_h2_log2e : -INFINITY;  // log2 units
        tmax = fmaxf(tmax, t[j]);
      }
      if (tmax == -INFINITY) continue;
      const float m_new = fmaxf(rm[i], tmax);
      float s = rs[i] * exp2f(rm[i] - m_new);
#pragma unroll
      for (int j = 0; j < 8; ++j) s += exp2f(t[j] - m_new);
      rs[i] = s;
      rm[i] = m_new;
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float M = warp_max16(rm[i]);
    const float s = (rm[i] == -INFINITY) ? 0.f : rs[i] * exp2f(rm[i] - M);
    const float S = warp_sum16(s);
    const int64_t row = m0 + tile_row(ty, i);
    if (tx == 0 && row < Nq) {
      part_m[(size_t)row * splits + split] = M;
      part_s[(size_t)row * splits + split] = S;
    }
  }
}

__global__ void __launch_bounds__(256) kde_finalize_kernel(const float *__restrict__ part_m,
                                                           const float *__restrict__ part_s, int64_t Nq,
                                                           int splits, double log_norm, double *out64,
                                                           float *out_max, float *out_sum) {
  const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= Nq) return;
  float M = -INFINITY;
  for (int s = 0; s < splits; ++s) M = fmaxf(M, part_m[(size_t)row * splits + s]);
  double S = 0.0;
  for (int s = 0; s < splits; ++s) {
    const float m = part_m[(size_t)row * splits + s];
    if (m != -INFINITY) S += (double)part_s[(size_t)row * splits + s] * exp2((double)m - (double)M);
  }
  const double LN2 = 0.693147180559945309417232121458;
  if (out64) out64[row] = (double)M * LN2 + log(S) - log_norm;
  if (out_max) out_max[row] = (float)((double)M * LN2);  // natural-log units
  if (out_sum) out_sum[row] = (float)S;
}

// Bank splits (grid.y): pick the split count that fills whole waves of the resident-CTA grid best.
static void pick_splits(int64_t rowblocks, int64_t panels, int max_splits, int64_t slots, int &splits, int64_t &pps) {
  int64_t best_s = 1;
  double best_u = -1.0;
  for (int64_t s = 1; s <= max_splits && s <= (panels > 0 ? panels : 1); ++s) {
    const int64_t eff = ceil_div(panels, ceil_div(panels, s));  // splits actually produced
    const int64_t total = rowblocks * eff;
    const double u = (double)total / (double)(ceil_div(total, slots) * slots);
    if (u > best_u + 0.02) {
      best_u = u;
      best_s = s;
    }
  }
  pps = ceil_div(panels > 0 ? panels : 1, best_s);
  splits = (int)ceil_div(panels > 0 ? panels : 1, pps);
  if (splits < 1) splits = 1;
}

// tensor-core pass: 256 x 256 panels per CTA pair, 74 pairs;  SIMT pass: 128 x 128 panels, 2 CTAs / SM
static KnnPlan make_knn_plan(int64_t Nq, int64_t Nb, int k, bool tensor, int products = 1) {
  KnnPlan p;
  p.kseed = 64;
  while (p.kseed < k + 8) p.kseed <<= 1;
  // selection capacity of the re-rank (and the floor a shrinking list keeps): the single-product pass certifies
  // through a ~50x wider rounding bound than the FP32 pass, so more rows lie within it of the k-th distance
  const bool filter1 = tensor && products == 1;
  p.kcap = filter1 ? std::min(1024, 2 * p.kseed) : p.kseed;
  const int pw = tensor ? 256 : BN;
  p.capp = 256;
  // tensor-core pass: room for a band of up to 2 kcap + 512 rows around the k-th distance before a list has to shrink
  // (a shrink lowers the list's final threshold into the band and sends the row to the exhaustive pass)
  while (p.capp < (filter1 ? std::min(4 * p.kcap, 2 * p.kcap + 512) : tensor ? 2 * p.kcap : p.kcap) + pw) p.capp <<= 1;
  if (tensor && p.capp < (filter1 ? 1024 : 512)) p.capp = filter1 ? 1024 : 512;
  p.counted = tensor ? 1 : 0;
  pick_splits(ceil_div(Nq, tensor ? 256 : BM), ceil_div(Nb, pw), 16, tensor ? kNumSMs / 2 : 2 * kNumSMs, p.splits,
              p.panels_per_split);
  // Query groups (tensor-core pass): the re-rank of group g runs on a side stream beside the candidate pass of group
  // g + 1.  The two cannot share an SM (the candidate pass takes all of its shared memory), so what a group hides is
  // the tail of a wave, and each group costs two launches and a pipeline fill: the largest of 1..4 groups whose
  // launches still fill whole waves of the 74 CTA-pair slots, and only for searches long enough to notice.
  p.groups = 1;
  if (tensor) {
    const int64_t tiles = ceil_div(Nq, 256), slots = kNumSMs / 2;
    double best = 0.0;
    for (int g = 1; g <= 4 && g <= tiles; ++g) {
      int64_t waves = 0;
      for (int i = 0; i < g; ++i) {
        const int64_t t = tiles / g + (i < tiles % g ? 1 : 0);
        waves += ceil_div(t * p.splits, slots);
      }
      const double eff = (double)(tiles * p.splits) / (double)(waves * slots);
      if (g > 1 && (double)waves * (double)p.panels_per_split < 2000.0) break;  // configs[1]: ~110 panel steps
      if (eff >= best - 0.03) {  // prefer more groups unless they cost more than 3 % of wave efficiency
        if (eff > best) best = eff;
        p.groups = g;
      }
    }
  }
  p.fin_max = tensor ? p.capp : p.kcap;
  return p;
}

struct KnnWorkspace {
  size_t qn, thr_key, thr_fin, counts, buf_d, buf_i, flag_count, flag_rows, flag_T, flag_I, fb_d, fb_i, total;
};
constexpr int FB_GRID = 64;
constexpr int kKnnMaxK = 1016;  // kcap = 1024 candidates per row survive the merge
static KnnWorkspace knn_layout(int64_t Nq, const KnnPlan &p) {
  KnnWorkspace w;
  size_t o = 0;
  auto take = [&](size_t bytes) {
    const size_t at = o;
    o += (bytes + 255) / 256 * 256;
    return at;
  };
  w.qn = take((size_t)Nq * 4);
  w.thr_key = take((size_t)Nq * 4);
  w.thr_fin = take((size_t)Nq * p.splits * 4);
  w.counts = take((size_t)Nq * p.splits * 4);
  const size_t rows32 = (size_t)ceil_div(Nq, 32) * 32;
  w.buf_d = take(rows32 * p.splits * p.capp * 4);
  w.buf_i = take(rows32 * p.splits * p.capp * 4);
  w.flag_count = take(256);
  w.flag_rows = take((size_t)Nq * 4);
  w.flag_T = take((size_t)Nq * 8);
  w.flag_I = take((size_t)Nq * 4);
  w.fb_d = take((size_t)FB_GRID * FB_CAP * 8);
  w.fb_i = take((size_t)FB_GRID * FB_CAP * 4);
  w.total = o;
  return w;
}

// One side stream + events per device for the fork / join inside runia_knn_search_f32 (re-rank of one query group
// beside the candidate pass of the next).  Work on it is ordered after an event of the caller's stream and the
// caller's stream waits for it before the call's last kernel, so the entry point stays "asynchronous on `stream`".
struct SideStream {
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr}, done = nullptr;
  static SideStream *get() {
    static SideStream per_dev[16];
    static std::mutex mu;
    int dev = 0;
    cudaGetDevice(&dev);
    SideStream &s = per_dev[dev & 15];
    std::lock_guard<std::mutex> lk(mu);
    if (!s.stream) {
      if (cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
      for (auto &e : s.ev)
        if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return nullptr;
      if (cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    }
    return &s;
  }
};

namespace tc {
bool usable(const void *A, int K, const void *B_hi, const void *B_lo);
int launch_knn_candidates_tc(const float *Q, const float *qn, int64_t Nq, const float *B_hi, const float *B_lo,
                             const float *bn, int64_t Nb, int d, int kseed, float seed_slack, int kcap, int fin_max, int capp,
                             int splits, int64_t panels_per_split, float *buf_d, int32_t *buf_i, int32_t *counts,
                             uint32_t *thr_key, float *thr_fin, const uint32_t *bn_max, int phase, int products,
                             cudaStream_t st);
int launch_kde_partial_tc(const float *Q, const float *qn, int64_t Nq, const float *B_hi, const float *B_lo,
                          const float *bn, int64_t Nb, int d, float scale, int splits, int64_t panels_per_split,
                          float *part_m, float *part_s, cudaStream_t st);
}  // namespace tc

}  // namespace runia

using namespace runia;

extern "C" int runia_center_cast(const void *in, int in_is_f64, int64_t N, int d, const double *center,
                                 float *out, void *stream) {
  RUNIA_NVTX();
  RUNIA_REQUIRE(N >= 0 && d > 0, RUNIA_E_BADARG, "center_cast: bad sizes");
  if (N == 0) return RUNIA_OK;
  RUNIA_REQUIRE(in && out, RUNIA_E_BADARG, "center_cast: null pointer");
  const int64_t total = N * d;
  const unsigned grid = (unsigned)std::min<int64_t>(ceil_div(total, 256), (int64_t)kNumSMs * 16);
  if (in_is_f64)
    center_cast_kernel<double><<<grid, 256, 0, (cudaStream_t)stream>>>((const double *)in, total, d, center, out);
  else
    center_cast_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float *)in, total, d, center, out);
  count_launch();
  return finish_launch("center_cast");
}

extern "C" int runia_normalize_rows(const void *in, int in_is_f64, int64_t N, int d, float *out, void *stream) {
  RUNIA_NVTX();
  RUNIA_REQUIRE(N >= 0 && d > 0, RUNIA_E_BADARG, "normalize_rows: bad sizes");
  if (N == 0) return RUNIA_OK;
  RUNIA_REQUIRE(in && out, RUNIA_E_BADARG, "normalize_rows: null pointer");
  const unsigned grid = (unsigned)ceil_div(N, 8);
  if (in_is_f64)
    normalize_rows_kernel<double><<<grid, 256, 0, (cudaStream_t)stream>>>((const double *)in, N, d, out);
  else
    normalize_rows_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float *)in, N, d, out);
  count_launch();
  return finish_launch("normalize_rows");
}

extern "C" int runia_row_sqnorm_f32(const float *X, int64_t N, int d, float *out, void *stream) {
  RUNIA_NVTX();
  RUNIA_REQUIRE(N >= 0 && d > 0, RUNIA_E_BADARG, "row_sqnorm: bad sizes");
  if (N == 0) return RUNIA_OK;
  RUNIA_REQUIRE(X && out, RUNIA_E_BADARG, "row_sqnorm: null pointer");
  row_sqnorm_kernel<<<(unsigned)ceil_div(N, 8), 256, 0, (cudaStream_t)stream>>>(X, N, d, out);
  count_launch();
  return finish_launch("row_sqnorm");
}

extern "C" int64_t runia_knn_workspace_bytes(int64_t Nq, int64_t Nb, int d, int k) {
  if (Nq <= 0 || Nb <= 0 || k <= 0 || k > kKnnMaxK) return 0;
  (void)d;
  const size_t a = knn_layout(Nq, make_knn_plan(Nq, Nb, k, false)).total;
  const size_t b = knn_layout(Nq, make_knn_plan(Nq, Nb, k, true, 1)).total;
  const size_t c = knn_layout(Nq, make_knn_plan(Nq, Nb, k, true, 3)).total;
  return (int64_t)std::max(a, std::max(b, c));
}

extern "C" int runia_knn_search_ex_f32(const float *Qn, int64_t Nq, const float *Bn, const float *Bn_sqnorm,
                                       const float *Bn_hi, const float *Bn_lo, int64_t Nb, int d, int k,
                                       int64_t idx_offset, float *out_dist, double *out_dist_f64, int64_t *out_idx,
                                       float *out_kth, int32_t *status, void *workspace, int64_t workspace_bytes,
                                       int filter_products, void *stream);

extern "C" int runia_knn_search_f32(const float *Qn, int64_t Nq, const float *Bn, const float *Bn_sqnorm,
                                    const float *Bn_hi, const float *Bn_lo, int64_t Nb, int d, int k,
                                    int64_t idx_offset, float *out_dist,
                                    double *out_dist_f64, int64_t *out_idx, float *out_kth, int32_t *status,
                                    void *workspace, int64_t workspace_bytes, void *stream) {
  return runia_knn_search_ex_f32(Qn, Nq, Bn, Bn_sqnorm, Bn_hi, Bn_lo, Nb, d, k, idx_offset, out_dist, out_dist_f64, out_idx,
                                 out_kth, status, workspace, workspace_bytes, 1, stream);
}

extern "C" int runia_knn_search_ex_f32(const float *Qn, int64_t Nq, const float *Bn, const float *Bn_sqnorm,
                                       const float *Bn_hi, const float *Bn_lo, int64_t Nb, int d, int k,
                                       int64_t idx_offset, float *out_dist, double *out_dist_f64, int64_t *out_idx,
                                       float *out_kth, int32_t *status, void *workspace, int64_t workspace_bytes,
                                       int filter_products, void *stream) {
  RUNIA_NVTX();
  RUNIA_REQUIRE(filter_products == 1 || filter_products == 3, RUNIA_E_BADARG, "knn_search: filter_products=%d (1 or 3)",
                filter_products);
  RUNIA_REQUIRE(Nq >= 0 && Nb > 0 && d > 0, RUNIA_E_BADARG, "knn_search: bad sizes Nq=%lld Nb=%lld d=%d",
                (long long)Nq, (long long)Nb, d);
  RUNIA_REQUIRE(k >= 1 && k <= kKnnMaxK, RUNIA_E_UNSUPPORTED, "knn_search: k=%d outside [1, %d]", k, kKnnMaxK);
  RUNIA_REQUIRE(Nb < (int64_t)0x7fffffff, RUNIA_E_UNSUPPORTED, "knn_search: bank shard too large");
  if (Nq == 0) return RUNIA_OK;
  RUNIA_REQUIRE(Qn && Bn && Bn_sqnorm && status && workspace, RUNIA_E_BADARG, "knn_search: null pointer");
  const bool tensor = Bn_hi && Bn_lo && tc::usable(Qn, d, Bn_hi, Bn_lo) &&
                      (reinterpret_cast<uintptr_t>(Bn_sqnorm) & 15) == 0;  // the epilogue reads the norms as float4
  const KnnPlan plan = make_knn_plan(Nq, Nb, k, tensor, filter_products);
  const KnnWorkspace w = knn_layout(Nq, plan);
  RUNIA_REQUIRE((size_t)workspace_bytes >= w.total, RUNIA_E_WORKSPACE, "knn_search: workspace %lld < %lld bytes",
                (long long)workspace_bytes, (long long)w.total);
  cudaStream_t st = (cudaStream_t)stream;
  unsigned char *ws = (unsigned char *)workspace;
  float *qn = (float *)(ws + w.qn);
  float *buf_d = (float *)(ws + w.buf_d);
  int32_t *buf_i = (int32_t *)(ws + w.buf_i);
  int32_t *flag_count = (int32_t *)(ws + w.flag_count);

  RUNIA_CUDA(cudaMemsetAsync(flag_count, 0, 256, st));  // [0] flagged rows, [1] max_b |b|^2
  RUNIA_CUDA(cudaMemsetAsync(status, 0, 4 * sizeof(int32_t), st));
  row_sqnorm_kernel<<<(unsigned)ceil_div(Nq, 8), 256, 0, st>>>(Qn, Nq, d, qn);
  max_sqnorm_kernel<<<(unsigned)std::min<int64_t>(ceil_div(Nb, 256 * 8), 4 * kNumSMs), 256, 0, st>>>(
      Bn_sqnorm, Nb, (uint32_t *)flag_count + 1);

  KnnRerankArgs a;
  a.Q = Qn; a.B = Bn; a.row0 = 0; a.Nq = Nq; a.Nb = Nb; a.d = d; a.k = k; a.plan = plan;
  a.buf_d = buf_d; a.buf_i = buf_i;
  a.counts = plan.counted ? (const int32_t *)(ws + w.counts) : nullptr;
  // |approx - exact| per unit of (|q|^2 + max|b|^2) / 2 (DESIGN.md "kNN certification"; the re-rank kernel applies
  // the scale, 1 for unit-norm rows):
  //   FP32 SIMT pass: accumulation and the norm terms, (2K + 8) * 2^-24 * 1.25;
  //   single TF32 product: the operands first -- A_hi truncates to 19 bits (2^-10 relative), B_hi rounds to nearest
  //   (2^-11), on sum |q_i b_i| <= (|q|^2 + |b|^2) / 2 and doubled by the -2 q.b term -- then the same accumulation.
  //   3xTF32: operand split error 2^-20 per unit of sum |q_i b_i| plus the accumulation -> 2.5 x the FP32 bound.
  const double acc_eps = (2.0 * d + 8.0) * 5.9604644775390625e-08;
  const double op_eps = 2.0 * (9.765625e-4 + 4.8828125e-4 + 4.76837158203125e-7);
  const double eps1 = op_eps * 1.0005 + 2.5 * acc_eps;  // single product (also the seed's, in both modes)
  a.eps = (float)(!tensor ? 1.25 * acc_eps : filter_products == 1 ? eps1 : 2.5 * acc_eps);
  // widening of the seed bound B (a single product in both modes), per unit of (|q|^2 + max|b|^2) / 2: the lists must
  // hold every row whose candidate-pass distance is below a_k + 2 s_c.  Same product in both passes: a_k <= B, so 2 s_1.
  // 3xTF32 candidates: a_k <= (exact k-th) + s_3 <= B + s_1 + s_3, so s_1 + 3 s_3.
  const float seed_slack = (float)((filter_products == 1 ? 2.0 * eps1 : eps1 + 3.0 * 2.5 * acc_eps) * 1.0001);
  a.thr_fin = tensor ? (const float *)(ws + w.thr_fin) : nullptr;
  a.qn = qn;
  a.bn_max = (const uint32_t *)flag_count + 1;
  a.idx_offset = idx_offset;
  a.out_dist = out_dist; a.out_dist64 = out_dist_f64; a.out_idx = out_idx; a.out_kth = out_kth;
  a.flag_count = flag_count;
  a.flag_rows = (int32_t *)(ws + w.flag_rows);
  a.flag_T = (double *)(ws + w.flag_T);
  a.flag_I = (int32_t *)(ws + w.flag_I);
  const size_t per_warp = knn_rerank_scratch(plan.kcap);
  const int rw = plan.kcap <= 128 ? RERANK_WARPS : plan.kcap <= 256 ? 4 : plan.kcap <= 512 ? 2 : 1;  // <= 40 KB per block
  const size_t dyn2 = per_warp * rw;
  static PerDeviceFlag attr2;
  if (!attr2) {
    RUNIA_CUDA(cudaFuncSetAttribute(knn_rerank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    attr2 = true;
  }
  auto rerank = [&](int64_t r0, int64_t r1, cudaStream_t s2) {
    KnnRerankArgs g = a;
    g.row0 = r0;
    g.Nq = r1;
    knn_rerank_kernel<<<(unsigned)ceil_div(r1 - r0, rw), rw * 32, dyn2, s2>>>(g);
    count_launch();
  };

  if (tensor) {
    // Query groups: the candidate pass of group g + 1 (tensor-bound, one CTA pair per SM pair) runs while the
    // re-rank of group g (gather- and latency-bound, small blocks) runs on a side stream (make_knn_plan picks the
    // number of groups together with the bank splits).
    const int64_t tiles = ceil_div(Nq, 256);
    int groups = plan.groups;
    SideStream *side = groups > 1 ? SideStream::get() : nullptr;
    if (groups > 1 && !side) groups = 1;
    if (groups > 1) {  // the seed thresholds of every row in one launch (a per-group seed would pay its latency per group)
      const int rc = tc::launch_knn_candidates_tc(Qn, qn, Nq, Bn_hi, Bn_lo, Bn_sqnorm, Nb, d, plan.kseed,
                                                  seed_slack, plan.kcap, plan.fin_max, plan.capp, plan.splits,
                                                  plan.panels_per_split, buf_d, buf_i, (int32_t *)(ws + w.counts),
                                                  (uint32_t *)(ws + w.thr_key), (float *)(ws + w.thr_fin),
                                                  (const uint32_t *)flag_count + 1, 1, filter_products, st);
      if (rc) return rc;
    }
    int64_t t0 = 0;
    for (int g = 0; g < groups; ++g) {
      const int64_t t1 = t0 + tiles / groups + (g < tiles % groups ? 1 : 0);
      const int64_t r0 = t0 * 256, r1 = std::min<int64_t>(Nq, t1 * 256);
      const size_t boff = (size_t)r0 * plan.splits * plan.capp;
      const int rc = tc::launch_knn_candidates_tc(Qn + r0 * d, qn + r0, r1 - r0, Bn_hi, Bn_lo, Bn_sqnorm, Nb, d, plan.kseed,
                                                  seed_slack, plan.kcap,
                                                  plan.fin_max, plan.capp, plan.splits, plan.panels_per_split,
                                                  buf_d + 2 * boff, buf_i, (int32_t *)(ws + w.counts) + r0 * plan.splits,
                                                  (uint32_t *)(ws + w.thr_key) + r0,
                                                  (float *)(ws + w.thr_fin) + r0 * plan.splits,
                                                  (const uint32_t *)flag_count + 1, groups > 1 ? 2 : 3, filter_products, st);
      if (rc) return rc;
      if (side && g + 1 < groups) {  // the last group's re-rank has nothing left to hide behind
        RUNIA_CUDA(cudaEventRecord(side->ev[g], st));
        RUNIA_CUDA(cudaStreamWaitEvent(side->stream, side->ev[g], 0));
        rerank(r0, r1, side->stream);
      } else {
        rerank(r0, r1, st);
      }
      t0 = t1;
    }
    if (side && groups > 1) {  // join: the exhaustive pass needs every group's flags
      RUNIA_CUDA(cudaEventRecord(side->done, side->stream));
      RUNIA_CUDA(cudaStreamWaitEvent(st, side->done, 0));
    }
  } else {
    const size_t dyn1 = (size_t)8 * plan.capp * 8;
    static PerDeviceFlag attr1;
    if (!attr1) {
      RUNIA_CUDA(cudaFuncSetAttribute(knn_candidates_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 140 * 1024));
      attr1 = true;
    }
    dim3 grid1((unsigned)ceil_div(Nq, BM), (unsigned)plan.splits);
    knn_candidates_kernel<<<grid1, GEMM_THREADS, dyn1, st>>>(Qn, qn, Nq, Bn, Bn_sqnorm, Nb, d, plan, buf_d, buf_i);
    count_launch();
    rerank(0, Nq, st);
  }
  knn_fallback_kernel<<<FB_GRID, FB_THREADS, 0, st>>>(a, status, (double *)(ws + w.fb_d), (int32_t *)(ws + w.fb_i));
  count_launch(3);
  return finish_launch("knn_search");
}

extern "C" int runia_topk_merge(const double *part_dist, const int64_t *part_idx, int R, int64_t Nq, int k,
                                float *out_dist, int64_t *out_idx, float *out_kth, void *stream) {
  RUNIA_NVTX();
  RUNIA_REQUIRE(R >= 1 && R <= 64 && Nq >= 0 && k >= 1, RUNIA_E_BADARG, "topk_merge: bad sizes R=%d k=%d", R, k);
  if (Nq == 0) return RUNIA_OK;
  RUNIA_REQUIRE(part_dist && part_idx, RUNIA_E_BADARG, "topk_merge: null pointer");
  topk_merge_kernel<<<(unsigned)ceil_div(Nq, 128), 128, 0, (cudaStream_t)stream>>>(part_dist, part_idx, R, Nq, k,
                                                                               out_dist, out_idx, out_kth);
  count_launch();
  return finish_launch("topk_merge");
}

static void kde_plan(int64_t Nq, int64_t Nb, bool tensor, int &splits, int64_t &pps) {
  pick_splits(ceil_div(Nq, tensor ? 256 : BM), ceil_div(Nb, tensor ? 256 : BN), 64, tensor ? kNumSMs / 2 : 2 * kNumSMs,
              splits, pps);
}

extern "C" int64_t runia_kde_workspace_bytes(int64_t Nq, int64_t Nb) {
  if (Nq <= 0 || Nb <= 0) return 0;
  int splits, s2;
  int64_t pps;
  kde_plan(Nq, Nb, false, splits, pps);
  kde_plan(Nq, Nb, true, s2, pps);
  if (s2 > splits) splits = s2;
  return (int64_t)(((size_t)Nq * 4 + 255) / 256 * 256 + ((size_t)Nb * 4 + 255) / 256 * 256 +
                   2 * (((size_t)Nq * splits * 4 + 255) / 256 * 256));
}

extern "C" int runia_kde_lse_f32(const float *Q, int64_t Nq, const float *B, const float *B_hi, const float *B_lo,
                                 int64_t Nb, int d, double bandwidth,
                                 int64_t Nb_total, double *out_f64, float *out_max, float *out_sum,
                                 void *workspace, int64_t workspace_bytes, void *stream) {
  RUNIA_NVTX();
  RUNIA_REQUIRE(Nq >= 0 && Nb > 0 && d > 0 && bandwidth > 0 && Nb_total >= Nb, RUNIA_E_BADARG, "kde_lse: bad sizes");
  if (Nq == 0) return RUNIA_OK;
  RUNIA_REQUIRE(Q && B && workspace && (out_f64 || (out_max && out_sum)), RUNIA_E_BADARG, "kde_lse: null pointer");
  RUNIA_REQUIRE(workspace_bytes >= runia_kde_workspace_bytes(Nq, Nb), RUNIA_E_WORKSPACE, "kde_lse: workspace too small");
  const bool tensor = B_hi && B_lo && tc::usable(Q, d, B_hi, B_lo);
  int splits;
  int64_t pps;
  kde_plan(Nq, Nb, tensor, splits, pps);
  cudaStream_t st = (cudaStream_t)stream;
  unsigned char *ws = (unsigned char *)workspace;
  size_t o = 0;
  float *qn = (float *)(ws + o); o += ((size_t)Nq * 4 + 255) / 256 * 256;
  float *bn = (float *)(ws + o); o += ((size_t)Nb * 4 + 255) / 256 * 256;
  float *pm = (float *)(ws + o); o += ((size_t)Nq * splits * 4 + 255) / 256 * 256;
  float *ps = (float *)(ws + o);
  row_sqnorm_kernel<<<(unsigned)ceil_div(Nq, 8), 256, 0, st>>>(Q, Nq, d, qn);
  row_sqnorm_kernel<<<(unsigned)ceil_div(Nb, 8), 256, 0, st>>>(B, Nb, d, bn);
  const double LOG2E = 1.4426950408889634073599246810019;
  const float scale = (float)(-0.5 / (bandwidth * bandwidth) * LOG2E);
  if (tensor) {
    const int rc = tc::launch_kde_partial_tc(Q, qn, Nq, B_hi, B_lo, bn, Nb, d, scale, splits, pps, pm, ps, st);
    if (rc) return rc;
  } else {
    dim3 grid((unsigned)ceil_div(Nq, BM), (unsigned)splits);
    kde_partial_kernel<<<grid, GEMM_THREADS, 0, st>>>(Q, qn, Nq, B, bn, Nb, d, scale, splits, pps, pm, ps);
  }
  const double log_norm = log((double)Nb_total) + 0.5 * d * log(2.0 * 3.14159265358979323846 * bandwidth * bandwidth);
  kde_finalize_kernel<<<(unsigned)ceil_div(Nq, 256), 256, 0, st>>>(pm, ps, Nq, splits, log_norm, out_f64, out_max, out_sum);
  count_launch(4);
  return finish_launch("kde_lse");
}
