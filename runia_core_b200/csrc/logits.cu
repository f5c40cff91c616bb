// HBM-bound single-pass scorers:
//  (a8)  Energy / MSP / GEN from logits in one read           postprocessors.py:519-691
//  (a10) ReAct / DICE / DICE+ReAct: clip -> linear -> LSE      postprocessors.py:1325-1621
//        ASH-S: per-row top-k pruning + rescale -> linear -> LSE   funcs.py:230-261
#include <algorithm>

#include "common.cuh"

namespace runia {

// ------------------------------------------------------------------------------------------
// logit scores, small C (<= 64): a block stages ROWS x C contiguous floats through shared memory
// with coalesced 128-bit loads, then one thread scores one row.
// ------------------------------------------------------------------------------------------
constexpr int LS_ROWS = 256;

__device__ __forceinline__ float gen_term(float p, float gamma) {
  // p^gamma * (1-p)^gamma, float32 like the reference (funcs.py:374)
  return powf(p, gamma) * powf(1.f - p, gamma);
}

constexpr float kLog2e = 1.4426950408889634f, kLn2f = 0.6931471805599453f;

// Softmax statistics in log2 units on the MUFU pipe: t_c = (l_c - max) log2 e, e_c = 2^t_c,
// s = sum e_c; energy = max + ln2 * log2 s; msp = 1 / s; log2 p_c = t_c - log2 s (no MUFU);
// GEN term (p (1 - p))^gamma = 2^(gamma (log2 p + log2(1 - p))).  31 MUFU per 10-class sample keeps the
// kernel under the HBM time (52 B/sample); powf / expf / logf would make it compute-bound 4x over.
// The approximate, flush-to-zero forms are safe here: e_c <= 1 only feeds s >= 1 and 1 - p, log2 p comes
// from the logits themselves, 1 - p is never denormal, and a flushed 2^x is below 1e-38.
__device__ __forceinline__ float ex2_fast(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float lg2_fast(float x) {
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// CMAX = 16: the row lives in registers (fully unrolled, guarded by c < C); CMAX = 64: in shared memory
template <int CMAX>
__global__ void __launch_bounds__(LS_ROWS)
logit_scores_small_kernel(const float *__restrict__ logits, int64_t N, int C, float gamma, int M,
                          float *__restrict__ energy, float *__restrict__ msp, float *__restrict__ gen) {
  extern __shared__ __align__(16) float tile[];  // [LS_ROWS * C] logits (CMAX = 64 + GEN: then the t_c)
  const int64_t r0 = (int64_t)blockIdx.x * LS_ROWS;
  const int rows = (int)((N - r0 < LS_ROWS) ? (N - r0) : LS_ROWS);
  const float *src = logits + r0 * C;
  // C <= 16 with 8-byte aligned rows: every thread reads its own row with LDG.64 (a warp covers a contiguous
  // 32 * C * 4 bytes, so every sector is used); otherwise the block stages its rows through shared memory
  const bool direct = CMAX <= 16 && (C % 2 == 0) && ((reinterpret_cast<uintptr_t>(logits) & 7) == 0);
  if (!direct) {
    const int total = rows * C;
    if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
      const int n4 = total >> 2;
      for (int i = threadIdx.x; i < n4; i += LS_ROWS)
        reinterpret_cast<float4 *>(tile)[i] = __ldg(reinterpret_cast<const float4 *>(src) + i);
      for (int e = (n4 << 2) + threadIdx.x; e < total; e += LS_ROWS) tile[e] = __ldg(src + e);
    } else {
      for (int e = threadIdx.x; e < total; e += LS_ROWS) tile[e] = __ldg(src + e);
    }
    __syncthreads();
  }
  if ((int)threadIdx.x >= rows) return;
  float *l = tile + threadIdx.x * C;
  const int64_t row = r0 + threadIdx.x;
  if (CMAX <= 16) {
    float t[CMAX], e[CMAX];
    float m = -INFINITY;
    if (direct) {
      const float2 *g2 = reinterpret_cast<const float2 *>(src + (size_t)threadIdx.x * C);
#pragma unroll
      for (int c = 0; c < CMAX; c += 2) {
        float2 u = make_float2(-INFINITY, -INFINITY);
        if (c < C) u = __ldg(g2 + (c >> 1));
        t[c] = u.x;
        t[c + 1] = u.y;
      }
    } else {
#pragma unroll
      for (int c = 0; c < CMAX; ++c) t[c] = c < C ? l[c] : -INFINITY;
    }
#pragma unroll
    for (int c = 0; c < CMAX; ++c) m = fmaxf(m, t[c]);
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < CMAX; ++c) {
      t[c] = (t[c] - m) * kLog2e;  // -inf for c >= C
      e[c] = ex2_fast(t[c]);
      s += e[c];
    }
    const float lg2s = lg2_fast(s);
    if (energy) energy[row] = fmaf(lg2s, kLn2f, m);
    const float inv_s = 1.f / s;
    if (msp) msp[row] = inv_s;  // exp(m - m) / s
    if (gen) {
      float g = 0.f;
#pragma unroll
      for (int c = 0; c < CMAX; ++c) {
        if (c < C) {
          bool take = true;
          if (M < C) {  // only the M largest probabilities under the total order (value, index)
            int greater = 0;
#pragma unroll
            for (int o = 0; o < CMAX; ++o) greater += (o < C && (t[o] > t[c] || (t[o] == t[c] && o > c))) ? 1 : 0;
            take = greater < M;
          }
          const float lp = t[c] - lg2s;  // log2 p, straight from the logits
          const float term = ex2_fast(gamma * (lp + lg2_fast(1.f - e[c] * inv_s)));
          g += take ? term : 0.f;
        }
      }
      gen[row] = -g;
    }
  } else {
    float m = -INFINITY;
    for (int c = 0; c < C; ++c) m = fmaxf(m, l[c]);
    float s = 0.f;
    for (int c = 0; c < C; ++c) {
      const float tc = (l[c] - m) * kLog2e;
      s += ex2_fast(tc);
      if (gen) l[c] = tc;  // this thread's row only
    }
    const float lg2s = lg2_fast(s);
    if (energy) energy[row] = fmaf(lg2s, kLn2f, m);
    if (msp) msp[row] = 1.f / s;
    if (gen) {
      float g = 0.f;
      for (int c = 0; c < C; ++c) {
        const float tc = l[c];
        if (M < C) {
          int greater = 0;
          for (int o = 0; o < C; ++o) greater += (l[o] > tc || (l[o] == tc && o > c)) ? 1 : 0;
          if (greater >= M) continue;
        }
        const float lp = tc - lg2s;
        g += ex2_fast(gamma * (lp + lg2_fast(1.f - ex2_fast(lp))));
      }
      gen[row] = -g;
    }
  }
}

// Any C: one warp per row.  The row is staged in shared memory when it fits (8 rows per block), else re-read from
// global memory (L1 / L2).  GEN over the M largest probabilities (funcs.py:371: np.sort(probs)[:, -M:]): the M-th
// largest logit is found by select_kth_key on order-preserving keys (softmax is monotone, so the logits order the
// probabilities), one warp-wide count per round, stopping as soon as exactly M keys lie at or above the
// candidate; elements tied with the M-th value have equal terms, so only their number matters.
// PROBS: the row already holds probabilities (generalized_entropy(probs, ...) called directly, funcs.py:347-375):
// no softmax, terms p^gamma (1 - p)^gamma with powf like NumPy's float32 power.
__device__ __forceinline__ uint32_t okey(float v) {
  const uint32_t u = __float_as_uint(v);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// k-th largest of a row's order-preserving keys, warp-cooperative.  `for_each(f)` calls f(key) for every key this
// lane holds; [kmin, kmax] are the row's smallest and largest key, n the number of keys, 1 <= k <= n.
// Returns P with count(key >= P) >= k; exact = (that count == k): the top k are exactly the keys >= P (P itself need
// not be a key).  Otherwise P IS the k-th largest key and it is tied across the cut.
// Search: an interval [lo, hi) with count(>= lo) >= k > count(>= hi) shrinks by alternating an interpolation step
// (the count is assumed linear in the key between the two ends -- activations are spread over a few binades, where
// the float bit pattern is piecewise linear in the value) with a bisection step (which bounds the rounds for any
// input).  A candidate that lands in a gap of the key set (the count equals that of an end point) is replaced by the
// nearest key on the far side of the gap, found with one more reduction: rows with an atom at the cut (post-ReLU
// zeros, quantised activations) then finish in 4-5 rounds instead of walking the gap down to one ulp.  Typical rows
// of 512 activations need 8-10 rounds where a bit-by-bit radix select needs ~22 (sign and exponent bits alone: 9).
template <class ForEach>
__device__ __forceinline__ uint32_t select_kth_key(ForEach for_each, uint32_t kmin, uint32_t kmax, int n, int k, bool &exact) {
  auto count_ge = [&](uint32_t c) {
    int m = 0;
    for_each([&](uint32_t key) { m += key >= c ? 1 : 0; });
    return __reduce_add_sync(0xffffffffu, m);
  };
  int c_hi = count_ge(kmax);  // multiplicity of the maximum
  if (c_hi >= k) {
    exact = c_hi == k;
    return kmax;
  }
  uint32_t lo = kmin, hi = kmax;
  int c_lo = n;
  for (int round = 0;; ++round) {
    if (c_lo == k) {
      exact = true;
      return lo;
    }
    const uint32_t span = hi - lo;
    if (span == 1u) {  // count(>= lo) > k > count(>= lo + 1): lo is the k-th largest key, tied
      exact = false;
      return lo;
    }
    uint32_t step = span >> 1;
    if (!(round & 1)) {
      const float frac = ((float)(c_lo - k) + 0.5f) / (float)(c_lo - c_hi);
      step = (uint32_t)__float2uint_rd(__uint2float_rz(span) * frac);
    }
    step = min(max(step, 1u), span - 1u);
    uint32_t cand = lo + step;
    const int c = count_ge(cand);
    if (c >= k) {
      if (c == c_lo) {  // no key in [lo, cand): move up to the smallest key >= cand (same count)
        uint32_t m = 0xffffffffu;
        for_each([&](uint32_t key) { m = key >= cand ? min(m, key) : m; });
        cand = __reduce_min_sync(0xffffffffu, m);
      }
      lo = cand;
      c_lo = c;
    } else if (c == c_hi) {  // no key in [cand, hi): the largest key below cand decides
      uint32_t m = 0u;
      for_each([&](uint32_t key) { m = key < cand ? max(m, key) : m; });
      const uint32_t below = __reduce_max_sync(0xffffffffu, m);  // >= lo: count(>= lo) >= k > c
      const int cb = count_ge(below);
      if (cb >= k) {  // count(>= below) >= k > count(>= below + 1): `below` is the k-th largest key
        exact = cb == k;
        return below;
      }
      hi = below;
      c_hi = cb;
    } else {
      hi = cand;
      c_hi = c;
    }
  }
}

template <bool PROBS>
__global__ void __launch_bounds__(256)
logit_scores_wide_kernel(const float *__restrict__ logits, int64_t N, int C, float gamma, int M, int stage_floats,
                         float *__restrict__ energy, float *__restrict__ msp, float *__restrict__ gen) {
  extern __shared__ __align__(16) float srow[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t row = (int64_t)blockIdx.x * 8 + warp;
  if (row >= N) return;
  const float *l = logits + row * (int64_t)C;
  if (stage_floats) {
    float *st = srow + (size_t)warp * stage_floats;
    for (int c = lane; c < C; c += 32) st[c] = __ldg(l + c);
    __syncwarp();
    l = st;
  }
  float m = 0.f, lg2s = 0.f;
  if (!PROBS) {
    m = -INFINITY;
    for (int c = lane; c < C; c += 32) m = fmaxf(m, l[c]);
    m = warp_max32(m);
    float sum = 0.f;
    for (int c = lane; c < C; c += 32) sum += ex2_fast((l[c] - m) * kLog2e);
    sum = warp_sum32(sum);
    lg2s = lg2_fast(sum);
    if (lane == 0) {
      if (energy) energy[row] = fmaf(lg2s, kLn2f, m);
      if (msp) msp[row] = 1.f / sum;
    }
  }
  if (!gen) return;
  auto term = [&](float x) -> float {
    if (PROBS) return gen_term(x, gamma);
    const float lp = (x - m) * kLog2e - lg2s;  // log2 p
    return ex2_fast(gamma * (lp + lg2_fast(1.f - ex2_fast(lp))));
  };
  float g = 0.f;
  if (M <= 0 || M >= C) {  // [:, -M:] with M >= C (or M = 0) is the whole row
    for (int c = lane; c < C; c += 32) g += term(l[c]);
  } else {
    uint32_t kmin = 0xffffffffu, kmax = 0u;
    for (int c = lane; c < C; c += 32) {
      const uint32_t k = okey(l[c]);
      kmin = min(kmin, k);
      kmax = max(kmax, k);
    }
    kmin = __reduce_min_sync(0xffffffffu, kmin);
    kmax = __reduce_max_sync(0xffffffffu, kmax);
    bool exact = false;
    const uint32_t prefix = select_kth_key(
        [&](auto f) {
          for (int c = lane; c < C; c += 32) f(okey(l[c]));
        },
        kmin, kmax, C, M, exact);
    if (exact) {
      for (int c = lane; c < C; c += 32) g += okey(l[c]) >= prefix ? term(l[c]) : 0.f;
    } else {  // prefix = key of the M-th largest value, and it is tied
      int n_gt = 0;
      float tie = -INFINITY;
      for (int c = lane; c < C; c += 32) {
        const uint32_t k = okey(l[c]);
        if (k > prefix) {
          g += term(l[c]);
          ++n_gt;
        } else if (k == prefix) {
          tie = term(l[c]);
        }
      }
      n_gt = __reduce_add_sync(0xffffffffu, n_gt);
      tie = warp_max32(tie);
      if (lane == 0) g += (float)(M - n_gt) * tie;
    }
  }
  g = warp_sum32(g);
  if (lane == 0) gen[row] = -g;
}

// ------------------------------------------------------------------------------------------
// clip -> linear -> log-sum-exp for C <= 16 classes (ReAct / DICE / DICE+ReAct heads): HBM-bound.
// One warp per row, two rows per step.  Lane l owns elements 4l + 128 i + {0..3} of a row: four
// coalesced LDG.128 bring 512 elements into registers; the weight rows come from shared memory as
// conflict-free LDS.128, shared by the two rows.  The 16 per-lane partial dot products are reduced
// across the warp with a halving butterfly (8 + 4 + 2 + 1 + 1 = 16 shuffles instead of 16 x 5): after
// it lane l holds the full dot product of class (l >> 1) & 15, and the log-sum-exp is one more
// warp reduction.
// ------------------------------------------------------------------------------------------
constexpr int LH_C = 16;

__device__ __forceinline__ float butterfly16(float (&p)[LH_C], int lane) {
  // step off = 16: lanes with bit 4 clear keep classes 0..7, the others 8..15
  float q8[8];
  {
    const bool up = lane & 16;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float send = up ? p[i] : p[i + 8];
      const float keep = up ? p[i + 8] : p[i];
      q8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
  }
  float q4[4];
  {
    const bool up = lane & 8;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float send = up ? q8[i] : q8[i + 4];
      const float keep = up ? q8[i + 4] : q8[i];
      q4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
  }
  float q2[2];
  {
    const bool up = lane & 4;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const float send = up ? q4[i] : q4[i + 2];
      const float keep = up ? q4[i + 2] : q4[i];
      q2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
  }
  float q1;
  {
    const bool up = lane & 2;
    const float send = up ? q2[0] : q2[1];
    const float keep = up ? q2[1] : q2[0];
    q1 = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  }
  return q1 + __shfl_xor_sync(0xffffffffu, q1, 1);  // class ((lane>>4)&1)*8 + ((lane>>3)&1)*4 + ((lane>>2)&1)*2 + ((lane>>1)&1)
}

// CN = number of classes rounded up to a multiple of 4 (weight rows >= C are zero): no guards in the
// inner loops.  The next pair of rows is fetched into registers while the current one is reduced.
template <int CN>
__global__ void __launch_bounds__(256, 2)
linear_lse_c16_kernel(const float *__restrict__ X, int64_t N, int d, const float *__restrict__ W,
                      const float *__restrict__ b, int C, float clip, float *__restrict__ out) {
  extern __shared__ __align__(16) float sW[];  // [CN][dpad], rows >= C and columns >= d are zero
  const int dpad = (d + 511) & ~511;
  for (int e = threadIdx.x; e < CN * dpad; e += blockDim.x) {
    const int c = e / dpad, j = e - c * dpad;
    sW[e] = (c < C && j < d) ? __ldg(W + (size_t)c * d + j) : 0.f;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int my_class = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
  const float my_bias = my_class < C ? __ldg(b + my_class) : 0.f;
  const bool vec = (d % 4 == 0) && ((reinterpret_cast<uintptr_t>(X) & 15) == 0);
  const int64_t pairs = (N + 1) >> 1;
  const int64_t wstride = (int64_t)gridDim.x * 8;
  const int nchunk = dpad >> 9;
  auto load_chunk = [&](int64_t pr, int ch, float4 (&a0)[4], float4 (&a1)[4]) {
    const int64_t row0 = 2 * pr, row1 = (2 * pr + 1 < N) ? 2 * pr + 1 : 2 * pr;
    const float *x0 = X + row0 * (int64_t)d, *x1 = X + row1 * (int64_t)d;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int j = ch * 512 + 128 * i + 4 * lane;
      float4 u = make_float4(0.f, 0.f, 0.f, 0.f), v = u;
      if (pr < pairs) {
        if (vec && j + 3 < d) {
          u = __ldg(reinterpret_cast<const float4 *>(x0 + j));
          v = __ldg(reinterpret_cast<const float4 *>(x1 + j));
        } else {
          if (j + 0 < d) { u.x = __ldg(x0 + j + 0); v.x = __ldg(x1 + j + 0); }
          if (j + 1 < d) { u.y = __ldg(x0 + j + 1); v.y = __ldg(x1 + j + 1); }
          if (j + 2 < d) { u.z = __ldg(x0 + j + 2); v.z = __ldg(x1 + j + 2); }
          if (j + 3 < d) { u.w = __ldg(x0 + j + 3); v.w = __ldg(x1 + j + 3); }
        }
      }
      a0[i] = u;
      a1[i] = v;
    }
  };
  float4 n0[4], n1[4];  // prefetched chunk
  int64_t pr = (int64_t)blockIdx.x * 8 + warp;
  load_chunk(pr, 0, n0, n1);
  for (; pr < pairs; pr += wstride) {
    float p0[LH_C], p1[LH_C];
#pragma unroll
    for (int c = 0; c < LH_C; ++c) p0[c] = p1[c] = 0.f;
    for (int ch = 0; ch < nchunk; ++ch) {
      float4 a0[4], a1[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        a0[i] = make_float4(fminf(n0[i].x, clip), fminf(n0[i].y, clip), fminf(n0[i].z, clip), fminf(n0[i].w, clip));
        a1[i] = make_float4(fminf(n1[i].x, clip), fminf(n1[i].y, clip), fminf(n1[i].z, clip), fminf(n1[i].w, clip));
      }
      if (ch + 1 < nchunk)
        load_chunk(pr, ch + 1, n0, n1);
      else
        load_chunk(pr + wstride, 0, n0, n1);
      const float *wbase = sW + ch * 512 + 4 * lane;
#pragma unroll
      for (int c = 0; c < CN; ++c) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 w = *reinterpret_cast<const float4 *>(wbase + c * dpad + 128 * i);
          p0[c] = fmaf(a0[i].x, w.x, fmaf(a0[i].y, w.y, fmaf(a0[i].z, w.z, fmaf(a0[i].w, w.w, p0[c]))));
          p1[c] = fmaf(a1[i].x, w.x, fmaf(a1[i].y, w.y, fmaf(a1[i].z, w.z, fmaf(a1[i].w, w.w, p1[c]))));
        }
      }
    }
    const float d0 = butterfly16(p0, lane), d1 = butterfly16(p1, lane);  // all lanes take part
    const float l0 = my_class < C ? d0 + my_bias : -INFINITY;
    const float l1 = my_class < C ? d1 + my_bias : -INFINITY;
    const float m0 = warp_max32(l0), m1 = warp_max32(l1);
    // every class is held by two lanes (l, l ^ 1): halve the sum
    const float s0 = 0.5f * warp_sum32(exp2f((l0 - m0) * kLog2e));
    const float s1 = 0.5f * warp_sum32(exp2f((l1 - m1) * kLog2e));
    if (lane == 0) {
      out[2 * pr] = fmaf(__log2f(s0), kLn2f, m0);
      if (2 * pr + 1 < N) out[2 * pr + 1] = fmaf(__log2f(s1), kLn2f, m1);
    }
  }
}

// ------------------------------------------------------------------------------------------
// ASH-S head for C <= 16, d <= 1024: keep the k largest activations of the row, scale by
// exp(sum_all / sum_kept), linear layer, log-sum-exp (funcs.py:230-261 + postprocessors.py:1212-1220).
// The row lives in registers (lane l owns elements 4l + 128 i + q); the k-th largest value is found by
// an interpolation / bisection search on order-preserving integer keys (select_kth_key: one warp-wide REDUX per
// round, 5-8 rounds for distinct activations), which stops as soon as exactly k keys lie at or above the
// candidate; only when the k-th value is tied does it narrow down to that key and rank the ties by index.
// ------------------------------------------------------------------------------------------
template <int CN, int NCH>
__global__ void __launch_bounds__(256, (NCH == 1 ? 2 : 1))  // two chunks: row + keys + prefetched row need > 128 registers
ash_lse_c16_kernel(const float *__restrict__ X, int64_t N, int d, const float *__restrict__ W,
                   const float *__restrict__ b, int C, int k_keep, float *__restrict__ out) {
  constexpr int NV = NCH * 16;
  extern __shared__ __align__(16) float sW[];  // [CN][dpad], rows >= C and columns >= d are zero
  constexpr int dpad = NCH * 512;
  for (int e = threadIdx.x; e < CN * dpad; e += blockDim.x) {
    const int c = e / dpad, j = e - c * dpad;
    sW[e] = (c < C && j < d) ? __ldg(W + (size_t)c * d + j) : 0.f;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int my_class = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
  const float my_bias = my_class < C ? __ldg(b + my_class) : 0.f;
  const bool vec = (d % 4 == 0) && ((reinterpret_cast<uintptr_t>(X) & 15) == 0);
  const int64_t wstride = (int64_t)gridDim.x * 8;
  // the next row of this warp is in flight while the current one is selected and scored: the selection is a chain of
  // dependent warp reductions, and without the prefetch the SM holds too few bytes in flight to cover HBM latency
  auto load_row = [&](int64_t row, float4 (&u)[NV / 4]) {
    const float *x = X + row * (int64_t)d;
#pragma unroll
    for (int g = 0; g < NV / 4; ++g) {  // g = 4 ch + i: elements 128 g + 4 lane + q
      const int j = 128 * g + 4 * lane;
      u[g] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row < N) {
        if (vec && j + 3 < d) {
          u[g] = __ldg(reinterpret_cast<const float4 *>(x + j));
        } else {
          if (j + 0 < d) u[g].x = __ldg(x + j + 0);
          if (j + 1 < d) u[g].y = __ldg(x + j + 1);
          if (j + 2 < d) u[g].z = __ldg(x + j + 2);
          if (j + 3 < d) u[g].w = __ldg(x + j + 3);
        }
      }
    }
  };
  float4 nx[NV / 4];
  load_row((int64_t)blockIdx.x * 8 + warp, nx);
  for (int64_t row = (int64_t)blockIdx.x * 8 + warp; row < N; row += wstride) {
    float v[NV];
    uint32_t key[NV];
#pragma unroll
    for (int g = 0; g < NV / 4; ++g) {
      const int j = 128 * g + 4 * lane;
      v[4 * g + 0] = nx[g].x; v[4 * g + 1] = nx[g].y; v[4 * g + 2] = nx[g].z; v[4 * g + 3] = nx[g].w;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint32_t bits = __float_as_uint(v[4 * g + q]);
        const uint32_t k = (bits & 0x80000000u) ? ~bits : (bits | 0x80000000u);
        key[4 * g + q] = (j + q < d) ? k : 0u;  // padding can never be selected (real keys are > 0)
      }
    }
    load_row(row + wstride, nx);
    float s1 = 0.f;
#pragma unroll
    for (int e = 0; e < NV; ++e) s1 += v[e];
    s1 = warp_sum32(s1);
    // the k-th largest key (padding keys are 0: below every real key, never counted for a candidate >= kmin > 0)
    uint32_t kmin = 0xffffffffu, kmax = 0u;
#pragma unroll
    for (int e = 0; e < NV; ++e) {
      kmin = min(kmin, key[e] ? key[e] : 0xffffffffu);
      kmax = max(kmax, key[e]);
    }
    kmin = __reduce_min_sync(0xffffffffu, kmin);
    kmax = __reduce_max_sync(0xffffffffu, kmax);
    bool exact = false;
    const uint32_t prefix = select_kth_key(
        [&](auto f) {
#pragma unroll
          for (int e = 0; e < NV; ++e) f(key[e]);
        },
        kmin, kmax, d, k_keep, exact);
    float s2 = 0.f;
    if (exact) {
#pragma unroll
      for (int e = 0; e < NV; ++e) {
        v[e] = key[e] >= prefix ? v[e] : 0.f;
        s2 += v[e];
      }
    } else {
      // prefix is the k-th largest key and it is tied: keep everything above it and the ties with the
      // lowest indices (index order: block of 128, then lane, then q)
      int n_gt = 0;
#pragma unroll
      for (int e = 0; e < NV; ++e) n_gt += (key[e] > prefix) ? 1 : 0;
      n_gt = __reduce_add_sync(0xffffffffu, n_gt);
      int ties_left = k_keep - n_gt;
#pragma unroll
      for (int g = 0; g < NV / 4; ++g) {
        int mine = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) mine += (key[4 * g + q] == prefix) ? 1 : 0;
        int incl = mine;  // inclusive scan over lanes
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
          const int t = __shfl_up_sync(0xffffffffu, incl, off);
          if (lane >= off) incl += t;
        }
        int rank = incl - mine;  // ties of this block in lower lanes
        const int total = __shfl_sync(0xffffffffu, incl, 31);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int e = 4 * g + q;
          const bool tie = key[e] == prefix;
          const bool keep = key[e] > prefix || (tie && rank < ties_left);
          rank += tie ? 1 : 0;
          v[e] = keep ? v[e] : 0.f;
          s2 += v[e];
        }
        ties_left -= total;  // may go negative: no more ties are kept
      }
    }
    s2 = warp_sum32(s2);
    const float scale = expf(s1 / s2);
    float p[LH_C];
#pragma unroll
    for (int c = 0; c < LH_C; ++c) p[c] = 0.f;
#pragma unroll
    for (int c = 0; c < CN; ++c) {
#pragma unroll
      for (int g = 0; g < NV / 4; ++g) {
        const float4 w = *reinterpret_cast<const float4 *>(sW + c * dpad + 128 * g + 4 * lane);
        p[c] = fmaf(v[4 * g], w.x, fmaf(v[4 * g + 1], w.y, fmaf(v[4 * g + 2], w.z, fmaf(v[4 * g + 3], w.w, p[c]))));
      }
    }
    const float dot = butterfly16(p, lane);
    const float lg = my_class < C ? fmaf(scale, dot, my_bias) : -INFINITY;
    const float m = warp_max32(lg);
    const float sm = 0.5f * warp_sum32(exp2f((lg - m) * kLog2e));  // every class sits in two lanes
    if (lane == 0) out[row] = fmaf(__log2f(sm), kLn2f, m);
  }
}

template <int CN>
static int launch_ash16(int nch, unsigned blocks, size_t smem, cudaStream_t st, const float *X, int64_t N, int d,
                        const float *W, const float *b, int C, int k_keep, float *out) {
  static PerDeviceFlag attr;
  if (!attr) {
    RUNIA_CUDA(cudaFuncSetAttribute(ash_lse_c16_kernel<CN, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    RUNIA_CUDA(cudaFuncSetAttribute(ash_lse_c16_kernel<CN, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    attr = true;
  }
  if (nch == 1)
    ash_lse_c16_kernel<CN, 1><<<blocks, 256, smem, st>>>(X, N, d, W, b, C, k_keep, out);
  else
    ash_lse_c16_kernel<CN, 2><<<blocks, 256, smem, st>>>(X, N, d, W, b, C, k_keep, out);
  return RUNIA_OK;
}

// ------------------------------------------------------------------------------------------
// clip -> linear -> log-sum-exp, one warp per row, W (C x d) staged in shared memory
// ------------------------------------------------------------------------------------------
constexpr int CL_MAXC = 64;

template <bool ASH>
__global__ void __launch_bounds__(256)
linear_lse_kernel(const float *__restrict__ X, int64_t N, int d, const float *__restrict__ W,
                  const float *__restrict__ b, int C, float clip, int k_keep, float *__restrict__ out) {
  extern __shared__ float sW[];  // [C][d]
  for (int e = threadIdx.x; e < C * d; e += blockDim.x) sW[e] = __ldg(W + e);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t wstride = (int64_t)gridDim.x * 8;
  for (int64_t row = (int64_t)blockIdx.x * 8 + warp; row < N; row += wstride) {
    const float *x = X + row * (int64_t)d;
    float scale = 1.f;
    uint32_t tkey = 0;   // ASH: order-preserving key of the k-th largest activation
    int n_ties_keep = 0; // ASH: how many elements equal to the threshold are kept (lowest indices)
    if (ASH) {
      // k-th largest order-preserving integer key
      auto keyof = [](float v) -> uint32_t {
        const uint32_t u = __float_as_uint(v);
        return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
      };
      uint32_t kmin = 0xffffffffu, kmax = 0u;
      for (int j = lane; j < d; j += 32) {
        const uint32_t k = keyof(__ldg(x + j));
        kmin = min(kmin, k);
        kmax = max(kmax, k);
      }
      kmin = __reduce_min_sync(0xffffffffu, kmin);
      kmax = __reduce_max_sync(0xffffffffu, kmax);
      bool exact = false;  // exact: the top k are the keys >= tkey (tkey need not be a key: then no element ties with it)
      tkey = select_kth_key(
          [&](auto f) {
            for (int j = lane; j < d; j += 32) f(keyof(__ldg(x + j)));
          },
          kmin, kmax, d, k_keep, exact);
      int n_gt = 0;
      float s1 = 0.f, s_gt = 0.f;
      for (int j = lane; j < d; j += 32) {
        const float v = __ldg(x + j);
        s1 += v;
        if (keyof(v) > tkey) {
          ++n_gt;
          s_gt += v;
        }
      }
      n_gt = __reduce_add_sync(0xffffffffu, n_gt);
      s1 = warp_sum32(s1);
      s_gt = warp_sum32(s_gt);
      n_ties_keep = k_keep - n_gt;
      uint32_t tb = tkey;
      const float tval = __uint_as_float((tb & 0x80000000u) ? (tb & 0x7fffffffu) : ~tb);
      const float s2 = s_gt + (n_ties_keep > 0 ? (float)n_ties_keep * tval : 0.f);
      scale = expf(s1 / s2);
    }
    float mx = -INFINITY, sm = 0.f;
    // ASH tie handling: element j equal to the threshold is kept iff fewer than n_ties_keep
    // equal elements precede it (lowest indices win).
    for (int c = 0; c < C; ++c) {
      const float *w = sW + c * d;
      float p = 0.f;
      if (!ASH) {
        for (int j = lane; j < d; j += 32) {
          float v = __ldg(x + j);
          v = v > clip ? clip : v;
          p = fmaf(v, w[j], p);
        }
      } else {
        auto keyof = [](float v) -> uint32_t {
          const uint32_t u = __float_as_uint(v);
          return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
        };
        int ties_before = 0;
        for (int j0 = 0; j0 < d; j0 += 32) {
          const int j = j0 + lane;
          const float v = j < d ? __ldg(x + j) : 0.f;
          const uint32_t kk = j < d ? keyof(v) : 0u;
          const bool is_tie = (j < d) && (kk == tkey);
          const unsigned tie_mask = __ballot_sync(0xffffffffu, is_tie);
          const int my_tie_rank = ties_before + __popc(tie_mask & ((1u << lane) - 1u));
          const bool keep = (j < d) && (kk > tkey || (is_tie && my_tie_rank < n_ties_keep));
          if (keep) p = fmaf(v, w[j], p);
          ties_before += __popc(tie_mask);
        }
      }
      p = warp_sum32(p);
      const float lg = fmaf(scale, p, __ldg(b + c));
      const float m_new = fmaxf(mx, lg);
      sm = sm * expf(mx - m_new) + expf(lg - m_new);
      mx = m_new;
    }
    if (lane == 0) out[row] = mx + logf(sm);
  }
}

// ------------------------------------------------------------------------------------------
// clip -> linear -> log-sum-exp for ANY head (C, d): one warp per four rows, W and the rows streamed from
// L1 / L2, online log-sum-exp over the classes.  The shapes the faster kernels cannot take end up here
// (d % 4 != 0 or d > 4096 with a head that does not fit shared memory).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
linear_lse_rows4_kernel(const float *__restrict__ X, int64_t N, int d, const float *__restrict__ W,
                        const float *__restrict__ b, int C, float clip, float *__restrict__ out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t groups = (N + 3) >> 2;
  for (int64_t gq = (int64_t)blockIdx.x * 8 + warp; gq < groups; gq += (int64_t)gridDim.x * 8) {
    const float *x[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) x[r] = X + (4 * gq + r < N ? 4 * gq + r : 4 * gq) * (int64_t)d;
    float mx[4], sm[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      mx[r] = -INFINITY;
      sm[r] = 0.f;
    }
    for (int c = 0; c < C; ++c) {
      const float *w = W + (size_t)c * d;
      float p[4] = {0.f, 0.f, 0.f, 0.f};
      for (int j = lane; j < d; j += 32) {
        const float wj = __ldg(w + j);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          float v = __ldg(x[r] + j);
          v = v > clip ? clip : v;  // NaN stays NaN like numpy.clip
          p[r] = fmaf(v, wj, p[r]);
        }
      }
      const float bc = __ldg(b + c);
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float lg = warp_sum32(p[r]) + bc;
        const float m_new = fmaxf(mx[r], lg);
        sm[r] = sm[r] * expf(mx[r] - m_new) + expf(lg - m_new);
        mx[r] = m_new;
      }
    }
    if (lane < 4 && 4 * gq + lane < N) {
      float mo = mx[0], so = sm[0];
#pragma unroll
      for (int r = 1; r < 4; ++r)
        if (lane == r) {
          mo = mx[r];
          so = sm[r];
        }
      out[4 * gq + lane] = mo + logf(so);
    }
  }
}

// ------------------------------------------------------------------------------------------
// ASH-S pruning for any row width (funcs.py:230-261): keep the k largest activations of the row (ties with the
// k-th value: lowest indices first), zero the rest, scale by exp(sum_all / sum_kept).  One warp per row; the
// k-th largest value by select_kth_key on order-preserving keys.  Feeds the general heads above (C > 16 or
// d > 1024); the small heads fuse the same selection into the head kernel (ash_lse_c16_kernel).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
ash_prune_kernel(const float *__restrict__ X, int64_t N, int d, int k_keep, float *__restrict__ out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int64_t row = (int64_t)blockIdx.x * 8 + warp; row < N; row += (int64_t)gridDim.x * 8) {
    const float *x = X + row * (int64_t)d;
    uint32_t kmin = 0xffffffffu, kmax = 0u;
    for (int j = lane; j < d; j += 32) {
      const uint32_t k = okey(__ldg(x + j));
      kmin = min(kmin, k);
      kmax = max(kmax, k);
    }
    kmin = __reduce_min_sync(0xffffffffu, kmin);
    kmax = __reduce_max_sync(0xffffffffu, kmax);
    bool exact = false;
    const uint32_t prefix = select_kth_key(
        [&](auto f) {
          for (int j = lane; j < d; j += 32) f(okey(__ldg(x + j)));
        },
        kmin, kmax, d, k_keep, exact);  // exact or tied, the code below only needs count(>= prefix) >= k
    int n_gt = 0;
    float s1 = 0.f, s_gt = 0.f;
    for (int j = lane; j < d; j += 32) {
      const float v = __ldg(x + j);
      s1 += v;
      if (okey(v) > prefix) {
        ++n_gt;
        s_gt += v;
      }
    }
    n_gt = __reduce_add_sync(0xffffffffu, n_gt);
    s1 = warp_sum32(s1);
    s_gt = warp_sum32(s_gt);
    // after an early break `prefix` need not be a key of the row: then every kept element compares greater, and
    // the number of ties to keep is k - n_gt = (elements >= prefix) - n_gt = elements == prefix, possibly 0
    const int n_ties_keep = k_keep - n_gt;
    // the tied value itself is the float behind `prefix` (all ties are the same float)
    const float tv = (prefix & 0x80000000u) ? __uint_as_float(prefix & 0x7fffffffu) : __uint_as_float(~prefix);
    const float s2 = s_gt + (n_ties_keep > 0 ? (float)n_ties_keep * tv : 0.f);
    const float scale = expf(s1 / s2);
    float *o = out + row * (int64_t)d;
    int ties_before = 0;
    for (int j0 = 0; j0 < d; j0 += 32) {
      const int j = j0 + lane;
      const float v = j < d ? __ldg(x + j) : 0.f;
      const uint32_t k = j < d ? okey(v) : 0u;
      const bool is_tie = j < d && k == prefix;
      const unsigned tie_mask = __ballot_sync(0xffffffffu, is_tie);
      const int my_rank = ties_before + __popc(tie_mask & ((1u << lane) - 1u));
      const bool keep = j < d && (k > prefix || (is_tie && my_rank < n_ties_keep));
      if (j < d) o[j] = keep ? v * scale : 0.f;
      ties_before += __popc(tie_mask);
    }
  }
}

}  // namespace runia

using namespace runia;

template <bool PROBS>
static cudaError_t launch_logit_wide(const float *logits, int64_t N, int C, float gamma, int M, float *energy, float *msp,
                                     float *gen, cudaStream_t st) {
  static PerDeviceFlag attr;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(logit_scores_wide_kernel<PROBS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    if (e != cudaSuccess) return e;
    attr = true;
  }
  const int stage = (C + 3) & ~3;  // floats per warp; 8 warps per block
  const size_t smem = (size_t)8 * stage * sizeof(float);
  const bool fits = smem <= 160 * 1024;
  logit_scores_wide_kernel<PROBS><<<(unsigned)ceil_div(N, 8), 256, fits ? smem : 0, st>>>(logits, N, C, gamma, M,
                                                                                        fits ? stage : 0, energy, msp, gen);
  return cudaSuccess;
}

extern "C" int runia_gen_entropy_f32(const float *probs, int64_t N, int C, float gamma, int M, float *out, void *stream) {
  RUNIA_NVTX();
  RUNIA_REQUIRE(N >= 0 && C > 0, RUNIA_E_BADARG, "gen_entropy: bad sizes");
  if (N == 0) return RUNIA_OK;
  RUNIA_REQUIRE(probs && out, RUNIA_E_BADARG, "gen_entropy: null pointer");
  RUNIA_CUDA(launch_logit_wide<true>(probs, N, C, gamma, M, nullptr, nullptr, out, (cudaStream_t)stream));
  count_launch();
  return finish_launch("gen_entropy");
}

extern "C" int runia_logit_scores_f32(const float *logits, int64_t N, int C, float gamma, int M, float *energy,
                                      float *msp, float *gen, void *stream) {
  RUNIA_NVTX();
  RUNIA_REQUIRE(N >= 0 && C > 0, RUNIA_E_BADARG, "logit_scores: bad sizes");
  if (N == 0) return RUNIA_OK;
  RUNIA_REQUIRE(logits && (energy || msp || gen), RUNIA_E_BADARG, "logit_scores: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if (C <= 64) {
    const size_t smem = (size_t)LS_ROWS * C * sizeof(float);
    static PerDeviceFlag attr;
    if (!attr) {
      RUNIA_CUDA(cudaFuncSetAttribute(logit_scores_small_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 66 * 1024));
      attr = true;
    }
    const unsigned grid = (unsigned)ceil_div(N, LS_ROWS);
    if (C <= 16) {
      // the register row is sized to the class count rounded up to even: no dead (-inf) columns in the unrolled body
      switch ((C + 1) & ~1) {
#define RUNIA_LS_CASE(CM)                                                                                   \
  case CM:                                                                                                  \
    logit_scores_small_kernel<CM><<<grid, LS_ROWS, smem, st>>>(logits, N, C, gamma, M, energy, msp, gen); \
    break;
        RUNIA_LS_CASE(2) RUNIA_LS_CASE(4) RUNIA_LS_CASE(6) RUNIA_LS_CASE(8) RUNIA_LS_CASE(10) RUNIA_LS_CASE(12)
        RUNIA_LS_CASE(14) RUNIA_LS_CASE(16)
#undef RUNIA_LS_CASE
      }
    } else {
      logit_scores_small_kernel<64><<<grid, LS_ROWS, smem, st>>>(logits, N, C, gamma, M, energy, msp, gen);
    }
  } else {
    RUNIA_CUDA(launch_logit_wide<false>(logits, N, C, gamma, M, energy, msp, gen, st));
  }
  count_launch();
  return finish_launch("logit_scores");
}

static int launch_linear_lse(bool ash, const float *X, int64_t N, int d, const float *W, const float *b, int C,
                             float clip, int k_keep, float *out, void *stream) {
  RUNIA_REQUIRE(N >= 0 && d > 0 && C > 0, RUNIA_E_BADARG, "linear_lse: bad sizes");
  if (N == 0) return RUNIA_OK;
  RUNIA_REQUIRE(X && W && b && out, RUNIA_E_BADARG, "linear_lse: null pointer");
  const bool fits = C <= CL_MAXC && (size_t)C * d * 4 <= 200 * 1024;  // head resident in shared memory
  if (!fits) {
    RUNIA_REQUIRE(!ash, RUNIA_E_UNSUPPORTED,
                  "ash_linear_lse: C=%d, d=%d exceed the fused kernel (C <= %d, C*d*4 <= 200 KiB): prune with "
                  "runia_ash_prune_f32, then call runia_clip_linear_lse_*", C, d, CL_MAXC);
    const unsigned blocks = (unsigned)std::min<int64_t>(ceil_div(ceil_div(N, 4), 8), (int64_t)kNumSMs * 8);
    linear_lse_rows4_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(X, N, d, W, b, C, clip, out);
    count_launch();
    return finish_launch("linear_lse(general)");
  }
  const size_t smem = (size_t)C * d * sizeof(float);
  static PerDeviceFlag attr;
  if (!attr) {
    RUNIA_CUDA(cudaFuncSetAttribute(linear_lse_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    RUNIA_CUDA(cudaFuncSetAttribute(linear_lse_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr = true;
  }
  const int cn = (C + 3) & ~3;
  const size_t smem16 = (size_t)cn * ((d + 511) & ~511) * sizeof(float);
  if (!ash && C <= LH_C && smem16 <= 100 * 1024) {
    static PerDeviceFlag attr16;
    if (!attr16) {
      RUNIA_CUDA(cudaFuncSetAttribute(linear_lse_c16_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
      RUNIA_CUDA(cudaFuncSetAttribute(linear_lse_c16_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
      RUNIA_CUDA(cudaFuncSetAttribute(linear_lse_c16_kernel<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
      RUNIA_CUDA(cudaFuncSetAttribute(linear_lse_c16_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
      attr16 = true;
    }
    const unsigned blocks16 = (unsigned)std::min<int64_t>(ceil_div((N + 1) / 2, 8), (int64_t)kNumSMs * 2);
    cudaStream_t st = (cudaStream_t)stream;
    if (cn == 4) linear_lse_c16_kernel<4><<<blocks16, 256, smem16, st>>>(X, N, d, W, b, C, clip, out);
    else if (cn == 8) linear_lse_c16_kernel<8><<<blocks16, 256, smem16, st>>>(X, N, d, W, b, C, clip, out);
    else if (cn == 12) linear_lse_c16_kernel<12><<<blocks16, 256, smem16, st>>>(X, N, d, W, b, C, clip, out);
    else linear_lse_c16_kernel<16><<<blocks16, 256, smem16, st>>>(X, N, d, W, b, C, clip, out);
    count_launch();
    return finish_launch("linear_lse(c16)");
  }
  if (ash && C <= LH_C && d <= 1024) {
    const int nch = d <= 512 ? 1 : 2;
    const size_t smem_a = (size_t)cn * nch * 512 * sizeof(float);
    const unsigned blocks_a = (unsigned)std::min<int64_t>(ceil_div(N, 8), (int64_t)kNumSMs * (nch == 1 ? 2 : 1));
    cudaStream_t st = (cudaStream_t)stream;
    int rc;
    if (cn == 4) rc = launch_ash16<4>(nch, blocks_a, smem_a, st, X, N, d, W, b, C, k_keep, out);
    else if (cn == 8) rc = launch_ash16<8>(nch, blocks_a, smem_a, st, X, N, d, W, b, C, k_keep, out);
    else if (cn == 12) rc = launch_ash16<12>(nch, blocks_a, smem_a, st, X, N, d, W, b, C, k_keep, out);
    else rc = launch_ash16<16>(nch, blocks_a, smem_a, st, X, N, d, W, b, C, k_keep, out);
    if (rc) return rc;
    count_launch();
    return finish_launch("ash_lse(c16)");
  }
  int64_t blocks = ceil_div(N, 8);
  const int64_t cap = (int64_t)kNumSMs * 8;
  if (blocks > cap) blocks = cap;
  if (ash)
    linear_lse_kernel<true><<<(unsigned)blocks, 256, smem, (cudaStream_t)stream>>>(X, N, d, W, b, C, clip, k_keep, out);
  else
    linear_lse_kernel<false><<<(unsigned)blocks, 256, smem, (cudaStream_t)stream>>>(X, N, d, W, b, C, clip, k_keep, out);
  count_launch();
  return finish_launch("linear_lse");
}

extern "C" int runia_clip_linear_lse_f32(const float *X, int64_t N, int d, const float *W, const float *b, int C,
                                         float clip, float *out, void *stream) {
  RUNIA_NVTX();
  return launch_linear_lse(false, X, N, d, W, b, C, clip, 0, out, stream);
}

extern "C" int runia_ash_prune_f32(const float *X, int64_t N, int d, int k_keep, float *out, void *stream) {
  RUNIA_NVTX();
  RUNIA_REQUIRE(N >= 0 && d > 0, RUNIA_E_BADARG, "ash_prune: bad sizes");
  RUNIA_REQUIRE(k_keep >= 1 && k_keep <= d, RUNIA_E_BADARG, "ash_prune: k_keep=%d outside [1, d=%d]", k_keep, d);
  if (N == 0) return RUNIA_OK;
  RUNIA_REQUIRE(X && out, RUNIA_E_BADARG, "ash_prune: null pointer");
  const unsigned blocks = (unsigned)std::min<int64_t>(ceil_div(N, 8), (int64_t)kNumSMs * 8);
  ash_prune_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(X, N, d, k_keep, out);
  count_launch();
  return finish_launch("ash_prune");
}

extern "C" int runia_ash_linear_lse_f32(const float *X, int64_t N, int d, const float *W, const float *b, int C,
                                        int k_keep, float *out, void *stream) {
  RUNIA_NVTX();
  RUNIA_REQUIRE(k_keep >= 1 && k_keep <= d, RUNIA_E_BADARG, "ash_linear_lse: k_keep=%d outside [1, d=%d]", k_keep, d);
  return launch_linear_lse(true, X, N, d, W, b, C, INFINITY, k_keep, out, stream);
}

// ------------------------------------------------------------------------------------------
// (f4) predictive entropy and mutual information of MC-dropout logits -- inference/funcs.py:430-465
// (`get_predictive_uncertainty_score`): logits [N * n_mc, C] item-major; p = softmax per row;
//   pred_h = -sum_c mean_s(p) log mean_s(p);   mi = pred_h - mean_s( -sum_c p log p ).
// One warp per item; a group of CP = 2^ceil(log2 C) <= 32 lanes owns one MC sample at a time (32 / CP samples
// per pass), or the whole warp strides over the classes when C > 32.  Like upstream, a probability that
// underflows to exactly 0 gives NaN (0 * log 0).
// ------------------------------------------------------------------------------------------
namespace runia {

__global__ void __launch_bounds__(256)
pred_uncertainty_kernel(const float *__restrict__ logits, int64_t n_items, int n_mc, int C, float *__restrict__ pred_h,
                        float *__restrict__ mi) {
  const int lane = threadIdx.x & 31;
  const int64_t item = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (item >= n_items) return;
  const float *base = logits + item * (int64_t)n_mc * C;
  if (C <= 32) {
    int cp = 1;
    while (cp < C) cp <<= 1;
    const int groups = 32 / cp, g = lane / cp, c = lane % cp;
    const bool act = c < C;
    float mean_p = 0.f, ent = 0.f;  // this lane: running sum over its samples of p_c and of -sum_c p log p (group-reduced)
    for (int s0 = 0; s0 < n_mc; s0 += groups) {
      const int s = s0 + g;
      const bool ok = act && s < n_mc;
      const float l = ok ? __ldg(base + (int64_t)s * C + c) : -INFINITY;
      float m = l;
      for (int off = cp >> 1; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
      const float e = ok ? expf(l - m) : 0.f;
      float sum = e;
      for (int off = cp >> 1; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
      const float p = ok ? e / sum : 0.f;
      float t = ok ? p * logf(p) : 0.f;  // NaN when p == 0, like torch
      for (int off = cp >> 1; off > 0; off >>= 1) t += __shfl_xor_sync(0xffffffffu, t, off);
      mean_p += p;
      if (s < n_mc) ent -= t;
    }
    // combine the groups: lanes with the same class c
    for (int off = cp; off < 32; off <<= 1) {
      mean_p += __shfl_xor_sync(0xffffffffu, mean_p, off);
      ent += __shfl_xor_sync(0xffffffffu, ent, off);
    }
    mean_p /= (float)n_mc;
    float h = act ? mean_p * logf(mean_p) : 0.f;
    for (int off = cp >> 1; off > 0; off >>= 1) h += __shfl_xor_sync(0xffffffffu, h, off);
    if (lane == 0) {
      const float ph = -h;
      if (pred_h) pred_h[item] = ph;
      if (mi) mi[item] = ph - ent / (float)n_mc;
    }
  } else {
    // wide rows: the mean probabilities of the item are accumulated in shared memory (8 warps x C floats)
    extern __shared__ float acc[];
    float *mp = acc + (size_t)(threadIdx.x >> 5) * C;
    for (int c = lane; c < C; c += 32) mp[c] = 0.f;
    float ent = 0.f;
    for (int s = 0; s < n_mc; ++s) {
      const float *l = base + (int64_t)s * C;
      float m = -INFINITY;
      for (int c = lane; c < C; c += 32) m = fmaxf(m, __ldg(l + c));
      m = warp_max32(m);
      float sum = 0.f;
      for (int c = lane; c < C; c += 32) sum += expf(__ldg(l + c) - m);
      sum = warp_sum32(sum);
      float t = 0.f;
      for (int c = lane; c < C; c += 32) {
        const float p = expf(__ldg(l + c) - m) / sum;
        t += p * logf(p);
        mp[c] += p;
      }
      ent -= warp_sum32(t);
    }
    float h = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float q = mp[c] / (float)n_mc;
      h += q * logf(q);
    }
    h = warp_sum32(h);
    if (lane == 0) {
      if (pred_h) pred_h[item] = -h;
      if (mi) mi[item] = -h - ent / (float)n_mc;
    }
  }
}

}  // namespace runia

extern "C" int runia_pred_uncertainty_f32(const float *logits, int64_t n_items, int n_mc, int C, float *pred_h, float *mi,
                                          void *stream) {
  RUNIA_NVTX();
  RUNIA_REQUIRE(n_items >= 0 && n_mc >= 1 && C >= 1, RUNIA_E_BADARG, "pred_uncertainty: bad sizes");
  RUNIA_REQUIRE(C <= 4096, RUNIA_E_UNSUPPORTED, "pred_uncertainty: C=%d > 4096", C);
  if (n_items == 0) return RUNIA_OK;
  RUNIA_REQUIRE(logits && (pred_h || mi), RUNIA_E_BADARG, "pred_uncertainty: null pointer");
  const size_t smem = C > 32 ? (size_t)8 * C * sizeof(float) : 0;
  static PerDeviceFlag attr;
  if (!attr) {
    RUNIA_CUDA(cudaFuncSetAttribute(runia::pred_uncertainty_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 132 * 1024));
    attr = true;
  }
  runia::pred_uncertainty_kernel<<<(unsigned)runia::ceil_div(n_items, 8), 256, smem, (cudaStream_t)stream>>>(
      logits, n_items, n_mc, C, pred_h, mi);
  runia::count_launch();
  return runia::finish_launch("pred_uncertainty");
}

// ------------------------------------------------------------------------------------------
// (f3) spatial reduction of convolutional activation maps feeding the entropy path --
// feature_extraction/utils.py:70-92 (`get_mean_or_fullmean_ls_sample`): x [P, H, W] (P = batch * channels)
//   fullmean: out[p] = mean_h mean_w x[p, h, w];   mean: out[p, h] = mean_w x[p, h, w].
// One warp per plane (fullmean) or per row group (mean); every byte is read once, coalesced.
// ------------------------------------------------------------------------------------------
namespace runia {

__global__ void __launch_bounds__(256) spatial_mean_kernel(const float *__restrict__ x, int64_t P, int H, int W, int full,
                                                           float *__restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t wid = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5), nw = (int64_t)gridDim.x * 8;
  if (full) {
    const int hw = H * W;
    const float inv_w = 1.f / (float)W, inv_h = 1.f / (float)H;
    for (int64_t p = wid; p < P; p += nw) {
      const float *src = x + p * hw;
      float s = 0.f;
      for (int e = lane; e < hw; e += 32) s += __ldg(src + e);
      s = warp_sum32(s);
      if (lane == 0) out[p] = (s * inv_w) * inv_h;
    }
  } else {
    const int64_t rows = P * H;
    const float inv_w = 1.f / (float)W;
    if (W <= 32) {  // several rows per warp pass: lane l reads element l of a 32-float window, rows resolved per lane
      for (int64_t r = wid; r < rows; r += nw) {
        const float v = lane < W ? __ldg(x + r * W + lane) : 0.f;
        const float s = warp_sum32(v);
        if (lane == 0) out[r] = s * inv_w;
      }
    } else {
      for (int64_t r = wid; r < rows; r += nw) {
        float s = 0.f;
        for (int e = lane; e < W; e += 32) s += __ldg(x + r * W + e);
        s = warp_sum32(s);
        if (lane == 0) out[r] = s * inv_w;
      }
    }
  }
}

}  // namespace runia

extern "C" int runia_spatial_mean_f32(const float *x, int64_t P, int H, int W, int fullmean, float *out, void *stream) {
  RUNIA_NVTX();
  RUNIA_REQUIRE(P >= 0 && H > 0 && W > 0, RUNIA_E_BADARG, "spatial_mean: bad sizes");
  if (P == 0) return RUNIA_OK;
  RUNIA_REQUIRE(x && out, RUNIA_E_BADARG, "spatial_mean: null pointer");
  const int64_t units = fullmean ? P : P * H;
  const unsigned grid = (unsigned)std::min<int64_t>(runia::ceil_div(units, 8), (int64_t)runia::kNumSMs * 16);
  runia::spatial_mean_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, P, H, W, fullmean, out);
  runia::count_launch();
  return runia::finish_launch("spatial_mean");
}
