// HBM-bound single-pass scorers:
//  (a8)  Energy / MSP / GEN from logits in one read           postprocessors.py:519-691
//  (a10) ReAct / DICE / DICE+ReAct: clip -> linear -> LSE      postprocessors.py:1325-1621
//        ASH-S: per-row top-k pruning + rescale -> linear -> LSE   funcs.py:230-261
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

__global__ void __launch_bounds__(LS_ROWS)
logit_scores_small_kernel(const float *__restrict__ logits, int64_t N, int C, float gamma, int M,
                          float *__restrict__ energy, float *__restrict__ msp, float *__restrict__ gen) {
  extern __shared__ float tile[];  // [LS_ROWS][C + pad]
  const int ldc = C | 1;           // odd stride -> conflict-free row reads
  const int64_t r0 = (int64_t)blockIdx.x * LS_ROWS;
  const int64_t rows = (N - r0 < LS_ROWS) ? (N - r0) : LS_ROWS;
  const int64_t total = rows * C;
  const float *src = logits + r0 * C;
  for (int64_t e = threadIdx.x; e < total; e += LS_ROWS) {
    const int rr = (int)(e / C), cc = (int)(e % C);
    tile[rr * ldc + cc] = __ldg(src + e);
  }
  __syncthreads();
  if (threadIdx.x >= rows) return;
  const float *l = tile + threadIdx.x * ldc;
  float m = -INFINITY;
  for (int c = 0; c < C; ++c) m = fmaxf(m, l[c]);
  float s = 0.f;
  for (int c = 0; c < C; ++c) s += expf(l[c] - m);
  const int64_t row = r0 + threadIdx.x;
  if (energy) energy[row] = logf(s) + m;
  if (msp) msp[row] = 1.f / s;  // exp(m - m) / s
  if (gen) {
    float g = 0.f;
    if (M >= C) {
      for (int c = 0; c < C; ++c) g += gen_term(expf(l[c] - m) / s, gamma);
    } else {
      // the M largest probabilities under the total order (value, index)
      for (int c = 0; c < C; ++c) {
        const float lc = l[c];
        int greater = 0;
        for (int o = 0; o < C; ++o) greater += (l[o] > lc || (l[o] == lc && o > c)) ? 1 : 0;
        if (greater < M) g += gen_term(expf(lc - m) / s, gamma);
      }
    }
    gen[row] = -g;
  }
}

// large C: one warp per row (M must cover all classes)
__global__ void __launch_bounds__(256)
logit_scores_wide_kernel(const float *__restrict__ logits, int64_t N, int C, float gamma,
                         float *__restrict__ energy, float *__restrict__ msp, float *__restrict__ gen) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= N) return;
  const float *l = logits + row * (int64_t)C;
  float m = -INFINITY;
  for (int c = lane; c < C; c += 32) m = fmaxf(m, __ldg(l + c));
  m = warp_max32(m);
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += expf(__ldg(l + c) - m);
  s = warp_sum32(s);
  float g = 0.f;
  if (gen) {
    for (int c = lane; c < C; c += 32) g += gen_term(expf(__ldg(l + c) - m) / s, gamma);
    g = warp_sum32(g);
  }
  if (lane == 0) {
    if (energy) energy[row] = logf(s) + m;
    if (msp) msp[row] = 1.f / s;
    if (gen) gen[row] = -g;
  }
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
      // radix select on order-preserving integer keys
      auto keyof = [](float v) -> uint32_t {
        const uint32_t u = __float_as_uint(v);
        return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
      };
      uint32_t prefix = 0;
      for (int bit = 31; bit >= 0; --bit) {
        const uint32_t cand = prefix | (1u << bit);
        int cnt = 0;
        for (int j = lane; j < d; j += 32) cnt += (keyof(__ldg(x + j)) >= cand) ? 1 : 0;
        cnt = __reduce_add_sync(0xffffffffu, cnt);
        if (cnt >= k_keep) prefix = cand;
      }
      tkey = prefix;
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
      const float s2 = s_gt + (float)n_ties_keep * tval;
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

}  // namespace runia

using namespace runia;

extern "C" int runia_logit_scores_f32(const float *logits, int64_t N, int C, float gamma, int M, float *energy,
                                      float *msp, float *gen, void *stream) {
  RUNIA_REQUIRE(N >= 0 && C > 0, RUNIA_E_BADARG, "logit_scores: bad sizes");
  if (N == 0) return RUNIA_OK;
  RUNIA_REQUIRE(logits && (energy || msp || gen), RUNIA_E_BADARG, "logit_scores: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if (C <= 64) {
    const size_t smem = (size_t)LS_ROWS * (C | 1) * sizeof(float);
    static bool attr = false;
    if (!attr) {
      RUNIA_CUDA(cudaFuncSetAttribute(logit_scores_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024));
      attr = true;
    }
    logit_scores_small_kernel<<<(unsigned)ceil_div(N, LS_ROWS), LS_ROWS, smem, st>>>(logits, N, C, gamma, M, energy,
                                                                                 msp, gen);
  } else {
    RUNIA_REQUIRE(!gen || M >= C, RUNIA_E_UNSUPPORTED, "logit_scores: GEN with M=%d < C=%d needs C <= 64", M, C);
    logit_scores_wide_kernel<<<(unsigned)ceil_div(N, 8), 256, 0, st>>>(logits, N, C, gamma, energy, msp, gen);
  }
  count_launch();
  return finish_launch("logit_scores");
}

static int launch_linear_lse(bool ash, const float *X, int64_t N, int d, const float *W, const float *b, int C,
                             float clip, int k_keep, float *out, void *stream) {
  RUNIA_REQUIRE(N >= 0 && d > 0 && C > 0, RUNIA_E_BADARG, "linear_lse: bad sizes");
  RUNIA_REQUIRE(C <= CL_MAXC && (size_t)C * d * 4 <= 200 * 1024, RUNIA_E_UNSUPPORTED,
                "linear_lse: C=%d, d=%d exceed the shared-memory weight tile (C <= %d, C*d*4 <= 200 KiB)", C, d,
                CL_MAXC);
  if (N == 0) return RUNIA_OK;
  RUNIA_REQUIRE(X && W && b && out, RUNIA_E_BADARG, "linear_lse: null pointer");
  const size_t smem = (size_t)C * d * sizeof(float);
  static bool attr = false;
  if (!attr) {
    RUNIA_CUDA(cudaFuncSetAttribute(linear_lse_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    RUNIA_CUDA(cudaFuncSetAttribute(linear_lse_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr = true;
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
  return launch_linear_lse(false, X, N, d, W, b, C, clip, 0, out, stream);
}

extern "C" int runia_ash_linear_lse_f32(const float *X, int64_t N, int d, const float *W, const float *b, int C,
                                        int k_keep, float *out, void *stream) {
  RUNIA_REQUIRE(k_keep >= 1 && k_keep <= d, RUNIA_E_BADARG, "ash_linear_lse: k_keep=%d outside [1, d=%d]", k_keep, d);
  return launch_linear_lse(true, X, N, d, W, b, C, INFINITY, k_keep, out, stream);
}
