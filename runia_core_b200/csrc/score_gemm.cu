// Row scorers built on the FP32 tile contraction: PCA projection (a2), LaREM Mahalanobis (a3),
// ViM residual (a7), class-conditional Mahalanobis (a6), DDU/GMM log-density (a9).
// Each CTA owns 128 input rows and walks ALL output columns in 128-wide panels, so every
// row-wise reduction (sum of squares, max over classes, log-sum-exp) completes inside one CTA:
// no atomics, deterministic, and the [N, r] intermediate never reaches HBM.
#include <algorithm>

#include "rowgemm.cuh"

namespace runia {

// ------------------------------------------------------------------------------------------
// (a2) Z = ((X - mean) C^T) * inv_scale        dimensionality_reduction.py:75-87
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(GEMM_THREADS, 2)
pca_transform_kernel(const float *__restrict__ X, int64_t N, int D0, const float *__restrict__ mean,
                     const float *__restrict__ Ct, int d, const float *__restrict__ inv_scale,
                     const float *__restrict__ bias, float *__restrict__ Z) {
  __shared__ GemmSmem sm;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const Prologue pro{mean, INFINITY};
  const bool vec_store = (d % 4 == 0) && ((reinterpret_cast<uintptr_t>(Z) & 15) == 0);
  for (int n0 = 0; n0 < d; n0 += BN) {
    float acc[8][8];
    zero_acc(acc);
    gemm_mainloop(X, N, m0, Ct, d, n0, D0, pro, sm, acc);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int64_t row = m0 + tile_row(ty, i);
      if (row >= N) continue;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int col = n0 + tile_col(tx, h * 4);
        float v[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float sc = (inv_scale && col + q < d) ? __ldg(inv_scale + col + q) : 1.f;
          v[q] = acc[i][h * 4 + q] * sc;
          if (bias && col + q < d) v[q] += __ldg(bias + col + q);
        }
        if (vec_store && col + 3 < d) {
          *reinterpret_cast<float4 *>(Z + row * d + col) = make_float4(v[0], v[1], v[2], v[3]);
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (col + q < d) Z[row * d + col + q] = v[q];
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// (a3)/(a7) out = -sum_j sign_j ((x - mu) . w_j)^2   [MD]   or   -alpha*sqrt(.) + lse(logits) [ViM]
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(GEMM_THREADS, 2)
rownorm_kernel(const float *__restrict__ X, int64_t N, int d, const float *__restrict__ mu,
               const float *__restrict__ Wt, int r, const float *__restrict__ sign, int mode,
               const float *__restrict__ logits, int C, float alpha, double *__restrict__ out64,
               float *__restrict__ out32) {
  __shared__ GemmSmem sm;
  __shared__ float rowsum[BM];
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const Prologue pro{mu, INFINITY};
  float rowacc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int n0 = 0; n0 < r; n0 += BN) {
    float acc[8][8];
    zero_acc(acc);
    gemm_mainloop(X, N, m0, Wt, r, n0, d, pro, sm, acc);
    float sg[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int col = n0 + tile_col(tx, j);
      sg[j] = (col < r) ? (sign ? __ldg(sign + col) : 1.f) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) rowacc[i] = fmaf(sg[j] * acc[i][j], acc[i][j], rowacc[i]);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float v = warp_sum16(rowacc[i]);
    if (tx == 0) rowsum[tile_row(ty, i)] = v;
  }
  __syncthreads();
  if (threadIdx.x < BM) {
    const int64_t row = m0 + threadIdx.x;
    if (row < N) {
      const float v = rowsum[threadIdx.x];
      if (mode == RUNIA_ROWNORM_MD) {
        if (out64) out64[row] = -(double)v;
        if (out32) out32[row] = -v;
      } else {
        const float *l = logits + row * (int64_t)C;
        float m = -INFINITY;
        for (int c = 0; c < C; ++c) m = fmaxf(m, l[c]);
        float s = 0.f;
        for (int c = 0; c < C; ++c) s += expf(l[c] - m);
        const float sc = -alpha * sqrtf(fmaxf(v, 0.f)) + (m + logf(s));
        if (out64) out64[row] = (double)sc;
        if (out32) out32[row] = sc;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// (a6) out = max_c -sum_j sign_j (y_j - m_cj)^2,  y = (x - g) Wt^T      funcs.py:87-102
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(GEMM_THREADS, 2)
classcond_kernel(const float *__restrict__ X, int64_t N, int d, const float *__restrict__ g,
                 const float *__restrict__ Wt, int r, const float *__restrict__ sign,
                 const float *__restrict__ Mc, const int32_t *__restrict__ valid, int C, int accumulate,
                 double *__restrict__ out64, float *__restrict__ out32) {
  // C = classes of THIS launch (Mc / valid already point at the chunk); accumulate: fold with the score the
  // previous chunk left in out (more than 256 classes are scored 256 at a time)
  __shared__ GemmSmem sm;
  extern __shared__ float cls[];  // [BM][C]
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const Prologue pro{g, INFINITY};
  for (int e = threadIdx.x; e < BM * C; e += GEMM_THREADS) cls[e] = 0.f;
  for (int n0 = 0; n0 < r; n0 += BN) {
    float acc[8][8];
    zero_acc(acc);
    gemm_mainloop(X, N, m0, Wt, r, n0, d, pro, sm, acc);
    float sg[8];
    int colj[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      colj[j] = n0 + tile_col(tx, j);
      sg[j] = (colj[j] < r) ? (sign ? __ldg(sign + colj[j]) : 1.f) : 0.f;
    }
    for (int c = 0; c < C; ++c) {
      if (!valid[c]) continue;
      float mc[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) mc[j] = (colj[j] < r) ? __ldg(Mc + (int64_t)c * r + colj[j]) : 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float p = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float dlt = acc[i][j] - mc[j];
          p = fmaf(sg[j] * dlt, dlt, p);
        }
        p = warp_sum16(p);
        if (tx == 0) cls[tile_row(ty, i) * C + c] += p;  // single owner per (row, c)
      }
    }
  }
  __syncthreads();
  if (threadIdx.x < BM) {
    const int64_t row = m0 + threadIdx.x;
    if (row < N) {
      float best = -INFINITY;
      if (accumulate) best = out64 ? (float)out64[row] : out32[row];
      for (int c = 0; c < C; ++c)
        if (valid[c]) {
          const float sc = -cls[threadIdx.x * C + c];
          if (sc > best) best = sc;  // NaN never wins, like np.max after NaN -> -inf
        }
      if (out64) out64[row] = (double)best;
      if (out32) out32[row] = best;
    }
  }
}

// ------------------------------------------------------------------------------------------
// (a9) out = logsumexp_c ( -0.5 |A_c x - A_c mu_c|^2 + logconst_c )     funcs.py:265-344
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(GEMM_THREADS, 2)
gmm_lse_kernel(const float *__restrict__ X, int64_t N, int d, const float *__restrict__ At,
               const float *__restrict__ off, int dpad, const float *__restrict__ logconst, int C,
               float *__restrict__ out) {
  __shared__ GemmSmem sm;
  __shared__ float run_m[BM], run_s[BM];
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const Prologue pro{nullptr, INFINITY};
  if (threadIdx.x < BM) {
    run_m[threadIdx.x] = -INFINITY;
    run_s[threadIdx.x] = 0.f;
  }
  const int64_t NB = (int64_t)C * dpad;
  for (int c = 0; c < C; ++c) {
    float rowacc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int p0 = 0; p0 < dpad; p0 += BN) {
      const int64_t n0 = (int64_t)c * dpad + p0;
      float acc[8][8];
      zero_acc(acc);
      gemm_mainloop(X, N, m0, At, NB, n0, d, pro, sm, acc);
      float of[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) of[j] = __ldg(off + n0 + tile_col(tx, j));
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float dlt = acc[i][j] - of[j];
          rowacc[i] = fmaf(dlt, dlt, rowacc[i]);
        }
    }
    const float lc = __ldg(logconst + c);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float v = warp_sum16(rowacc[i]);
      if (tx == 0) {
        const int rr = tile_row(ty, i);
        const float lp = fmaf(-0.5f, v, lc);
        const float m_old = run_m[rr];
        const float m_new = fmaxf(m_old, lp);
        if (m_new == -INFINITY) continue;
        run_s[rr] = run_s[rr] * expf(m_old - m_new) + expf(lp - m_new);
        run_m[rr] = m_new;
      }
    }
  }
  __syncthreads();
  if (threadIdx.x < BM) {
    const int64_t row = m0 + threadIdx.x;
    if (row < N) out[row] = run_m[threadIdx.x] + logf(run_s[threadIdx.x]);
  }
}

}  // namespace runia

using namespace runia;

extern "C" int runia_pca_transform_f32(const float *X, int64_t N, int D0, const float *mean,
                                       const float *components, int d, const float *inv_scale, float *Z,
                                       void *stream) {
  RUNIA_NVTX();
  RUNIA_REQUIRE(N >= 0 && D0 > 0 && d > 0, RUNIA_E_BADARG, "pca_transform: bad sizes N=%lld D0=%d d=%d",
                (long long)N, D0, d);
  if (N == 0) return RUNIA_OK;
  RUNIA_REQUIRE(X && components && Z, RUNIA_E_BADARG, "pca_transform: null pointer");
  const unsigned grid = (unsigned)ceil_div(N, BM);
  pca_transform_kernel<<<grid, GEMM_THREADS, 0, (cudaStream_t)stream>>>(X, N, D0, mean, components, d,
                                                                      inv_scale, nullptr, Z);
  count_launch();
  return finish_launch("pca_transform");
}

extern "C" int runia_linear_f32(const float *X, int64_t N, int d, const float *W, const float *b, int C, float *out,
                                void *stream) {
  RUNIA_NVTX();
  RUNIA_REQUIRE(N >= 0 && d > 0 && C > 0, RUNIA_E_BADARG, "linear: bad sizes N=%lld d=%d C=%d", (long long)N, d, C);
  if (N == 0) return RUNIA_OK;
  RUNIA_REQUIRE(X && W && out, RUNIA_E_BADARG, "linear: null pointer");
  const unsigned grid = (unsigned)ceil_div(N, BM);
  pca_transform_kernel<<<grid, GEMM_THREADS, 0, (cudaStream_t)stream>>>(X, N, d, nullptr, W, C, nullptr, b, out);
  count_launch();
  return finish_launch("linear");
}

extern "C" int runia_rownorm_score_f32(const float *X, int64_t N, int d, const float *mu, const float *Wt,
                                       int r, const float *sign, int mode, const float *logits, int C,
                                       float alpha, double *out_f64, float *out_f32, void *stream) {
  RUNIA_NVTX();
  RUNIA_REQUIRE(N >= 0 && d > 0 && r > 0, RUNIA_E_BADARG, "rownorm_score: bad sizes N=%lld d=%d r=%d",
                (long long)N, d, r);
  RUNIA_REQUIRE(mode == RUNIA_ROWNORM_MD || mode == RUNIA_ROWNORM_VIM, RUNIA_E_BADARG, "rownorm_score: bad mode");
  if (N == 0) return RUNIA_OK;
  RUNIA_REQUIRE(X && Wt && (out_f64 || out_f32), RUNIA_E_BADARG, "rownorm_score: null pointer");
  RUNIA_REQUIRE(mode != RUNIA_ROWNORM_VIM || (logits && C > 0), RUNIA_E_BADARG, "rownorm_score: ViM needs logits");
  const unsigned grid = (unsigned)ceil_div(N, BM);
  rownorm_kernel<<<grid, GEMM_THREADS, 0, (cudaStream_t)stream>>>(X, N, d, mu, Wt, r, sign, mode, logits, C,
                                                                alpha, out_f64, out_f32);
  count_launch();
  return finish_launch("rownorm_score");
}

extern "C" int runia_classcond_mahalanobis_f32(const float *X, int64_t N, int d, const float *g,
                                               const float *Wt, int r, const float *sign, const float *Mc,
                                               const int32_t *class_valid, int C, double *out_f64,
                                               float *out_f32, void *stream) {
  RUNIA_NVTX();
  RUNIA_REQUIRE(N >= 0 && d > 0 && r > 0 && C > 0, RUNIA_E_BADARG, "classcond: bad sizes");
  if (N == 0) return RUNIA_OK;
  RUNIA_REQUIRE(X && Wt && Mc && class_valid && (out_f64 || out_f32), RUNIA_E_BADARG, "classcond: null pointer");
  constexpr int kChunk = 256;  // classes per launch: [BM][C] partial sums live in shared memory
  static PerDeviceFlag attr_set;
  if (!attr_set) {
    RUNIA_CUDA(cudaFuncSetAttribute(classcond_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 140 * 1024));
    attr_set = true;
  }
  const unsigned grid = (unsigned)ceil_div(N, BM);
  for (int c0 = 0; c0 < C; c0 += kChunk) {
    const int cc = std::min(kChunk, C - c0);
    classcond_kernel<<<grid, GEMM_THREADS, (size_t)BM * cc * sizeof(float), (cudaStream_t)stream>>>(
        X, N, d, g, Wt, r, sign, Mc + (size_t)c0 * r, class_valid + c0, cc, c0 > 0 ? 1 : 0, out_f64, out_f32);
    count_launch();
  }
  return finish_launch("classcond_mahalanobis");
}

extern "C" int runia_gmm_lse_f32(const float *X, int64_t N, int d, const float *At, const float *off, int dpad,
                                 const float *logconst, int C, float *out, void *stream) {
  RUNIA_NVTX();
  RUNIA_REQUIRE(N >= 0 && d > 0 && C > 0 && dpad > 0 && dpad % BN == 0, RUNIA_E_BADARG,
                "gmm_lse: bad sizes (dpad must be a multiple of %d)", BN);
  if (N == 0) return RUNIA_OK;
  RUNIA_REQUIRE(X && At && off && logconst && out, RUNIA_E_BADARG, "gmm_lse: null pointer");
  const unsigned grid = (unsigned)ceil_div(N, BM);
  gmm_lse_kernel<<<grid, GEMM_THREADS, 0, (cudaStream_t)stream>>>(X, N, d, At, off, dpad, logconst, C, out);
  count_launch();
  return finish_launch("gmm_lse");
}
