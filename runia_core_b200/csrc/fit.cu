// (f2) setup() statistics on the device: the mean / covariance halves of MDLatentSpace.setup
// (inference/postprocessors.py:202-226), cMDLatentSpace.setup (:283-318) and mahalanobis_preprocess
// (inference/funcs.py:33-66), which the reference computes with NumPy + sklearn EmpiricalCovariance.
//
//  * class means: NumPy reduces a C-contiguous [n, d] float32 array over axis 0 row by row, in float32
//    (no pairwise blocking on that axis), then divides by float32(n).  class_mean_seq_kernel keeps exactly
//    that order -- one thread per (class, column), rows in ascending order -- so the means are bit-identical
//    to `feats[labels == c].mean(0)`.
//  * covariance: np.cov promotes the float32 class-centred rows to float64 and takes X^T X with a float64
//    GEMM.  gram_f64_kernel forms f32(x - mu_c) exactly like NumPy's float32 subtraction, widens, and
//    accumulates the upper-triangular 64 x 64 tiles of G = sum_i r_i r_i^T with DFMA; rows are split over
//    CTAs and the per-split partial tiles are added in a fixed order (deterministic, no float atomics).
//    The column sums of the residuals come out of the diagonal tiles so the host can apply np.cov's own
//    re-centring: cov = (G - n a a^T) / n, a = colsum / n.
#include <algorithm>
#include <type_traits>

#include "common.cuh"

namespace runia {

constexpr int FIT_CHUNK = 2048;  // labels scanned per round of the mean kernel

__global__ void __launch_bounds__(128) class_mean_seq_kernel(const float *__restrict__ X, const int32_t *__restrict__ labels,
                                                             int64_t N, int d, float *__restrict__ means,
                                                             int64_t *__restrict__ counts) {
  __shared__ int rows[FIT_CHUNK];
  __shared__ int s_cnt;
  const int c = blockIdx.y;
  const int col = blockIdx.x * 128 + threadIdx.x;
  const bool live = col < d;
  const float *xc = X + (live ? col : 0);
  float acc = 0.f;
  int64_t total = 0;
  for (int64_t base = 0; base < N; base += FIT_CHUNK) {
    const int span = (int)min((int64_t)FIT_CHUNK, N - base);
    int cnt;
    if (labels) {
      __syncthreads();
      if (threadIdx.x < 32) {  // ordered compaction of this chunk's rows of class c
        int n = 0;
        for (int r = 0; r < span; r += 32) {
          const int i = r + threadIdx.x;
          const bool hit = i < span && labels[base + i] == c;
          const unsigned m = __ballot_sync(0xffffffffu, hit);
          if (hit) rows[n + __popc(m & ((1u << threadIdx.x) - 1u))] = i;
          n += __popc(m);
        }
        if (threadIdx.x == 0) s_cnt = n;
      }
      __syncthreads();
      cnt = s_cnt;
    } else {
      cnt = span;
    }
    total += cnt;
    if (!live) continue;
    const float *xb = xc + (size_t)base * d;
    int q = 0;
    for (; q + 8 <= cnt; q += 8) {
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = __ldg(xb + (size_t)(labels ? rows[q + u] : q + u) * d);
#pragma unroll
      for (int u = 0; u < 8; ++u) acc = __fadd_rn(acc, v[u]);
    }
    for (; q < cnt; ++q) acc = __fadd_rn(acc, __ldg(xb + (size_t)(labels ? rows[q] : q) * d));
  }
  if (live) means[(size_t)c * d + col] = __fdiv_rn(acc, (float)total);  // 0 / 0 = NaN for an empty class, like NumPy
  if (counts && blockIdx.x == 0 && threadIdx.x == 0) counts[c] = total;
}

constexpr int GT = 64;        // Gram tile edge
constexpr int GK = 32;        // rows per staged chunk (compute per chunk ~ one HBM round trip: the prefetch hides)
constexpr int GQ = GK / 4;    // rows fetched per thread and chunk
constexpr int G_THREADS = 256;

// residual of one element: f32(x - mu_label) widened to f64; rows whose label is outside [0, C) contribute 0
__device__ __forceinline__ float gram_fetch(const float *__restrict__ X, const float *__restrict__ centers, int d, int64_t row,
                                            int lab, int col, bool ok) {
  if (!ok) return 0.f;
  const float x = __ldg(X + (size_t)row * d + col);
  return centers ? __fsub_rn(x, __ldg(centers + (size_t)lab * d + col)) : x;
}

// same with ONE float64 centre: r = f64(x) - u, exact in float64 (what NumPy computes for `train_f32 - u_f64`:
// ViM.setup's EmpiricalCovariance(assume_centered=True).fit(train - u), postprocessors.py:1060-1064)
__device__ __forceinline__ double gram_fetch64(const float *__restrict__ X, const double *__restrict__ center, int d, int64_t row,
                                               int col, bool ok) {
  if (!ok) return 0.0;
  return __dsub_rn((double)__ldg(X + (size_t)row * d + col), __ldg(center + col));
}

template <bool F64C>
__global__ void __launch_bounds__(G_THREADS) gram_f64_kernel(const float *__restrict__ X, const int32_t *__restrict__ labels,
                                                             const float *__restrict__ centers,
                                                             const double *__restrict__ center64, int64_t N, int d, int C,
                                                             int nt, int64_t rows_per_split, double *__restrict__ part,
                                                             double *__restrict__ part_cs) {
  using RT = typename std::conditional<F64C, double, float>::type;
  __shared__ __align__(16) double As[GK][GT];
  __shared__ __align__(16) double Bs[GK][GT];
  // blockIdx.x enumerates the upper-triangular tile pairs (ti <= tj)
  int ti = 0, rem = blockIdx.x;
  while (rem >= nt - ti) {
    rem -= nt - ti;
    ++ti;
  }
  const int tj = ti + rem;
  const bool diag = ti == tj;
  const int split = blockIdx.y;
  const int64_t r0 = (int64_t)split * rows_per_split, r1 = min(N, r0 + rows_per_split);
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  // staging assignment: element e = tid + 256 q -> (row e / 64, column e % 64)
  const int lcol = tid & 63, lrow0 = tid >> 6;  // rows lrow0, lrow0 + 4, ... (GQ of them)
  const int colA = ti * GT + lcol, colB = tj * GT + lcol;
  const bool okA = colA < d, okB = colB < d;

  double acc[4][4] = {};
  double cs[4] = {};
  RT ra[GQ], rb[GQ];
  auto fetch = [&](int64_t rbase) {
#pragma unroll
    for (int q = 0; q < GQ; ++q) {
      const int64_t row = rbase + lrow0 + 4 * q;
      int lab = 0;
      bool in = row < r1;
      if (in && labels) {
        lab = labels[row];
        in = lab >= 0 && lab < C;
      }
      if constexpr (F64C) {
        ra[q] = gram_fetch64(X, center64, d, row, colA, in && okA);
        rb[q] = diag ? 0.0 : gram_fetch64(X, center64, d, row, colB, in && okB);
      } else {
        ra[q] = gram_fetch(X, centers, d, row, lab, colA, in && okA);
        rb[q] = diag ? 0.f : gram_fetch(X, centers, d, row, lab, colB, in && okB);
      }
    }
  };
  if (r0 < r1) fetch(r0);
  for (int64_t rbase = r0; rbase < r1; rbase += GK) {
    __syncthreads();
#pragma unroll
    for (int q = 0; q < GQ; ++q) {
      As[lrow0 + 4 * q][lcol] = (double)ra[q];
      if (!diag) Bs[lrow0 + 4 * q][lcol] = (double)rb[q];
    }
    __syncthreads();
    if (rbase + GK < r1) fetch(rbase + GK);
    const double(*Bt)[GT] = diag ? As : Bs;
#pragma unroll
    for (int kk = 0; kk < GK; ++kk) {
      const double2 a01 = *reinterpret_cast<const double2 *>(&As[kk][ty * 4]);
      const double2 a23 = *reinterpret_cast<const double2 *>(&As[kk][ty * 4 + 2]);
      const double2 b01 = *reinterpret_cast<const double2 *>(&Bt[kk][tx * 4]);
      const double2 b23 = *reinterpret_cast<const double2 *>(&Bt[kk][tx * 4 + 2]);
      const double a[4] = {a01.x, a01.y, a23.x, a23.y}, b[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
      if (diag && ty == 0) {
#pragma unroll
        for (int j = 0; j < 4; ++j) cs[j] += b[j];
      }
    }
  }
  double *out = part + ((size_t)split * gridDim.x + blockIdx.x) * (GT * GT);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    *reinterpret_cast<double2 *>(&out[(ty * 4 + i) * GT + tx * 4]) = make_double2(acc[i][0], acc[i][1]);
    *reinterpret_cast<double2 *>(&out[(ty * 4 + i) * GT + tx * 4 + 2]) = make_double2(acc[i][2], acc[i][3]);
  }
  if (diag && ty == 0) {
    double *o = part_cs + ((size_t)split * nt + ti) * GT + tx * 4;
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j] = cs[j];
  }
}

// G[i][j] = sum over splits (ascending) of the partial tile holding (min, max); colsum likewise
__global__ void gram_reduce_kernel(const double *__restrict__ part, const double *__restrict__ part_cs, int d, int nt, int npairs,
                                   int splits, double *__restrict__ G, double *__restrict__ colsum) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < (int64_t)d * d) {
    const int i = (int)(e / d), j = (int)(e % d);
    const int lo = min(i, j), hi = max(i, j);
    const int ti = lo / GT, tj = hi / GT;
    const int pair = ti * nt - ti * (ti - 1) / 2 + (tj - ti);
    const double *p = part + (size_t)pair * (GT * GT) + (lo % GT) * GT + (hi % GT);
    double s = 0.0;
    for (int k = 0; k < splits; ++k) s = __dadd_rn(s, p[(size_t)k * npairs * (GT * GT)]);
    G[e] = s;
  }
  if (colsum && e < d) {
    double s = 0.0;
    for (int k = 0; k < splits; ++k) s = __dadd_rn(s, part_cs[((size_t)k * nt + e / GT) * GT + e % GT]);
    colsum[e] = s;
  }
}

struct GramPlan {
  int nt, npairs, splits;
  int64_t rows_per_split;
  size_t part_bytes, cs_bytes;
};

static GramPlan gram_plan(int64_t N, int d) {
  GramPlan p;
  p.nt = (int)ceil_div(d, GT);
  p.npairs = p.nt * (p.nt + 1) / 2;
  int64_t s = ceil_div(4 * kNumSMs, p.npairs);
  s = std::min<int64_t>(s, ceil_div(N, 8 * GK));
  s = std::min<int64_t>(std::max<int64_t>(s, 1), 65535);
  p.rows_per_split = ceil_div(ceil_div(N, s), GK) * GK;
  p.splits = (int)ceil_div(N, p.rows_per_split);
  p.part_bytes = (size_t)p.splits * p.npairs * GT * GT * sizeof(double);
  p.cs_bytes = (size_t)p.splits * p.nt * GT * sizeof(double);
  return p;
}

}  // namespace runia

using namespace runia;

extern "C" int runia_class_mean_f32(const float *X, const int32_t *labels, int64_t N, int d, int C, float *means,
                                    int64_t *counts, void *stream) {
  RUNIA_NVTX();
  RUNIA_REQUIRE(N >= 0 && d >= 1 && C >= 1, RUNIA_E_BADARG, "class_mean: needs N >= 0, d >= 1, C >= 1");
  RUNIA_REQUIRE(C <= 65535, RUNIA_E_UNSUPPORTED, "class_mean: C=%d classes not supported (max 65535)", C);
  RUNIA_REQUIRE(labels || C == 1, RUNIA_E_BADARG, "class_mean: C > 1 needs labels");
  RUNIA_REQUIRE((X || N == 0) && means, RUNIA_E_BADARG, "class_mean: null pointer");
  dim3 grid((unsigned)ceil_div(d, 128), (unsigned)C);
  class_mean_seq_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(X, labels, N, d, means, counts);
  count_launch();
  return finish_launch("class_mean");
}

extern "C" size_t runia_centered_gram_workspace_bytes(int64_t N, int d) {
  if (N < 0 || d < 1) return 0;
  const GramPlan p = gram_plan(N > 0 ? N : 1, d);
  return p.part_bytes + p.cs_bytes;
}

extern "C" int runia_centered_gram_f64(const float *X, const int32_t *labels, const float *centers, int64_t N, int d, int C,
                                       double *G, double *colsum, void *ws, size_t ws_bytes, void *stream) {
  RUNIA_NVTX();
  RUNIA_REQUIRE(N >= 1 && d >= 1 && C >= 1, RUNIA_E_BADARG, "centered_gram: needs N >= 1, d >= 1, C >= 1");
  RUNIA_REQUIRE(X && G && ws, RUNIA_E_BADARG, "centered_gram: null pointer");
  RUNIA_REQUIRE(labels || C == 1, RUNIA_E_BADARG, "centered_gram: C > 1 needs labels");
  RUNIA_REQUIRE(d <= 8192, RUNIA_E_UNSUPPORTED, "centered_gram: d=%d not supported (max 8192)", d);
  const GramPlan p = gram_plan(N, d);
  RUNIA_REQUIRE(ws_bytes >= p.part_bytes + p.cs_bytes, RUNIA_E_BADARG, "centered_gram: workspace of %zu bytes, %zu needed", ws_bytes,
                p.part_bytes + p.cs_bytes);
  double *part = (double *)ws, *part_cs = (double *)((char *)ws + p.part_bytes);
  cudaStream_t st = (cudaStream_t)stream;
  gram_f64_kernel<false><<<dim3((unsigned)p.npairs, (unsigned)p.splits), G_THREADS, 0, st>>>(
      X, labels, centers, nullptr, N, d, C, p.nt, p.rows_per_split, part, part_cs);
  const int64_t total = (int64_t)d * d;
  gram_reduce_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, st>>>(part, part_cs, d, p.nt, p.npairs, p.splits, G, colsum);
  count_launch(2);
  return finish_launch("centered_gram");
}

extern "C" int runia_shifted_gram_f64(const float *X, const double *center, int64_t N, int d, double *G, double *colsum, void *ws,
                                      size_t ws_bytes, void *stream) {
  RUNIA_NVTX();
  RUNIA_REQUIRE(N >= 1 && d >= 1, RUNIA_E_BADARG, "shifted_gram: needs N >= 1, d >= 1");
  RUNIA_REQUIRE(X && center && G && ws, RUNIA_E_BADARG, "shifted_gram: null pointer");
  RUNIA_REQUIRE(d <= 8192, RUNIA_E_UNSUPPORTED, "shifted_gram: d=%d not supported (max 8192)", d);
  const GramPlan p = gram_plan(N, d);
  RUNIA_REQUIRE(ws_bytes >= p.part_bytes + p.cs_bytes, RUNIA_E_BADARG, "shifted_gram: workspace of %zu bytes, %zu needed", ws_bytes,
                p.part_bytes + p.cs_bytes);
  double *part = (double *)ws, *part_cs = (double *)((char *)ws + p.part_bytes);
  cudaStream_t st = (cudaStream_t)stream;
  gram_f64_kernel<true><<<dim3((unsigned)p.npairs, (unsigned)p.splits), G_THREADS, 0, st>>>(
      X, nullptr, nullptr, center, N, d, 1, p.nt, p.rows_per_split, part, part_cs);
  const int64_t total = (int64_t)d * d;
  gram_reduce_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, st>>>(part, part_cs, d, p.nt, p.npairs, p.splits, G, colsum);
  count_launch(2);
  return finish_launch("shifted_gram");
}
