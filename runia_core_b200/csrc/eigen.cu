// (f4) EigenScore -- llm_uncertainty/scores.py:49-66: mean log singular value of cov(E^T) + alpha I for
// an embedding matrix E [n samples, d hidden] (float32).  The reference builds the d x d covariance
// (4096 x 4096) and takes a full float64 SVD; the d x d matrix has rank <= n - 1 and shares its
// non-zero eigenvalues with the n x n Gram matrix of the centred samples, so
//   score = ( sum_{i<n} log(lambda_i(G) + alpha) + (d - n) log(alpha) ) / d,   G = Ec Ec^T / (n - 1).
// One CTA: column means, the Gram matrix in float64 (one warp per (i, j) pair, fixed summation order),
// cyclic Jacobi on the n x n matrix, the log-sum.  n <= 32.
#include "common.cuh"

namespace runia {

constexpr int EG_MAXN = 32;

__global__ void __launch_bounds__(256) eigen_score_kernel(const float *__restrict__ E, int n, int d, double alpha,
                                                          double *__restrict__ out) {
  extern __shared__ double mean[];  // [d]
  __shared__ double G[EG_MAXN][EG_MAXN + 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int k = threadIdx.x; k < d; k += blockDim.x) {
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += (double)E[(size_t)i * d + k];
    mean[k] = s / (double)n;
  }
  __syncthreads();
  const int npairs = n * (n + 1) / 2;
  for (int p = warp; p < npairs; p += blockDim.x / 32) {
    int i = 0, rem = p;  // p-th pair (i <= j), row-major upper triangle
    while (rem >= n - i) {
      rem -= n - i;
      ++i;
    }
    const int j = i + rem;
    double acc = 0.0;
    for (int k = lane; k < d; k += 32) {
      const double m = mean[k];
      acc = fma((double)E[(size_t)i * d + k] - m, (double)E[(size_t)j * d + k] - m, acc);
    }
    acc = warp_tree_sum_f64(acc);
    if (lane == 0) {
      const double g = acc / (double)(n > 1 ? n - 1 : 1);
      G[i][j] = g;
      G[j][i] = g;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    // cyclic Jacobi: eigenvalues end up on the diagonal
    for (int sweep = 0; sweep < 30; ++sweep) {
      double off = 0.0, diag = 0.0;
      for (int a = 0; a < n; ++a) {
        diag += G[a][a] * G[a][a];
        for (int b = a + 1; b < n; ++b) off += G[a][b] * G[a][b];
      }
      if (off <= 1e-30 * diag || off == 0.0) break;
      for (int a = 0; a < n - 1; ++a)
        for (int b = a + 1; b < n; ++b) {
          const double apq = G[a][b];
          if (apq == 0.0) continue;
          const double theta = (G[b][b] - G[a][a]) / (2.0 * apq);
          const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
          const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
          for (int k = 0; k < n; ++k) {  // columns a, b
            const double gka = G[k][a], gkb = G[k][b];
            G[k][a] = c * gka - s * gkb;
            G[k][b] = s * gka + c * gkb;
          }
          for (int k = 0; k < n; ++k) {  // rows a, b
            const double gak = G[a][k], gbk = G[b][k];
            G[a][k] = c * gak - s * gbk;
            G[b][k] = s * gak + c * gbk;
          }
        }
    }
    double sum = 0.0;
    const int m = n < d ? n : d;
    for (int a = 0; a < m; ++a) sum += log(fmax(G[a][a], 0.0) + alpha);  // G is positive semi-definite
    // for n > d (more samples than dimensions) the Gram spectrum already holds every non-zero value; the
    // d x d covariance has exactly d singular values: the d largest of G's
    if (d > n) sum += (double)(d - n) * log(alpha);
    out[0] = sum / (double)d;
  }
}

}  // namespace runia

using namespace runia;

extern "C" int runia_eigen_score_f32(const float *E, int n, int d, double alpha, double *out, void *stream) {
  RUNIA_NVTX();
  RUNIA_REQUIRE(n >= 2 && d >= 1 && alpha > 0.0, RUNIA_E_BADARG, "eigen_score: needs n >= 2 samples, d >= 1, alpha > 0");
  RUNIA_REQUIRE(n <= EG_MAXN && n <= d && (size_t)d * 8 <= 200 * 1024, RUNIA_E_UNSUPPORTED,
                "eigen_score: n=%d samples (max %d, at most d) or d=%d (max 25600) not supported", n, EG_MAXN, d);
  RUNIA_REQUIRE(E && out, RUNIA_E_BADARG, "eigen_score: null pointer");
  static PerDeviceFlag attr;
  if (!attr) {
    RUNIA_CUDA(cudaFuncSetAttribute(eigen_score_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr = true;
  }
  eigen_score_kernel<<<1, 256, (size_t)d * sizeof(double), (cudaStream_t)stream>>>(E, n, d, alpha, out);
  count_launch();
  return finish_launch("eigen_score");
}
