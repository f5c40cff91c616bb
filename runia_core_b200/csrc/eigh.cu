// (f2) Symmetric eigendecomposition and Cholesky factorisation in float64 for the setup() fits:
//   * pinvh(covariance) of MDLatentSpace / cMD / Mahalanobis.setup (sklearn EmpiricalCovariance -> scipy.linalg.pinvh,
//     inference/postprocessors.py:212-220, 296-314; inference/funcs.py:62-66) -- pinvh IS an eigendecomposition with
//     a cut-off, so A = V diag(lambda) V^T from here gives the precision and, at once, the factor the scorers need;
//   * the opt-in exact PCA fit (eigenvectors of the covariance) and ViM's residual space;
//   * the per-class Cholesky factors of gmm_fit (inference/funcs.py:296-342).
//
// Eigendecomposition: one-sided (Hestenes) Jacobi on the columns of G = A, V = I.  A rotation of columns (p, q)
// zeroes g_p . g_q; at convergence the columns of G are orthogonal, G = A V = V diag(lambda), and lambda_j = v_j . g_j.
// Every round of the round-robin ordering holds n / 2 DISJOINT column pairs: one CTA per pair (three dot products of
// length n, two axpy-like updates of G and V columns), n - 1 rounds per sweep, ~8-12 sweeps.  Columns are stored
// contiguously ([column][row]; A is symmetric, so its rows serve as its columns).  All arithmetic float64, the pair
// order is fixed: the result is deterministic.
#include <algorithm>

#include "common.cuh"

namespace runia {

__device__ __forceinline__ double block_sum_256(double v, double *red) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];  // fixed order
  return t;
}

// the pair of columns CTA `i` rotates in round `r` (circle method over m = n_even - 1 movable positions)
__device__ __forceinline__ void round_robin_pair(int n_even, int r, int i, int &p, int &q) {
  const int m = n_even - 1;
  if (i == 0) {
    p = m;
    q = r % m;
  } else {
    p = (r + i) % m;
    q = (r - i + m) % m;
  }
}

__global__ void __launch_bounds__(256)
jacobi_round_kernel(double *__restrict__ G, double *__restrict__ V, int n, int n_even, int round, double tol,
                    unsigned long long *__restrict__ off_flag) {
  __shared__ double red[8];
  int p, q;
  round_robin_pair(n_even, round, blockIdx.x, p, q);
  if (p >= n || q >= n) return;  // the dummy column of an odd n
  double *gp = G + (size_t)p * n, *gq = G + (size_t)q * n;
  double a = 0.0, b = 0.0, c = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double x = gp[i], y = gq[i];
    a = fma(x, x, a);
    b = fma(y, y, b);
    c = fma(x, y, c);
  }
  a = block_sum_256(a, red);
  b = block_sum_256(b, red);
  c = block_sum_256(c, red);
  const double lim = tol * sqrt(a * b);
  if (!(fabs(c) > lim) || c == 0.0) return;  // already orthogonal (or a zero column)
  if (threadIdx.x == 0) atomicAdd(off_flag, 1ull);
  const double zeta = (b - a) / (2.0 * c);
  const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
  const double cs = 1.0 / sqrt(1.0 + t * t), sn = cs * t;
  double *vp = V + (size_t)p * n, *vq = V + (size_t)q * n;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double x = gp[i], y = gq[i];
    gp[i] = cs * x - sn * y;
    gq[i] = sn * x + cs * y;
    const double u = vp[i], w = vq[i];
    vp[i] = cs * u - sn * w;
    vq[i] = sn * u + cs * w;
  }
}

__global__ void __launch_bounds__(256) eigh_init_kernel(const double *__restrict__ A, double *__restrict__ G,
                                                        double *__restrict__ V, int n) {
  const int64_t total = (int64_t)n * n;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int col = (int)(e / n), row = (int)(e % n);
    G[e] = 0.5 * (A[(size_t)col * n + row] + A[(size_t)row * n + col]);  // column `col` of the symmetrised matrix
    V[e] = col == row ? 1.0 : 0.0;
  }
}

// lambda_j = v_j . g_j (signed); v_j normalised (it already is, up to rounding)
__global__ void __launch_bounds__(256) eigh_finish_kernel(const double *__restrict__ G, double *__restrict__ V, int n,
                                                          double *__restrict__ evals) {
  __shared__ double red[8];
  const int j = blockIdx.x;
  const double *g = G + (size_t)j * n;
  double *v = V + (size_t)j * n;
  double d = 0.0, nn = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    d = fma(v[i], g[i], d);
    nn = fma(v[i], v[i], nn);
  }
  d = block_sum_256(d, red);
  nn = block_sum_256(nn, red);
  const double inv = 1.0 / sqrt(nn);
  for (int i = threadIdx.x; i < n; i += blockDim.x) v[i] *= inv;
  if (threadIdx.x == 0) evals[j] = d / nn;
}

// Cholesky A = L L^T (lower), one CTA per matrix of a batch, right-looking, in place on a copy; `fail[b]` = first
// column whose pivot is not positive (+1), 0 on success -- the signal gmm_fit's jitter ladder reacts to.
// `rel_pivot` > 0 additionally fails a pivot below rel_pivot * (its original diagonal entry): "numerically singular at
// float32 resolution", the condition under which the reference's float32 factorisation (torch, funcs.py:325-342) gives up.
__global__ void __launch_bounds__(256) cholesky_kernel(const double *__restrict__ A, double *__restrict__ L, int n,
                                                       double jitter, double rel_pivot, int32_t *__restrict__ fail) {
  __shared__ double pivot;
  __shared__ int bad;
  const double *a = A + (size_t)blockIdx.x * n * n;
  double *l = L + (size_t)blockIdx.x * n * n;
  for (int64_t e = threadIdx.x; e < (int64_t)n * n; e += blockDim.x) {
    const int r = (int)(e / n), c = (int)(e % n);
    l[e] = c <= r ? 0.5 * (a[e] + a[(size_t)c * n + r]) + (r == c ? jitter : 0.0) : 0.0;
  }
  if (threadIdx.x == 0) bad = 0;
  __syncthreads();
  for (int k = 0; k < n; ++k) {
    if (threadIdx.x == 0) {
      const double d = l[(size_t)k * n + k];
      const double d0 = a[(size_t)k * n + k] + jitter;
      if (!(d > rel_pivot * d0) || !(d > 0.0) || !isfinite(d)) bad = k + 1;
      pivot = sqrt(d);
      l[(size_t)k * n + k] = pivot;
    }
    __syncthreads();
    if (bad) break;
    const double inv = 1.0 / pivot;
    for (int r = k + 1 + threadIdx.x; r < n; r += blockDim.x) l[(size_t)r * n + k] *= inv;
    __syncthreads();
    // trailing update of the lower triangle: l[r][c] -= l[r][k] * l[c][k] for k < c <= r
    const int m = n - k - 1;
    for (int64_t e = threadIdx.x; e < (int64_t)m * m; e += blockDim.x) {
      const int r = k + 1 + (int)(e / m), c = k + 1 + (int)(e % m);
      if (c <= r) l[(size_t)r * n + c] -= l[(size_t)r * n + k] * l[(size_t)c * n + k];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) fail[blockIdx.x] = bad;
}

// X = L^{-1} for a batch of lower-triangular factors (gmm_prepare's whitening blocks: the reference evaluates the
// mixture through torch's MultivariateNormal, i.e. triangular solves against scale_tril).  Thread j of a CTA owns column j
// of X: forward substitution down the rows, X[i][j] = -(sum_{k=j}^{i-1} L[i][k] X[k][j]) / L[i][i]; the loads of L are
// warp-wide broadcasts, those of X coalesced; the summation order is fixed.
__global__ void __launch_bounds__(128) tril_inverse_kernel(const double *__restrict__ L, double *__restrict__ X, int n) {
  const double *l = L + (size_t)blockIdx.y * n * n;
  double *x = X + (size_t)blockIdx.y * n * n;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  for (int i = 0; i < j; ++i) x[(size_t)i * n + j] = 0.0;
  x[(size_t)j * n + j] = 1.0 / l[(size_t)j * n + j];
  for (int i = j + 1; i < n; ++i) {
    const double *li = l + (size_t)i * n;
    double s0 = 0.0, s1 = 0.0;
    int k = j;
    for (; k + 1 < i; k += 2) {
      s0 = fma(li[k], x[(size_t)k * n + j], s0);
      s1 = fma(li[k + 1], x[(size_t)(k + 1) * n + j], s1);
    }
    if (k < i) s0 = fma(li[k], x[(size_t)k * n + j], s0);
    x[(size_t)i * n + j] = -(s0 + s1) / li[i];
  }
}

}  // namespace runia

using namespace runia;

extern "C" int runia_tril_inverse_f64(const double *L, int batch, int n, double *X, void *stream) {
  RUNIA_NVTX();
  RUNIA_REQUIRE(batch >= 0 && n >= 1, RUNIA_E_BADARG, "tril_inverse: bad sizes");
  if (batch == 0) return RUNIA_OK;
  RUNIA_REQUIRE(L && X && L != X, RUNIA_E_BADARG, "tril_inverse: null or aliased pointer");
  RUNIA_REQUIRE(batch <= 65535, RUNIA_E_UNSUPPORTED, "tril_inverse: batch=%d not supported (max 65535)", batch);
  tril_inverse_kernel<<<dim3((unsigned)ceil_div(n, 128), (unsigned)batch), 128, 0, (cudaStream_t)stream>>>(L, X, n);
  count_launch();
  return finish_launch("tril_inverse");
}

extern "C" size_t runia_eigh_workspace_bytes(int n) {
  if (n < 1) return 0;
  return (size_t)n * n * sizeof(double) + 256;  // G + the rotation counter
}

extern "C" int runia_eigh_f64(const double *A, int n, double *evals, double *evecs, void *workspace, size_t workspace_bytes,
                              int max_sweeps, int *sweeps_out, void *stream) {
  RUNIA_NVTX();
  RUNIA_REQUIRE(n >= 1 && n <= 8192, RUNIA_E_BADARG, "eigh: n=%d outside [1, 8192]", n);
  RUNIA_REQUIRE(A && evals && evecs && workspace, RUNIA_E_BADARG, "eigh: null pointer");
  RUNIA_REQUIRE(workspace_bytes >= runia_eigh_workspace_bytes(n), RUNIA_E_WORKSPACE, "eigh: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  double *G = (double *)workspace;
  unsigned long long *flag = (unsigned long long *)((char *)workspace + (size_t)n * n * sizeof(double));
  double *V = evecs;  // [column][row]: row j of the output buffer is eigenvector j
  const unsigned ig = (unsigned)std::min<int64_t>(ceil_div((int64_t)n * n, 256), (int64_t)kNumSMs * 8);
  eigh_init_kernel<<<ig, 256, 0, st>>>(A, G, V, n);
  count_launch();
  const int n_even = (n + 1) & ~1;
  const int sweeps_cap = max_sweeps > 0 ? max_sweeps : 30;
  int done = 0;
  if (n > 1) {
    for (; done < sweeps_cap; ++done) {
      RUNIA_CUDA(cudaMemsetAsync(flag, 0, sizeof(unsigned long long), st));
      for (int r = 0; r < n_even - 1; ++r)
        jacobi_round_kernel<<<n_even / 2, 256, 0, st>>>(G, V, n, n_even, r, 1e-14, flag);
      count_launch(n_even - 1);
      unsigned long long rotations = 0;  // one small read-back per sweep: the sweep count depends on the matrix
      RUNIA_CUDA(cudaMemcpyAsync(&rotations, flag, sizeof(rotations), cudaMemcpyDeviceToHost, st));
      RUNIA_CUDA(cudaStreamSynchronize(st));
      if (rotations == 0) {
        ++done;
        break;
      }
    }
  }
  eigh_finish_kernel<<<n, 256, 0, st>>>(G, V, n, evals);
  count_launch();
  if (sweeps_out) *sweeps_out = done;
  return finish_launch("eigh");
}

extern "C" int runia_cholesky_f64(const double *A, int batch, int n, double jitter, double rel_pivot, double *L, int32_t *fail,
                                  void *stream) {
  RUNIA_NVTX();
  RUNIA_REQUIRE(batch >= 0 && n >= 1, RUNIA_E_BADARG, "cholesky: bad sizes");
  if (batch == 0) return RUNIA_OK;
  RUNIA_REQUIRE(A && L && fail, RUNIA_E_BADARG, "cholesky: null pointer");
  cholesky_kernel<<<batch, 256, 0, (cudaStream_t)stream>>>(A, L, n, jitter, rel_pivot, fail);
  count_launch();
  return finish_launch("cholesky");
}
