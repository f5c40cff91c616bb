// (a1) MC-dropout latent-sample entropy: evaluation/entropy.py:41-93.
//
// For every item (n_mc consecutive rows of z) and every latent dimension j the reference calls
// the Kozachenko-Leonenko estimator on the n_mc scalars z[:, j] (k-th nearest-neighbour distance
// in 1-D), and once per item on the n_mc D-vectors under the max-norm.  Both are HBM-bound:
// n_mc*D*4 bytes in, D*8 + 8 bytes out per item.
//
// Fast path (n_mc = 16, k = 5 -- the reference's default for n_mc > 5): one warp per item,
// lane l owns dimensions l, l+32, ...  Every global load is a fully coalesced 128-byte row
// segment; the 16 samples of a dimension live in registers, are sorted by a register sorting
// network, and the k-th neighbour distance of each sample is the minimum over the (k+1)-wide
// windows that contain it.  The same registers feed the 120 pairwise |x_i - x_l| running maxima
// of the joint (Chebyshev) estimator, reduced across the warp through a transposed
// shared-memory pass once per item -- so z is read from HBM exactly once.
// Generic path (any 2 <= n_mc <= 32, 1 <= k < n_mc): same numbers, local-memory arrays.
#include "common.cuh"

namespace runia {

constexpr float kLn2 = 0.693147180559945309f;

template <int N>
__device__ __forceinline__ void sort_network(float (&v)[N]) {
  // bitonic network, fully unrolled: all indices are compile-time, v stays in registers
#pragma unroll
  for (int k = 2; k <= N; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
#pragma unroll
      for (int i = 0; i < N; ++i) {
        const int p = i ^ j;
        if (p > i) {
          const float lo = fminf(v[i], v[p]), hi = fmaxf(v[i], v[p]);
          if ((i & k) == 0) {
            v[i] = lo; v[p] = hi;
          } else {
            v[i] = hi; v[p] = lo;
          }
        }
      }
    }
  }
}

// sum_i log2(max(r_i, min_dist)) for the sorted samples s[0..N): r_i = k-th neighbour distance
template <int N, int K>
__device__ __forceinline__ float sum_log2_knn_1d(const float (&s)[N], float min_dist) {
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    float r = INFINITY;
#pragma unroll
    for (int a = 0; a <= N - 1 - K; ++a) {
      if (a <= i && i <= a + K) r = fminf(r, fmaxf(s[i] - s[a], s[a + K] - s[i]));
    }
    acc += __log2f(fmaxf(r, min_dist));
  }
  return acc;
}

template <int N, int K, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
entropy_fast_kernel(const float *__restrict__ z, int64_t n_items, int D, float min_dist, double c_term,
                    double *__restrict__ h_z, double *__restrict__ h_mvn) {
  constexpr int NPAIR = N * (N - 1) / 2;
  extern __shared__ float smem[];  // per warp: [NPAIR][32] transposed pair maxima, then [N][N]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t item = (int64_t)blockIdx.x * WARPS + warp;
  if (item >= n_items) return;
  float *pmx = smem + (size_t)warp * (NPAIR * 32);
  const float *zi = z + item * (int64_t)N * D;
  const bool want_joint = (h_mvn != nullptr);

  float pm[NPAIR];
#pragma unroll
  for (int p = 0; p < NPAIR; ++p) pm[p] = 0.f;

  for (int j0 = 0; j0 < D; j0 += 32) {
    const int j = j0 + lane;
    const bool ok = j < D;
    float v[N];
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = ok ? __ldg(zi + (int64_t)i * D + j) : 0.f;
    if (want_joint) {
      int p = 0;
#pragma unroll
      for (int a = 0; a < N; ++a)
#pragma unroll
        for (int b = a + 1; b < N; ++b) {
          pm[p] = fmaxf(pm[p], fabsf(v[a] - v[b]));
          ++p;
        }
    }
    sort_network<N>(v);
    const float sl = sum_log2_knn_1d<N, K>(v, min_dist);
    // h = -psi(k) + psi(n) + (1/n) sum log(2 r)   [d = 1]
    if (ok) h_z[item * (int64_t)D + j] = c_term + (double)(kLn2 * (1.f + sl * (1.f / N)));
  }
  if (!want_joint) return;

  // transpose-reduce the pair maxima across the 32 lanes
#pragma unroll
  for (int p = 0; p < NPAIR; ++p) pmx[p * 32 + lane] = pm[p];
  __syncwarp();
  float red[(NPAIR + 31) / 32];
#pragma unroll
  for (int q = 0; q < (NPAIR + 31) / 32; ++q) {
    const int p = q * 32 + lane;
    float m = 0.f;
    if (p < NPAIR)
      for (int t = 0; t < 32; ++t) m = fmaxf(m, pmx[p * 32 + ((t + lane) & 31)]);
    red[q] = m;
  }
  __syncwarp();
  float *dm = pmx;  // reuse: dense [N][N] Chebyshev matrix
  for (int e = lane; e < N * N; e += 32) dm[e] = 0.f;
  __syncwarp();
#pragma unroll
  for (int q = 0; q < (NPAIR + 31) / 32; ++q) {
    const int p = q * 32 + lane;
    if (p < NPAIR) {
      // invert p -> (a, b), a < b, row-major upper triangle
      int a = 0, rem = p;
      while (rem >= N - 1 - a) {
        rem -= N - 1 - a;
        ++a;
      }
      const int b = a + 1 + rem;
      dm[a * N + b] = red[q];
      dm[b * N + a] = red[q];
    }
  }
  __syncwarp();
  float lg = 0.f;
  if (lane < N) {
    float row[N];
#pragma unroll
    for (int b = 0; b < N; ++b) row[b] = dm[lane * N + b];
    sort_network<N>(row);  // row[0] = 0 (self); row[K] = k-th neighbour
    lg = __log2f(fmaxf(row[K], min_dist));
  }
  lg = warp_sum32(lg);
  if (lane == 0) h_mvn[item] = c_term + (double)D * (double)(kLn2 * (1.f + lg * (1.f / N)));
}

// ---------------------------------- generic path ------------------------------------------
__global__ void __launch_bounds__(128)
entropy_generic_dim_kernel(const float *__restrict__ z, int64_t n_items, int n, int D, int k, float min_dist,
                           double c_term, double *__restrict__ h_z) {
  const int64_t total = n_items * (int64_t)D;
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const int64_t item = e / D;
  const int j = (int)(e % D);
  const float *zi = z + item * (int64_t)n * D + j;
  float s[32];
  for (int i = 0; i < n; ++i) {  // insertion sort
    const float x = zi[(int64_t)i * D];
    int p = i;
    while (p > 0 && s[p - 1] > x) {
      s[p] = s[p - 1];
      --p;
    }
    s[p] = x;
  }
  float acc = 0.f;
  for (int i = 0; i < n; ++i) {
    float r = INFINITY;
    const int a_lo = i - k > 0 ? i - k : 0;
    const int a_hi = i < n - 1 - k ? i : n - 1 - k;
    for (int a = a_lo; a <= a_hi; ++a) r = fminf(r, fmaxf(s[i] - s[a], s[a + k] - s[i]));
    acc += __log2f(fmaxf(r, min_dist));
  }
  h_z[e] = c_term + (double)(kLn2 * (1.f + acc / (float)n));
}

__global__ void __launch_bounds__(128)
entropy_generic_joint_kernel(const float *__restrict__ z, int64_t n_items, int n, int D, int k, float min_dist,
                             double c_term, double *__restrict__ h_mvn) {
  __shared__ float dm_all[4][32 * 33];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t item = (int64_t)blockIdx.x * 4 + warp;
  if (item >= n_items) return;
  float *dm = dm_all[warp];
  const float *zi = z + item * (int64_t)n * D;
  for (int a = 0; a < n; ++a)
    for (int b = a + 1; b < n; ++b) {
      float m = 0.f;
      for (int j = lane; j < D; j += 32) m = fmaxf(m, fabsf(zi[(int64_t)a * D + j] - zi[(int64_t)b * D + j]));
      m = warp_max32(m);
      if (lane == 0) {
        dm[a * 33 + b] = m;
        dm[b * 33 + a] = m;
      }
    }
  if (lane < n) dm[lane * 33 + lane] = 0.f;
  __syncwarp();
  float lg = 0.f;
  if (lane < n) {
    // (k+1)-th smallest of the row including self: element with exactly k predecessors under
    // the total order (value, index)
    float r = 0.f;
    for (int b = 0; b < n; ++b) {
      const float vb = dm[lane * 33 + b];
      int rank = 0;
      for (int c = 0; c < n; ++c) {
        const float vc = dm[lane * 33 + c];
        rank += (vc < vb || (vc == vb && c < b)) ? 1 : 0;
      }
      if (rank == k) r = vb;
    }
    lg = __log2f(fmaxf(r, min_dist));
  }
  lg = warp_sum32(lg);
  if (lane == 0) h_mvn[item] = c_term + (double)D * (double)(kLn2 * (1.f + lg / (float)n));
}

}  // namespace runia

using namespace runia;

extern "C" int runia_mcd_entropy_f32(const float *z, int64_t n_items, int n_mc, int D, int k, double min_dist,
                                     double digamma_term, double *h_z, double *h_mvn, void *stream) {
  RUNIA_REQUIRE(n_items >= 0 && D > 0, RUNIA_E_BADARG, "mcd_entropy: bad sizes n_items=%lld D=%d",
                (long long)n_items, D);
  RUNIA_REQUIRE(n_mc >= 2 && n_mc <= 32, RUNIA_E_UNSUPPORTED, "mcd_entropy: n_mc=%d outside [2, 32]", n_mc);
  RUNIA_REQUIRE(k >= 1 && k < n_mc, RUNIA_E_BADARG, "mcd_entropy: k=%d must satisfy 1 <= k < n_mc=%d", k, n_mc);
  if (n_items == 0) return RUNIA_OK;
  RUNIA_REQUIRE(z && h_z, RUNIA_E_BADARG, "mcd_entropy: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if (n_mc == 16 && k == 5) {
    constexpr int WARPS = 4;
    constexpr size_t smem = (size_t)WARPS * 120 * 32 * sizeof(float);
    static bool attr = false;
    if (!attr) {
      RUNIA_CUDA(cudaFuncSetAttribute(entropy_fast_kernel<16, 5, WARPS>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr = true;
    }
    entropy_fast_kernel<16, 5, WARPS><<<(unsigned)ceil_div(n_items, WARPS), WARPS * 32, smem, st>>>(
        z, n_items, D, (float)min_dist, digamma_term, h_z, h_mvn);
    count_launch();
    return finish_launch("mcd_entropy(fast)");
  }
  const int64_t total = n_items * (int64_t)D;
  entropy_generic_dim_kernel<<<(unsigned)ceil_div(total, 128), 128, 0, st>>>(z, n_items, n_mc, D, k, (float)min_dist,
                                                                            digamma_term, h_z);
  count_launch();
  if (h_mvn) {
    entropy_generic_joint_kernel<<<(unsigned)ceil_div(n_items, 4), 128, 0, st>>>(z, n_items, n_mc, D, k,
                                                                                (float)min_dist, digamma_term, h_mvn);
    count_launch();
  }
  return finish_launch("mcd_entropy(generic)");
}
