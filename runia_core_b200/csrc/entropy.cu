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
#include <cuda.h>

#include <algorithm>

#include "common.cuh"

namespace runia {

namespace tc {
int make_plain_map(CUtensorMap *map, const float *ptr, int64_t rows, int K, int box_rows, int box_cols);
}

constexpr float kLn2 = 0.693147180559945309f;

template <int N>
__device__ __forceinline__ void sort_network(float (&v)[N]) {
  // Batcher's odd-even merge sort (N a power of two), fully unrolled: every index is a compile-time constant, v stays
  // in registers.  19 / 63 / 191 comparators for N = 8 / 16 / 32 (the bitonic network needs 24 / 80 / 240).
#pragma unroll
  for (int p = 1; p < N; p <<= 1) {
#pragma unroll
    for (int k = p; k >= 1; k >>= 1) {
#pragma unroll
      for (int a = 0; a < N; ++a) {  // comparator (a, a + k) of this layer, if any
        const int off = k % p;
        if (a >= off && ((a - off) % (2 * k)) < k && a + k < N && (a / (2 * p)) == ((a + k) / (2 * p))) {
          const float lo = fminf(v[a], v[a + k]), hi = fmaxf(v[a], v[a + k]);
          v[a] = lo;
          v[a + k] = hi;
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

// ------------------------------ n_mc = 16, k = 5, D % 4 == 0 ---------------------------------
// The estimator is bound by the min/max (ALU) pipe, not by HBM: per (item, dimension) it needs
// 120 pair maxima, a 16-element sort and 66 window radii, and every FMNMX / FMNMX3 holds the half-rate
// ALU pipe for two cycles (scripts/probes/pipe_probe.cu -> profiles/r2c_pipe_probe.jsonl).  This kernel
// spends as few ALU-pipe instructions on them as the ISA allows:
//  * a lane owns TWO adjacent dimensions per step; the 120 pair differences of both are one packed
//    FADD2 each (FMA pipe) and fold into the running Chebyshev maxima with one 3-input FMNMX3
//    max(pm, |d.x|, |d.y|);
//  * the sort is the 60-comparator, 10-layer network (optimal size for 16 keys);
//  * the k-th neighbour distance of sample i is min over the windows [a, a+5] containing it of
//    max(s_i - s_a, s_{a+5} - s_i); the two end points of a window need no max at all, the four
//    interior points take a two-input max, the min over (up to) six windows is a chain of 3-input
//    mins and the min_dist clamp is applied once per point; the differences are packed FADD2;
//  * log2 is the bare MUFU (the radii are >= min_dist > 0: no denormal fix-up).
// z is streamed HBM -> shared memory with 16-byte cp.async into a per-warp 3-deep ring (one step =
// 16 samples x 64 dimensions = 4 KB), two steps ahead of the arithmetic, so that no register is spent
// on prefetch (the 120 maxima + 32 samples already need ~250) and HBM latency never reaches a warp.
// Warps are persistent over items; the ring runs across item boundaries.
__device__ __forceinline__ float2 sub2(float2 a, float2 b) {
  float2 r;
  asm("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; sub.f32x2 rc, ra, rb; mov.b64 {%0, %1}, rc;}"
      : "=f"(r.x), "=f"(r.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return r;
}
__device__ __forceinline__ float lg2_pos(float x) {  // x >= min_dist > 0, normal
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float warp_max_f32(float v) {  // CREDUX on sm_100a
  float r;
  asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));
  return r;
}
#define RUNIA_CE(a, b)                   \
  {                                      \
    const float lo_ = fminf(v[a], v[b]); \
    const float hi_ = fmaxf(v[a], v[b]); \
    v[a] = lo_;                          \
    v[b] = hi_;                          \
  }
// 60 compare-exchanges in 10 layers; tests/test_cabi_and_host.py checks the comparator list with the 0-1 principle
__device__ __forceinline__ void sort16(float (&v)[16]) {
  RUNIA_CE(0, 13) RUNIA_CE(1, 12) RUNIA_CE(2, 15) RUNIA_CE(3, 14) RUNIA_CE(4, 8) RUNIA_CE(5, 6) RUNIA_CE(7, 11) RUNIA_CE(9, 10)
  RUNIA_CE(0, 5) RUNIA_CE(1, 7) RUNIA_CE(2, 9) RUNIA_CE(3, 4) RUNIA_CE(6, 13) RUNIA_CE(8, 14) RUNIA_CE(10, 15) RUNIA_CE(11, 12)
  RUNIA_CE(0, 1) RUNIA_CE(2, 3) RUNIA_CE(4, 5) RUNIA_CE(6, 8) RUNIA_CE(7, 9) RUNIA_CE(10, 11) RUNIA_CE(12, 13) RUNIA_CE(14, 15)
  RUNIA_CE(0, 2) RUNIA_CE(1, 3) RUNIA_CE(4, 10) RUNIA_CE(5, 11) RUNIA_CE(6, 7) RUNIA_CE(8, 9) RUNIA_CE(12, 14) RUNIA_CE(13, 15)
  RUNIA_CE(1, 2) RUNIA_CE(3, 12) RUNIA_CE(4, 6) RUNIA_CE(5, 7) RUNIA_CE(8, 10) RUNIA_CE(9, 11) RUNIA_CE(13, 14)
  RUNIA_CE(1, 4) RUNIA_CE(2, 6) RUNIA_CE(5, 8) RUNIA_CE(7, 10) RUNIA_CE(9, 13) RUNIA_CE(11, 14)
  RUNIA_CE(2, 4) RUNIA_CE(3, 6) RUNIA_CE(9, 12) RUNIA_CE(11, 13)
  RUNIA_CE(3, 5) RUNIA_CE(6, 8) RUNIA_CE(7, 9) RUNIA_CE(10, 12)
  RUNIA_CE(3, 4) RUNIA_CE(5, 6) RUNIA_CE(7, 8) RUNIA_CE(9, 10) RUNIA_CE(11, 12)
  RUNIA_CE(6, 7) RUNIA_CE(8, 9)
}
#undef RUNIA_CE

constexpr int E16_WARPS = 4;
constexpr int E16_RING = 2;
constexpr int E16_STEP_FLOATS = 16 * 64;                                    // one step: 16 samples x 64 dims
constexpr int E16_NREG = 60;                                                // pair maxima kept in registers
constexpr int E16_NSM = 120 - E16_NREG;                                     // pair maxima kept in shared memory ([q][lane] float4)
constexpr int E16_WARP_FLOATS = E16_RING * E16_STEP_FLOATS + 16 * 16 + E16_NSM * 32;  // ring + Chebyshev matrix + maxima
constexpr size_t kEntropy16Smem = (size_t)E16_WARPS * E16_WARP_FLOATS * sizeof(float) + E16_WARPS * E16_RING * 8;  // + mbarriers

// (a, b) of pair p in the row-major upper triangle of the 16 x 16 matrix
__constant__ uint8_t kPairA16[120], kPairB16[120];

__device__ __forceinline__ uint32_t e16_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void e16_mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "E16_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra E16_DONE;\n\t"
      "bra E16_WAIT;\n\t"
      "E16_DONE:\n\t"
      "}" ::"r"(bar),
      "r"(parity)
      : "memory");
}

// z is streamed by the TMA unit: one elected lane per warp issues ONE cp.async.bulk.tensor per step (a 16 x 64 box of
// the [rows, D] matrix, columns past D zero-filled) into the warp's 2-slot ring, completion on a per-slot mbarrier;
// the warp's instruction stream carries no address arithmetic and no per-lane copies for the loads any more
// (8 LDGSTS + ~60 integer instructions per step before).
__global__ void __launch_bounds__(E16_WARPS * 32, 3)
entropy16_kernel(const __grid_constant__ CUtensorMap tmZ, int64_t n_items, int D, float min_dist, double c_term,
                 double *__restrict__ h_z, double *__restrict__ h_mvn) {
  constexpr int N = 16, K = 5;
  extern __shared__ __align__(128) float smem16[];  // per warp: ring | Chebyshev matrix | maxima; then the mbarriers
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float *ring = smem16 + (size_t)warp * E16_WARP_FLOATS;
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem16 + (size_t)E16_WARPS * E16_WARP_FLOATS) + warp * E16_RING;
  float *dm = ring + E16_RING * E16_STEP_FLOATS;
  float *pmsf = dm + 16 * 16;                            // [q][lane][4]: shared-memory half of the pair maxima
  float4 *pms = reinterpret_cast<float4 *>(pmsf) + lane;  // this lane's maxima: pms[32 * q], q < E16_NSM / 4
  const uint32_t ring_u32 = e16_smem_u32(ring);
  const uint32_t bar_u32 = e16_smem_u32(bars);
  const int64_t gw = (int64_t)blockIdx.x * E16_WARPS + warp;  // this warp's first item
  const int64_t GW = (int64_t)gridDim.x * E16_WARPS;          // item stride
  const int spi = (D + 63) >> 6;                              // steps per item
  const int64_t n_my = gw < n_items ? (n_items - gw + GW - 1) / GW : 0;
  const int64_t n_steps = n_my * spi;

  if (lane == 0) {
    for (int s = 0; s < E16_RING; ++s)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_u32 + 8u * s) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmZ) : "memory");
  }
  __syncwarp();

  // copy stream (runs E16_RING - 1 steps ahead of the compute stream), driven by lane 0
  int64_t c_item = gw;
  int c_j = 0, c_buf = 0;
  int64_t c_step = 0;
  auto issue = [&]() {
    if (c_step < n_steps) {
      if (lane == 0) {
        const uint32_t bar = bar_u32 + 8u * (uint32_t)c_buf;
        const uint32_t dst = ring_u32 + (uint32_t)(c_buf * E16_STEP_FLOATS) * 4u;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the slot's last readers (generic proxy) are done
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(E16_STEP_FLOATS * 4) : "memory");
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
            "l"(&tmZ), "r"(bar), "r"(c_j * 64), "r"((int)(c_item * N))
            : "memory");
      }
      if (++c_j == spi) {
        c_j = 0;
        c_item += GW;
      }
      if (++c_buf == E16_RING) c_buf = 0;
      ++c_step;
    }
  };
  issue();
  issue();  // both slots in flight: the slot a step has just read is refilled with the step TWO ahead

  int buf = 0;
  uint32_t phase = 0;
  for (int64_t item = gw; item < n_items; item += GW) {
    // half of the 120 maxima live in registers, half in shared memory: 168 registers -> 12 resident warps
    float pm[E16_NREG];
#pragma unroll
    for (int p = 0; p < E16_NREG; ++p) pm[p] = 0.f;
#pragma unroll
    for (int q = 0; q < E16_NSM / 4; ++q) pms[32 * q] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
    for (int jstep = 0; jstep < spi; ++jstep) {
      e16_mbar_wait(bar_u32 + 8u * (uint32_t)buf, phase);
      float2 x[N];
      {
        const float2 *b2 = reinterpret_cast<const float2 *>(ring + buf * E16_STEP_FLOATS) + lane;
#pragma unroll
        for (int i = 0; i < N; ++i) x[i] = b2[i * 32];
      }
      __syncwarp();  // every lane has read the slot that the next issue() refills (the one read a step ago)
      issue();
      if (++buf == E16_RING) {
        buf = 0;
        phase ^= 1u;
      }
      const int j = jstep * 32 + lane;  // float2 column: dimensions 2j, 2j+1
      {
        int p = 0;
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int a = 0; a < N; ++a)
#pragma unroll
          for (int b = a + 1; b < N; ++b) {
            const float2 d = sub2(x[a], x[b]);
            if (p < E16_NREG) {
              pm[p] = fmaxf(fmaxf(pm[p], fabsf(d.x)), fabsf(d.y));
            } else {
              const int q = (p - E16_NREG) >> 2, c = (p - E16_NREG) & 3;
              if (c == 0) t = pms[32 * q];
              if (c == 0) t.x = fmaxf(fmaxf(t.x, fabsf(d.x)), fabsf(d.y));
              if (c == 1) t.y = fmaxf(fmaxf(t.y, fabsf(d.x)), fabsf(d.y));
              if (c == 2) t.z = fmaxf(fmaxf(t.z, fabsf(d.x)), fabsf(d.y));
              if (c == 3) t.w = fmaxf(fmaxf(t.w, fabsf(d.x)), fabsf(d.y));
              if (c == 3) pms[32 * q] = t;
            }
            ++p;
          }
      }
      float2 s[N];
      {
        float v[N];
#pragma unroll
        for (int i = 0; i < N; ++i) v[i] = x[i].x;
        sort16(v);
#pragma unroll
        for (int i = 0; i < N; ++i) s[i].x = v[i];
#pragma unroll
        for (int i = 0; i < N; ++i) v[i] = x[i].y;
        sort16(v);
#pragma unroll
        for (int i = 0; i < N; ++i) s[i].y = v[i];
      }
      float ax = 0.f, ay = 0.f;
      {
        // Two-input maxima for the interior points and ONE min_dist clamp per point, after the minimum over the windows.
        // Every min / max instruction, 2- or 3-input, holds the half-rate ALU pipe for two cycles
        // (profiles/r2c_pipe_probe.jsonl), so a 3-input form only pays when all three inputs are needed: the clamp
        // folded into each candidate cost an instruction per candidate (44 per dimension) against 16 now.  Forming the
        // candidates with adds only (w_a + |L - R| = 2 max(L, R)) and scalar instead of packed differences were both
        // measured slower (profiles/r2c_entropy_variants.json).
        float2 wc[N - K];
#pragma unroll
        for (int a = 0; a < N - K; ++a) wc[a] = sub2(s[a + K], s[a]);
#pragma unroll
        for (int i = 0; i < N; ++i) {
          float cx[K + 1], cy[K + 1];
          int nc = 0;
#pragma unroll
          for (int a = 0; a < N - K; ++a) {
            if (a <= i && i <= a + K) {
              if (i == a || i == a + K) {
                cx[nc] = wc[a].x;
                cy[nc] = wc[a].y;
              } else {
                const float2 L = sub2(s[i], s[a]);
                const float2 R = sub2(s[a + K], s[i]);
                cx[nc] = fmaxf(L.x, R.x);
                cy[nc] = fmaxf(L.y, R.y);
              }
              ++nc;
            }
          }
          float rx = cx[0], ry = cy[0];
          if (nc == 2) rx = fminf(rx, cx[1]), ry = fminf(ry, cy[1]);
          if (nc >= 3) rx = fminf(fminf(rx, cx[1]), cx[2]), ry = fminf(fminf(ry, cy[1]), cy[2]);
          if (nc == 4) rx = fminf(rx, cx[3]), ry = fminf(ry, cy[3]);
          if (nc >= 5) rx = fminf(fminf(rx, cx[3]), cx[4]), ry = fminf(fminf(ry, cy[3]), cy[4]);
          if (nc == 6) rx = fminf(rx, cx[5]), ry = fminf(ry, cy[5]);
          ax += lg2_pos(fmaxf(rx, min_dist));
          ay += lg2_pos(fmaxf(ry, min_dist));
        }
      }
      if (2 * j < D) {
        double2 o;  // h = -psi(k) + psi(n) + (1/n) sum log(2 r)   [d = 1]
        o.x = c_term + (double)(kLn2 * (1.f + ax * (1.f / N)));
        o.y = c_term + (double)(kLn2 * (1.f + ay * (1.f / N)));
        reinterpret_cast<double2 *>(h_z + item * (int64_t)D)[j] = o;
      }
    }
    // ---- item complete: joint (Chebyshev) estimator from the 120 pair maxima ----
    // The per-lane maxima are reduced across the lanes through shared memory, 60 pairs at a time in the [q][lane][4]
    // block the shared-memory half already lives in: lane L takes pairs L and L + 32 of the block and walks the 32
    // lanes' values of each (rotated by L >> 2 so that the 32 lanes hit 32 banks).
    if (h_mvn != nullptr) {
      __syncwarp();
#pragma unroll 1
      for (int half = 0; half < 2; ++half) {
        if (half == 1) {  // second half: the register-resident maxima take the block's place
          __syncwarp();
#pragma unroll
          for (int q = 0; q < E16_NREG / 4; ++q) pms[32 * q] = make_float4(pm[4 * q], pm[4 * q + 1], pm[4 * q + 2], pm[4 * q + 3]);
          __syncwarp();
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const int pl = lane + 32 * r;  // pair inside the block
          if (pl < E16_NSM) {
            const float *col = pmsf + (pl >> 2) * 128 + (pl & 3);
            const int rot = lane >> 2;
            float m0 = 0.f, m1 = 0.f;
#pragma unroll
            for (int t = 0; t < 32; t += 2) {
              m0 = fmaxf(m0, col[((t + rot) & 31) * 4]);
              m1 = fmaxf(m1, col[((t + 1 + rot) & 31) * 4]);
            }
            const int pg = half == 0 ? E16_NREG + pl : pl;  // global pair index
            const int a = kPairA16[pg], b = kPairB16[pg];
            const float m = fmaxf(m0, m1);
            dm[a * N + b] = m;
            dm[b * N + a] = m;
          }
        }
      }
      if (lane < N) dm[lane * N + lane] = 0.f;
      __syncwarp();
      float lg = 0.f;
      if (lane < N) {
        float v[N];
#pragma unroll
        for (int b = 0; b < N; ++b) v[b] = dm[lane * N + b];
        sort16(v);  // v[0] = 0 (self); v[K] = k-th neighbour
        lg = lg2_pos(fmaxf(v[K], min_dist));
      }
      lg = warp_sum32(lg);
      if (lane == 0) h_mvn[item] = c_term + (double)D * (double)(kLn2 * (1.f + lg * (1.f / N)));
      __syncwarp();  // dm / pms are rewritten by the next item
    }
  }
}

// ------------------------- any n_mc in [6, 32], k = 5 (the reference's default is 32) -------------------------
// One warp per item, 32 dimensions per step.  NP = n_mc rounded up to a power of two; missing samples are +inf
// sentinels (they sort to the end and every window that touches one has an infinite radius, so they drop out of
// the minima by themselves).
//  * per-dimension part: lane = dimension, the n samples of the lane's dimension in registers (coalesced loads,
//    one 128-byte line per sample), bitonic network, window minima as above;
//  * joint (Chebyshev) part: lane = SAMPLE.  The step's [n][32] block is staged in shared memory (row stride 36);
//    lane i keeps its own 32 values in registers and walks the rows j < n as broadcast LDS.128, folding
//    |x_i - x_j| into acc_j with FADD2 + 3-input max -- the whole row i of the distance matrix stays in lane i's
//    registers (32 accumulators instead of n (n - 1) / 2 per lane), and the k-th neighbour of sample i is read off
//    a sort of that row at the end.
constexpr int ENP_WARPS = 4;
constexpr int ENP_STRIDE = 36;  // floats per staged row: 16-byte aligned, rows land in different bank groups

template <int NP>
__global__ void __launch_bounds__(ENP_WARPS * 32, 4)
entropy_np_kernel(const float *__restrict__ z, int64_t n_items, int n, int D, float min_dist, double c_term,
                  double *__restrict__ h_z, double *__restrict__ h_mvn) {
  constexpr int K = 5;
  __shared__ __align__(16) float stage_all[ENP_WARPS][NP * ENP_STRIDE];
  __shared__ float acc_all[ENP_WARPS][(NP / 2 + 1) * 32];  // [t][lane]: running max_d |x_lane[d] - x_(lane + t)[d]|
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float *stage = stage_all[warp];
  float *acc = acc_all[warp] + lane;
  const bool want_joint = h_mvn != nullptr;
  for (int64_t item = (int64_t)blockIdx.x * ENP_WARPS + warp; item < n_items; item += (int64_t)gridDim.x * ENP_WARPS) {
    const float *zi = z + item * (int64_t)n * D;
    for (int jj = 0; jj <= (n >> 1); ++jj) acc[jj * 32] = 0.f;  // slot t: pair (lane, lane + t), t = 1 .. n / 2
    for (int j0 = 0; j0 < D; j0 += 32) {
      const int j = j0 + lane;
      const bool ok = j < D;
      float v[NP];
#pragma unroll
      for (int i = 0; i < NP; ++i) v[i] = i < n ? (ok ? __ldg(zi + (int64_t)i * D + j) : 0.f) : INFINITY;
      if (want_joint) {
        __syncwarp();  // the previous step's readers are done with the stage
#pragma unroll
        for (int i = 0; i < NP; ++i)
          if (i < n) stage[i * ENP_STRIDE + lane] = v[i];
      }
      // ---- per-dimension estimator (lane = dimension) ----
      sort_network<NP>(v);
      float wd[NP - K];  // window widths, clamped: the radius of a window for its two end points (no max needed there)
#pragma unroll
      for (int a = 0; a + K < NP; ++a) wd[a] = fmaxf(v[a + K] - v[a], min_dist);
      float sl = 0.f;
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        // candidates of sample i: one per window [a, a + K] containing it; interior points fold the min_dist clamp into a
        // 3-input max, and the minimum over the (up to six) candidates is a chain of 3-input mins
        float cand[K + 1];
        int nc = 0;
#pragma unroll
        for (int a = 0; a + K < NP; ++a) {
          if (a <= i && i <= a + K) {
            if (i == a || i == a + K)
              cand[nc] = wd[a];
            else
              cand[nc] = fmaxf(fmaxf(v[i] - v[a], v[a + K] - v[i]), min_dist);
            ++nc;
          }
        }
        float r = cand[0];
        if (nc == 2) r = fminf(r, cand[1]);
        if (nc >= 3) r = fminf(fminf(r, cand[1]), cand[2]);
        if (nc == 4) r = fminf(r, cand[3]);
        if (nc >= 5) r = fminf(fminf(r, cand[3]), cand[4]);
        if (nc == 6) r = fminf(r, cand[5]);
        sl += i < n ? lg2_pos(r) : 0.f;
      }
      if (ok) h_z[item * (int64_t)D + j] = c_term + (double)(kLn2 * (1.f + sl / (float)n));
      // ---- joint estimator: fold this step into row `lane` of the Chebyshev matrix (lane = sample) ----
      if (want_joint) {
        __syncwarp();
        float4 own[8];
        const float *mine = stage + (lane < n ? lane : 0) * ENP_STRIDE;
#pragma unroll
        for (int q = 0; q < 8; ++q) own[q] = *reinterpret_cast<const float4 *>(mine + 4 * q);
        // a loop over the other samples, FOUR rows per trip (the accumulators live in shared memory for that).  Fully
        // unrolled (32 rows) this body was 1,300 instructions and the kernel ran out of instruction cache
        // (stall_no_instruction 1.2); one row per trip left a single dependent chain of 16 three-input maxima per row
        // (plus the LDS / STS of its accumulator) with four warps per scheduler to hide it: 5,300 cycles per step where
        // the ALU pipe needs 2,300.  Four independent chains per trip fill the pipe.
        auto fold_row = [&](const float *row, float m) {
          float ma = m, mb = 0.f;  // two chains per row, merged at the end
#pragma unroll
          for (int q = 0; q < 8; q += 2) {
            const float4 o0 = *reinterpret_cast<const float4 *>(row + 4 * q);  // broadcast
            const float4 o1 = *reinterpret_cast<const float4 *>(row + 4 * q + 4);
            const float2 d0 = sub2(make_float2(own[q].x, own[q].y), make_float2(o0.x, o0.y));
            const float2 d1 = sub2(make_float2(own[q].z, own[q].w), make_float2(o0.z, o0.w));
            const float2 d2 = sub2(make_float2(own[q + 1].x, own[q + 1].y), make_float2(o1.x, o1.y));
            const float2 d3 = sub2(make_float2(own[q + 1].z, own[q + 1].w), make_float2(o1.z, o1.w));
            ma = fmaxf(fmaxf(ma, fabsf(d0.x)), fabsf(d0.y));  // one 3-input FMNMX3 each (ptxas fuses this nesting only)
            ma = fmaxf(fmaxf(ma, fabsf(d1.x)), fabsf(d1.y));
            mb = fmaxf(fmaxf(mb, fabsf(d2.x)), fabsf(d2.y));
            mb = fmaxf(fmaxf(mb, fabsf(d3.x)), fabsf(d3.y));
          }
          return fmaxf(ma, mb);
        };
        // every unordered pair once: lane i takes the rows i + 1 .. i + n / 2 (mod n); accumulator slot t of lane i
        // holds the distance of the pair (i, i + t).  (For even n the pairs at t = n / 2 are evaluated from both ends.)
        // The row loads are no longer broadcasts -- 32 lanes read 32 different rows, four shared-memory wavefronts per
        // LDS.128 with the 36-float row stride -- but the table's 2,048 lane-ops per dimension become 1,024.
        const int T = n >> 1;
        auto row_of = [&](int t) {
          int r = lane + t;
          r = r >= n ? r - n : r;
          return stage + (lane < n ? r : 0) * ENP_STRIDE;
        };
        int t = 1;
#pragma unroll 1
        for (; t + 3 <= T; t += 4) {
          const float m0 = acc[(t + 0) * 32], m1 = acc[(t + 1) * 32], m2 = acc[(t + 2) * 32], m3 = acc[(t + 3) * 32];
          const float r0 = fold_row(row_of(t + 0), m0);
          const float r1 = fold_row(row_of(t + 1), m1);
          const float r2 = fold_row(row_of(t + 2), m2);
          const float r3 = fold_row(row_of(t + 3), m3);
          acc[(t + 0) * 32] = r0;
          acc[(t + 1) * 32] = r1;
          acc[(t + 2) * 32] = r2;
          acc[(t + 3) * 32] = r3;
        }
#pragma unroll 1
        for (; t <= T; ++t) acc[t * 32] = fold_row(row_of(t), acc[t * 32]);
      }
    }
    if (want_joint) {
      __syncwarp();  // every lane's accumulators are in shared memory: row i also needs the pairs (i - t, i) of other lanes
      float lg = 0.f;
      if (lane < n) {
        const int T = n >> 1;
        const bool even = (n & 1) == 0;
        float row[NP];
        row[0] = 0.f;  // self
#pragma unroll
        for (int t = 1; t <= NP / 2; ++t) {
          int j2 = lane - t;
          j2 = j2 < 0 ? j2 + n : j2;
          const bool use = t <= T && t < n;
          const float v1 = use ? acc_all[warp][t * 32 + lane] : INFINITY;                      // pair (i, i + t)
          const float v2 = (use && !(even && t == T)) ? acc_all[warp][t * 32 + j2] : INFINITY;  // pair (i - t, i)
          row[2 * t - 1] = v1;
          if (2 * t < NP) row[2 * t] = v2;
        }
        sort_network<NP>(row);  // row[0] = 0 (self); row[K] = k-th neighbour
        lg = lg2_pos(fmaxf(row[K], min_dist));
      }
      lg = warp_sum32(lg);
      if (lane == 0) h_mvn[item] = c_term + (double)D * (double)(kLn2 * (1.f + lg / (float)n));
    }
  }
}

template <int NP>
static void launch_entropy_np(const float *z, int64_t n_items, int n, int D, float min_dist, double c_term, double *h_z,
                              double *h_mvn, cudaStream_t st) {
  int per_sm = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, entropy_np_kernel<NP>, ENP_WARPS * 32, 0);
  const unsigned grid = (unsigned)std::min<int64_t>(ceil_div(n_items, ENP_WARPS), (int64_t)kNumSMs * std::max(per_sm, 1));
  entropy_np_kernel<NP><<<grid, ENP_WARPS * 32, 0, st>>>(z, n_items, n, D, min_dist, c_term, h_z, h_mvn);
}

// ------------------------------ n_mc = 32, k = 5 (the reference's default mcd_samples_nro) -------------------------
// 496 pair maxima do not fit one lane's registers, so FOUR warps share an item: the CTA streams 32 x 128 tiles of z
// (one cp.async.bulk.tensor per tile into a 4-slot ring, three tiles ahead, full / empty mbarriers per slot) and every
// warp does a quarter of each part on the tile, all operands read from shared memory without bank conflicts:
//  * joint (Chebyshev) part: the pairs are cut cyclically -- warp w holds the samples i = 8w .. 8w + 7 in registers and
//    streams the samples i + 1 .. i + 16 (mod 32) past them: 8 x 16 = 128 running maxima per lane, a lane owning two
//    adjacent dimensions per sub-step (FADD2 + max(pm, |d.x|, |d.y|)), two sub-steps of 64 dimensions per tile.
//    4 x 128 = 512 slots cover the 496 pairs (the 16 pairs at cyclic distance 16 are evaluated from both ends);
//  * per-dimension part: warp w takes the dimensions 32w .. 32w + 31 of the tile, lane = dimension: 32 keys in
//    registers, Batcher's 191-comparator network, window minima;
//  * item end: the 128 maxima are reduced over the lanes (redux.sync.max.f32), scattered into a 32 x 32 table
//    (eight of them), and every fourth item the CTA meets at a barrier and each warp sorts the rows of one table.
// Per tile and warp: 256 FADD2 + 256 FMNMX3 (pairs), 382 FMNMX (sort), ~220 min / max + ~180 FADD (windows) -- the same
// arithmetic as entropy_np_kernel<32>, but no operand is read through shared-memory bank conflicts, the load stream is
// the TMA unit's, and the two parts of a step no longer alternate between a lane = dimension and a lane = sample layout.
// The bound is the half-rate ALU pipe (every min / max instruction holds it for two cycles: scripts/probes/
// pipe_probe.cu); 30k x 32 x 512: 1.77 -> 1.24 ms.
constexpr int E32_WARPS = 4;
constexpr int E32_RING = 4;
constexpr int E32_COLS = 128;
constexpr int E32_SLOT_FLOATS = 32 * E32_COLS;
constexpr int E32_DM = 32 * 33;
constexpr int E32_NDM = 8;  // Chebyshev tables: two groups of four (see the item end)
constexpr size_t kEntropy32Smem = (size_t)(E32_RING * E32_SLOT_FLOATS + E32_NDM * E32_DM) * sizeof(float) + 2 * E32_RING * 8;

__device__ __forceinline__ void e32_mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.u32 %0, 1, 0, P1;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}

// sum_i log2(max(r_i, min_dist)) over the sorted keys, k = 5: r_i = min over the windows [a, a + 5] holding i of
// max(s_i - s_a, s_(a+5) - s_i); two-input maxima and one clamp per point, after the minimum.
template <int N, bool FULL = true>
__device__ __forceinline__ float sum_log2_knn5_sorted(const float (&s)[N], float min_dist, int n = N) {
  constexpr int K = 5;  // !FULL: only the first n keys are samples, the rest +inf sentinels (they sort to the end; a
                        // window that touches one has an infinite radius and drops out of a sample's minimum)
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    float cand[K + 1];
    int nc = 0;
#pragma unroll
    for (int a = 0; a + K < N; ++a) {
      if (a <= i && i <= a + K) {
        if (i == a || i == a + K)
          cand[nc] = s[a + K] - s[a];
        else
          cand[nc] = fmaxf(s[i] - s[a], s[a + K] - s[i]);
        ++nc;
      }
    }
    float r = cand[0];
    if (nc == 2) r = fminf(r, cand[1]);
    if (nc >= 3) r = fminf(fminf(r, cand[1]), cand[2]);
    if (nc == 4) r = fminf(r, cand[3]);
    if (nc >= 5) r = fminf(fminf(r, cand[3]), cand[4]);
    if (nc == 6) r = fminf(r, cand[5]);
    const float lg = lg2_pos(fmaxf(r, min_dist));
    acc += (FULL || i < n) ? lg : 0.f;
  }
  return acc;
}

// FULL: n = 32.  Otherwise 17 <= n < 32 samples per item: the TMA box has n rows and the rows n .. 31 of every ring
// slot hold +inf, written once -- a pair with a sentinel gets an infinite (or, sentinel with sentinel, NaN -> ignored)
// distance and sorts behind the real ones, the sentinel keys of a dimension sort to the end (see above).
template <bool FULL>
__global__ void __launch_bounds__(E32_WARPS * 32, 2)
entropy32_kernel(const __grid_constant__ CUtensorMap tmZ, int64_t n_items, int n, int D, float min_dist, double c_term,
                 double *__restrict__ h_z, double *__restrict__ h_mvn) {
  constexpr int N = 32, K = 5;
  const float inv_n = FULL ? 1.f / N : 1.f / (float)n;
  extern __shared__ __align__(128) float smem32[];  // ring | Chebyshev tables | full[RING], empty[RING] mbarriers
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float *ring = smem32;
  float *dm = ring + E32_RING * E32_SLOT_FLOATS;
  const uint32_t ring_u32 = e16_smem_u32(ring);
  const uint32_t full_u32 = e16_smem_u32(dm + E32_NDM * E32_DM);
  const uint32_t empty_u32 = full_u32 + 8u * E32_RING;
  const int spi = (D + E32_COLS - 1) / E32_COLS;  // tiles per item
  const int64_t G = gridDim.x;
  const int64_t n_my = (int64_t)blockIdx.x < n_items ? (n_items - blockIdx.x + G - 1) / G : 0;
  const int64_t n_steps = n_my * spi;
  const bool want_joint = h_mvn != nullptr;

  if (threadIdx.x == 0) {
    for (int s = 0; s < E32_RING; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(full_u32 + 8u * s) : "memory");
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(empty_u32 + 8u * s), "r"(E32_WARPS) : "memory");
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmZ) : "memory");
  }
  for (int e = threadIdx.x; e < E32_NDM * E32_DM; e += E32_WARPS * 32) dm[e] = 0.f;  // the diagonals stay zero
  if (!FULL) {
    for (int s = 0; s < E32_RING; ++s)
      for (int e = n * E32_COLS + threadIdx.x; e < E32_SLOT_FLOATS; e += E32_WARPS * 32) ring[s * E32_SLOT_FLOATS + e] = INFINITY;
  }
  __syncthreads();

  // copy stream: thread 0, three tiles ahead of the arithmetic
  int64_t c_item = blockIdx.x, c_step = 0;
  int c_j = 0;
  auto issue = [&]() {
    if (c_step < n_steps) {
      if (threadIdx.x == 0) {
        const uint32_t slot = (uint32_t)c_step & (E32_RING - 1);
        if (c_step >= E32_RING) e32_mbar_wait(empty_u32 + 8u * slot, (uint32_t)((c_step >> 2) - 1) & 1u);
        const uint32_t bar = full_u32 + 8u * slot;
        const uint32_t dst = ring_u32 + slot * (uint32_t)(E32_SLOT_FLOATS * 4);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar),
                     "r"((FULL ? N : n) * E32_COLS * 4)
                     : "memory");
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
            "l"(&tmZ), "r"(bar), "r"(c_j * E32_COLS), "r"((int)(c_item * (FULL ? N : n)))
            : "memory");
      }
      if (++c_j == spi) {
        c_j = 0;
        c_item += G;
      }
      ++c_step;
    }
  };
  if (warp == 0) {
    issue();
    issue();
    issue();
  }

  int64_t step = 0;
  int it = 0;
  for (int64_t item = blockIdx.x; item < n_items; item += G, ++it) {
    float acc[128];  // slot 16 h + t - 1: max over the lane's dimensions of |x_(8w+h) - x_(8w+h+t)|, t = 1 .. 16
#pragma unroll
    for (int p = 0; p < 128; ++p) acc[p] = 0.f;
#pragma unroll 1
    for (int jstep = 0; jstep < spi; ++jstep, ++step) {
      const uint32_t slot = (uint32_t)step & (E32_RING - 1);
      if (warp == 0) issue();  // tile step + 3 goes to the slot every warp released at the end of step - 1
      e32_mbar_wait(full_u32 + 8u * slot, (uint32_t)(step >> 2) & 1u);
      const float *tile = ring + slot * E32_SLOT_FLOATS;
      if (want_joint) {
#pragma unroll 1
        for (int sub = 0; sub < 2; ++sub) {
          const float2 *t2 = reinterpret_cast<const float2 *>(tile) + sub * 32 + lane;  // row stride: 64 float2
          const int base = warp * 8;
          float2 x[8];
#pragma unroll
          for (int h = 0; h < 8; ++h) x[h] = t2[(base + h) * 64];
#pragma unroll
          for (int r = 1; r <= 23; ++r) {
            const float2 y = t2[((base + r) & 31) * 64];
#pragma unroll
            for (int h = 0; h < 8; ++h) {
              const int t = r - h;
              if (t >= 1 && t <= 16) {
                const float2 d = sub2(x[h], y);
                acc[16 * h + t - 1] = fmaxf(fmaxf(acc[16 * h + t - 1], fabsf(d.x)), fabsf(d.y));
              }
            }
          }
        }
      }
      {
        float v[N];
        const float *col = tile + warp * 32 + lane;
#pragma unroll
        for (int i = 0; i < N; ++i) v[i] = col[i * E32_COLS];
        sort_network<N>(v);
        const float sl = sum_log2_knn5_sorted<N, FULL>(v, min_dist, n);
        const int j = jstep * E32_COLS + warp * 32 + lane;
        if (j < D) h_z[item * (int64_t)D + j] = c_term + (double)(kLn2 * (1.f + sl * inv_n));
      }
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(empty_u32 + 8u * slot) : "memory");
    }
    if (want_joint) {
      float *dmb = dm + (it & (E32_NDM - 1)) * E32_DM;
      float keep[4];
#pragma unroll
      for (int p = 0; p < 128; ++p) {
        const float m = warp_max_f32(acc[p]);
        if ((p & 31) == lane) keep[p >> 5] = m;
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int p = lane + 32 * q;
        const int a = warp * 8 + (p >> 4), b = (a + (p & 15) + 1) & 31;
        dmb[a * 33 + b] = keep[q];
        dmb[b * 33 + a] = keep[q];
      }
      // The tables are finished FOUR items at a time, one per warp: a finishing warp that sorts its table while the
      // other three wait at the next barrier cost a sort per item; four warps sorting four tables side by side cost a
      // quarter of that, and the CTA meets at a barrier once per four items.  Eight tables in two groups: a group is
      // rewritten after the NEXT group's barrier, which every warp passes only after its own sort of this one.
      if ((it & 3) == 3 || item + G >= n_items) {
        __syncthreads();
        const int t = (it & ~3) + warp;  // the item (counted per CTA) this warp finishes
        if (t <= it) {
          const float *tb = dm + (t & (E32_NDM - 1)) * E32_DM;
          float v[N];
#pragma unroll
          for (int b = 0; b < N; ++b) v[b] = tb[lane * 33 + b];
          sort_network<N>(v);  // v[0] = 0 (self); v[K] = k-th neighbour
          float lg = lg2_pos(fmaxf(v[K], min_dist));
          if (!FULL && lane >= n) lg = 0.f;  // sentinel rows
          lg = warp_sum32(lg);
          if (lane == 0)
            h_mvn[(int64_t)blockIdx.x + (int64_t)t * G] = c_term + (double)D * (double)(kLn2 * (1.f + lg * inv_n));
        }
      }
    }
  }
}

// ---------------------------------- generic path ------------------------------------------
constexpr int kEntropyMaxN = 128;  // largest n_mc of the generic kernels (local arrays / the n x n matrix in shared memory)

template <int NMAX>
__global__ void __launch_bounds__(128)
entropy_generic_dim_kernel(const float *__restrict__ z, int64_t n_items, int n, int D, int k, float min_dist,
                           double c_term, double *__restrict__ h_z) {
  const int64_t total = n_items * (int64_t)D;
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const int64_t item = e / D;
  const int j = (int)(e % D);
  const float *zi = z + item * (int64_t)n * D + j;
  float s[NMAX];
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

// n_mc > 32: one block per item, the n x n Chebyshev matrix in dynamic shared memory (row stride n + 1)
__global__ void __launch_bounds__(128)
entropy_generic_joint_big_kernel(const float *__restrict__ z, int64_t n_items, int n, int D, int k, float min_dist,
                                 double c_term, double *__restrict__ h_mvn) {
  extern __shared__ float dmb[];  // [n][n + 1]
  __shared__ float part[4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t item = blockIdx.x;
  const float *zi = z + item * (int64_t)n * D;
  const int ld = n + 1;
  for (int p = warp; p < n * n; p += 4) {
    const int a = p / n, b = p - a * n;
    if (b <= a) continue;
    float m = 0.f;
    for (int j = lane; j < D; j += 32) m = fmaxf(m, fabsf(zi[(int64_t)a * D + j] - zi[(int64_t)b * D + j]));
    m = warp_max32(m);
    if (lane == 0) {
      dmb[a * ld + b] = m;
      dmb[b * ld + a] = m;
    }
  }
  for (int a = threadIdx.x; a < n; a += blockDim.x) dmb[a * ld + a] = 0.f;
  __syncthreads();
  float lg = 0.f;
  for (int a = threadIdx.x; a < n; a += blockDim.x) {
    float r = 0.f;  // element of row a with exactly k predecessors under the total order (value, index)
    for (int b = 0; b < n; ++b) {
      const float vb = dmb[a * ld + b];
      int rank = 0;
      for (int c = 0; c < n; ++c) {
        const float vc = dmb[a * ld + c];
        rank += (vc < vb || (vc == vb && c < b)) ? 1 : 0;
      }
      if (rank == k) r = vb;
    }
    lg += __log2f(fmaxf(r, min_dist));
  }
  lg = warp_sum32(lg);
  if (lane == 0) part[warp] = lg;
  __syncthreads();
  if (threadIdx.x == 0) {
    const float tot = (part[0] + part[1]) + (part[2] + part[3]);
    h_mvn[item] = c_term + (double)D * (double)(kLn2 * (1.f + tot / (float)n));
  }
}

}  // namespace runia

using namespace runia;

extern "C" int runia_mcd_entropy_f32(const float *z, int64_t n_items, int n_mc, int D, int k, double min_dist,
                                     double digamma_term, double *h_z, double *h_mvn, void *stream) {
  RUNIA_NVTX();
  RUNIA_REQUIRE(n_items >= 0 && D > 0, RUNIA_E_BADARG, "mcd_entropy: bad sizes n_items=%lld D=%d",
                (long long)n_items, D);
  RUNIA_REQUIRE(n_mc >= 2 && n_mc <= kEntropyMaxN, RUNIA_E_UNSUPPORTED, "mcd_entropy: n_mc=%d outside [2, %d]", n_mc,
                kEntropyMaxN);
  RUNIA_REQUIRE(k >= 1 && k < n_mc, RUNIA_E_BADARG, "mcd_entropy: k=%d must satisfy 1 <= k < n_mc=%d", k, n_mc);
  if (n_items == 0) return RUNIA_OK;
  RUNIA_REQUIRE(z && h_z, RUNIA_E_BADARG, "mcd_entropy: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if (n_mc == 16 && k == 5 && D % 4 == 0 && (reinterpret_cast<uintptr_t>(z) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(h_z) & 15) == 0 && n_items * 16 < (int64_t)0x7fffffff) {
    static PerDeviceFlag attr16;
    if (!attr16) {
      RUNIA_CUDA(cudaFuncSetAttribute(entropy16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kEntropy16Smem));
      uint8_t pa[120], pb[120];
      int p = 0;
      for (int a = 0; a < 16; ++a)
        for (int b = a + 1; b < 16; ++b) {
          pa[p] = (uint8_t)a;
          pb[p] = (uint8_t)b;
          ++p;
        }
      RUNIA_CUDA(cudaMemcpyToSymbol(kPairA16, pa, sizeof(pa)));
      RUNIA_CUDA(cudaMemcpyToSymbol(kPairB16, pb, sizeof(pb)));
      attr16 = true;
    }
    CUtensorMap tmz;
    int rc = tc::make_plain_map(&tmz, z, n_items * 16, D, 16, 64);
    if (rc) return rc;
    // persistent warps: three CTAs of four warps per SM (register-limited), items round-robin over warps
    int ctas_per_sm = 3;
    if (const char *e = getenv("RUNIA_B200_E16_CTAS")) ctas_per_sm = atoi(e);  // occupancy experiments
    const unsigned grid = (unsigned)std::min<int64_t>(ceil_div(n_items, E16_WARPS), (int64_t)ctas_per_sm * kNumSMs);
    entropy16_kernel<<<grid, E16_WARPS * 32, kEntropy16Smem, st>>>(tmz, n_items, D, (float)min_dist, digamma_term, h_z,
                                                                   h_mvn);
    count_launch();
    return finish_launch("mcd_entropy(16)");
  }
  if (n_mc == 16 && k == 5) {
    constexpr int WARPS = 4;
    constexpr size_t smem = (size_t)WARPS * 120 * 32 * sizeof(float);
    static PerDeviceFlag attr;
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
  if (n_mc > 16 && n_mc <= 32 && k == 5 && D % 4 == 0 && (reinterpret_cast<uintptr_t>(z) & 15) == 0 &&
      n_items * n_mc < (int64_t)0x7fffffff && !getenv("RUNIA_B200_E32_OFF")) {
    static PerDeviceFlag attr32;
    if (!attr32) {
      RUNIA_CUDA(cudaFuncSetAttribute(entropy32_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kEntropy32Smem));
      RUNIA_CUDA(cudaFuncSetAttribute(entropy32_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kEntropy32Smem));
      attr32 = true;
    }
    CUtensorMap tmz;
    int rc = tc::make_plain_map(&tmz, z, n_items * n_mc, D, n_mc, E32_COLS);
    if (rc) return rc;
    // one item per CTA at a time, two CTAs of four warps per SM (register-limited), items round-robin over CTAs
    const unsigned grid = (unsigned)std::min<int64_t>(n_items, (int64_t)2 * kNumSMs);
    if (n_mc == 32)
      entropy32_kernel<true><<<grid, E32_WARPS * 32, kEntropy32Smem, st>>>(tmz, n_items, n_mc, D, (float)min_dist,
                                                                           digamma_term, h_z, h_mvn);
    else
      entropy32_kernel<false><<<grid, E32_WARPS * 32, kEntropy32Smem, st>>>(tmz, n_items, n_mc, D, (float)min_dist,
                                                                            digamma_term, h_z, h_mvn);
    count_launch();
    return finish_launch("mcd_entropy(32)");
  }
  if (k == 5 && n_mc >= 6 && n_mc <= 32) {  // the reference's k for every n_mc > 5 (evaluation/entropy.py:66)
    if (n_mc <= 8)
      launch_entropy_np<8>(z, n_items, n_mc, D, (float)min_dist, digamma_term, h_z, h_mvn, st);
    else if (n_mc <= 16)
      launch_entropy_np<16>(z, n_items, n_mc, D, (float)min_dist, digamma_term, h_z, h_mvn, st);
    else
      launch_entropy_np<32>(z, n_items, n_mc, D, (float)min_dist, digamma_term, h_z, h_mvn, st);
    count_launch();
    return finish_launch("mcd_entropy(np)");
  }
  const int64_t total = n_items * (int64_t)D;
  if (n_mc <= 32)
    entropy_generic_dim_kernel<32><<<(unsigned)ceil_div(total, 128), 128, 0, st>>>(z, n_items, n_mc, D, k, (float)min_dist,
                                                                                  digamma_term, h_z);
  else
    entropy_generic_dim_kernel<kEntropyMaxN><<<(unsigned)ceil_div(total, 128), 128, 0, st>>>(
        z, n_items, n_mc, D, k, (float)min_dist, digamma_term, h_z);
  count_launch();
  if (h_mvn && n_mc <= 32) {
    entropy_generic_joint_kernel<<<(unsigned)ceil_div(n_items, 4), 128, 0, st>>>(z, n_items, n_mc, D, k,
                                                                                (float)min_dist, digamma_term, h_mvn);
    count_launch();
  } else if (h_mvn) {
    const size_t dyn = (size_t)n_mc * (n_mc + 1) * sizeof(float);
    static PerDeviceFlag attrb;
    if (!attrb) {
      RUNIA_CUDA(cudaFuncSetAttribute(entropy_generic_joint_big_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)((size_t)kEntropyMaxN * (kEntropyMaxN + 1) * sizeof(float))));
      attrb = true;
    }
    entropy_generic_joint_big_kernel<<<(unsigned)n_items, 128, dyn, st>>>(z, n_items, n_mc, D, k, (float)min_dist,
                                                                         digamma_term, h_mvn);
    count_launch();
  }
  return finish_launch("mcd_entropy(generic)");
}
