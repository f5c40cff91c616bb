// FP32 SIMT tile contraction shared by every "stream rows against a small matrix" scorer.
//
//   acc[m, n] = sum_k f(A[m0+m, k]) * B[n0+n, k]          (both operands K-contiguous)
//
// 128 x 128 output tile per CTA, BK = 16, 256 threads, 8 x 8 register tile per thread laid out as
// two 4-wide strips in each direction (rows ty*4.., 64+ty*4..; cols tx*4.., 64+tx*4..) so that
// shared-memory reads are LDS.128 without bank conflicts (A: broadcast, B: 256 contiguous bytes
// per half-warp).  Global -> register -> shared double buffering, one __syncthreads per k-step.
// f() is the fused prologue: subtract a per-column centre and/or clip from above (ReAct).
//
// This is the FP32-faithful path (error ~ sqrt(K) * 2^-24 relative to sum |terms|); the tcgen05
// 3xTF32 variant lives in tc_gemm.cuh and is selected by measured error.
#pragma once
#include "common.cuh"

namespace runia {

constexpr int BM = 128, BN = 128, BK = 16, GEMM_THREADS = 256;
constexpr int LDS_PAD = 4;

struct __align__(16) GemmSmem {
  float As[2][BK][BM + LDS_PAD];
  float Bs[2][BK][BN + LDS_PAD];
};

struct Prologue {
  const float *sub;  // [K] or nullptr
  float clip;        // +inf = none
};

__device__ __forceinline__ int tile_row(int ty, int i) { return (i < 4) ? ty * 4 + i : 64 + ty * 4 + (i - 4); }
__device__ __forceinline__ int tile_col(int tx, int j) { return (j < 4) ? tx * 4 + j : 64 + tx * 4 + (j - 4); }

// Loads 4 consecutive k of one row (zero beyond [rows, K]); applies the prologue when APPLY.
template <bool APPLY>
__device__ __forceinline__ float4 load_k4(const float *__restrict__ P, int64_t rows, int64_t row, int K, int k,
                                          bool vec_ok, const Prologue &pro) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (row < rows) {
    const float *p = P + row * (int64_t)K + k;
    if (vec_ok && k + 3 < K) {
      v = __ldg(reinterpret_cast<const float4 *>(p));
      if (APPLY) {
        if (pro.sub) {
          const float4 s = __ldg(reinterpret_cast<const float4 *>(pro.sub + k));
          v.x -= s.x; v.y -= s.y; v.z -= s.z; v.w -= s.w;
        }
        if (pro.clip < INFINITY) {  // fminf would turn NaN into clip; keep NaN like numpy.clip
          v.x = v.x > pro.clip ? pro.clip : v.x; v.y = v.y > pro.clip ? pro.clip : v.y;
          v.z = v.z > pro.clip ? pro.clip : v.z; v.w = v.w > pro.clip ? pro.clip : v.w;
        }
      }
    } else {
      float t[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (k + q < K) {
          float x = __ldg(p + q);
          if (APPLY) {
            if (pro.sub) x -= __ldg(pro.sub + k + q);
            x = x > pro.clip ? pro.clip : x;
          }
          t[q] = x;
        }
      }
      v = make_float4(t[0], t[1], t[2], t[3]);
    }
  }
  return v;
}

// acc += tile product.  A: [M, K] rows m0.., B: [NB, K] rows n0...
__device__ __forceinline__ void gemm_mainloop(const float *__restrict__ A, int64_t M, int64_t m0,
                                              const float *__restrict__ B, int64_t NB, int64_t n0, int K,
                                              const Prologue &pro, GemmSmem &s, float (&acc)[8][8]) {
  const int t = threadIdx.x;
  const int tx = t & 15, ty = t >> 4;
  const int lr = t >> 2, kq = (t & 3) * 4;
  const bool vec_ok = (K % 4 == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0) &&
                      ((reinterpret_cast<uintptr_t>(B) & 15) == 0) &&
                      (pro.sub == nullptr || (reinterpret_cast<uintptr_t>(pro.sub) & 15) == 0);
  const int nk = (K + BK - 1) / BK;

  float4 ra0, ra1, rb0, rb1;
  ra0 = load_k4<true>(A, M, m0 + lr, K, kq, vec_ok, pro);
  ra1 = load_k4<true>(A, M, m0 + lr + 64, K, kq, vec_ok, pro);
  rb0 = load_k4<false>(B, NB, n0 + lr, K, kq, vec_ok, pro);
  rb1 = load_k4<false>(B, NB, n0 + lr + 64, K, kq, vec_ok, pro);
  int buf = 0;
  auto stage = [&](int b) {
    s.As[b][kq + 0][lr] = ra0.x; s.As[b][kq + 1][lr] = ra0.y; s.As[b][kq + 2][lr] = ra0.z; s.As[b][kq + 3][lr] = ra0.w;
    s.As[b][kq + 0][lr + 64] = ra1.x; s.As[b][kq + 1][lr + 64] = ra1.y; s.As[b][kq + 2][lr + 64] = ra1.z; s.As[b][kq + 3][lr + 64] = ra1.w;
    s.Bs[b][kq + 0][lr] = rb0.x; s.Bs[b][kq + 1][lr] = rb0.y; s.Bs[b][kq + 2][lr] = rb0.z; s.Bs[b][kq + 3][lr] = rb0.w;
    s.Bs[b][kq + 0][lr + 64] = rb1.x; s.Bs[b][kq + 1][lr + 64] = rb1.y; s.Bs[b][kq + 2][lr + 64] = rb1.z; s.Bs[b][kq + 3][lr + 64] = rb1.w;
  };
  __syncthreads();  // previous users of the shared tiles are done
  stage(0);
  __syncthreads();

  for (int kt = 0; kt < nk; ++kt) {
    if (kt + 1 < nk) {
      const int k = (kt + 1) * BK + kq;
      ra0 = load_k4<true>(A, M, m0 + lr, K, k, vec_ok, pro);
      ra1 = load_k4<true>(A, M, m0 + lr + 64, K, k, vec_ok, pro);
      rb0 = load_k4<false>(B, NB, n0 + lr, K, k, vec_ok, pro);
      rb1 = load_k4<false>(B, NB, n0 + lr + 64, K, k, vec_ok, pro);
    }
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4 *>(&s.As[buf][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4 *>(&s.As[buf][k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4 *>(&s.Bs[buf][k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4 *>(&s.Bs[buf][k][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nk) stage(buf ^ 1);
    __syncthreads();
    buf ^= 1;
  }
}

__device__ __forceinline__ void zero_acc(float (&acc)[8][8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
}

}  // namespace runia
