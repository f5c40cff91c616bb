// tcgen05 (3xTF32) versions of the contraction-shaped scorers: LaREM / ViM row norm (a3, a7),
// PCA projection (a2), kNN candidate filter (a5) and KDE log-sum-exp (a4).  See tc_gemm.cuh for the
// kernel anatomy.  The SIMT kernels in score_gemm.cu / distance.cu stay as the FP32 reference
// path (K % 4 != 0, unaligned pointers) and as the accuracy yardstick (DESIGN.md "3xTF32 vs FP32").
#include <cuda.h>

#include <algorithm>
#include <mutex>

#include "tc_gemm.cuh"

namespace runia {
namespace tc {

// ---------------------------------------------------------------------------------------------
// operand split: hi = tf32(x), lo = tf32(x - hi)          (weights / banks, once at setup)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) split_tf32_kernel(const float *__restrict__ x, int64_t total,
                                                         float *__restrict__ hi, float *__restrict__ lo) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    const float v = x[e];
    const float h = to_tf32(v);
    hi[e] = h;
    lo[e] = to_tf32(v - h);
  }
}

// ---------------------------------------------------------------------------------------------
// epilogues
// ---------------------------------------------------------------------------------------------
struct RowNormEpi {  // MD: -sum sign*v^2 ; ViM: -alpha*sqrt(sum v^2) + lse(logits)
  const float *sign;
  int r, mode, C;
  const float *logits;
  float alpha;
  double *out64;
  float *out32;
  int64_t M;
  int64_t row;
  float acc;
  __device__ void set_stage(uint32_t) {}
  template <class P> __device__ void bind(const P *) {}
  __device__ void begin(int, int64_t row_) {
    row = row_;
    acc = 0.f;
  }
  __device__ void consume(int64_t col0, const float (&v)[32], int, int) {
    if (sign) {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float s = (col0 + j < r) ? __ldg(sign + col0 + j) : 0.f;
        acc = fmaf(s * v[j], v[j], acc);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) acc = fmaf(v[j], v[j], acc);  // columns >= r are exact zeros (TMA zero fill)
    }
  }
  __device__ void panel_done(int) {}
  __device__ void finish() {
    if (row >= M) return;
    float sc;
    if (mode == RUNIA_ROWNORM_MD) {
      sc = -acc;
    } else {
      const float *l = logits + row * (int64_t)C;
      float m = -INFINITY;
      for (int c = 0; c < C; ++c) m = fmaxf(m, l[c]);
      float s = 0.f;
      for (int c = 0; c < C; ++c) s += expf(l[c] - m);
      sc = -alpha * sqrtf(fmaxf(acc, 0.f)) + (m + logf(s));
    }
    if (out64) out64[row] = (double)sc;
    if (out32) out32[row] = sc;
  }
};

struct PcaEpi {  // Z[row, col] = v * inv_scale[col], written with TMA stores (full 128-byte lines, clipped at the edges)
  alignas(64) CUtensorMap tmZ;  // [N, d] fp32, box 32 x 32, 128B swizzle
  const float *inv_scale;
  int d;
  int64_t M;
  const CUtensorMap *mapZ;
  uint32_t stage;
  int64_t row;
  bool pending;
  __device__ void set_stage(uint32_t smem) {
    stage = smem;
    pending = false;
  }
  template <class P>
  __device__ void bind(const P *param) { mapZ = &param->tmZ; }
  __device__ void begin(int, int64_t row_) { row = row_; }
  __device__ void consume(int64_t col0, const float (&v)[32], int, int lane) {
    const int64_t row0 = row - lane;  // warp-uniform
    if (row0 >= M || col0 >= d) return;
    if (pending) {  // the previous box must have been read out of the staging buffer
      if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      __syncwarp();
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      float4 o;
      const int j = 4 * c;
      o.x = v[j + 0] * ((inv_scale && col0 + j + 0 < d) ? __ldg(inv_scale + col0 + j + 0) : 1.f);
      o.y = v[j + 1] * ((inv_scale && col0 + j + 1 < d) ? __ldg(inv_scale + col0 + j + 1) : 1.f);
      o.z = v[j + 2] * ((inv_scale && col0 + j + 2 < d) ? __ldg(inv_scale + col0 + j + 2) : 1.f);
      o.w = v[j + 3] * ((inv_scale && col0 + j + 3 < d) ? __ldg(inv_scale + col0 + j + 3) : 1.f);
      const uint32_t addr = stage + (uint32_t)lane * 128u + (uint32_t)((c ^ (lane & 7)) << 4);
      asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(o.x), "f"(o.y), "f"(o.z), "f"(o.w) : "memory");
    }
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
      asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(mapZ),
                   "r"((int)col0), "r"((int)row0), "r"(stage)
                   : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    pending = true;
  }
  __device__ void panel_done(int) {}
  __device__ void finish() {  // the staging buffer must outlive the last store's read (tile end / kernel exit)
    if (pending && (threadIdx.x & 31) == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    __syncwarp();
    pending = false;
  }
};

// (a6) class-conditional Mahalanobis: out = max_c -sum_j sign_j (y_j - m_cj)^2 over the classes that
// had training samples; one accumulator per class in registers (C <= kClassMax)
constexpr int kClassMax = 16;
// packed FP32 pairs (FADD2 / FFMA2 on sm_100a): one issue slot for two lanes of the per-class distance
__device__ __forceinline__ float2 sub_f32x2(float2 a, float2 b) {
  float2 r;
  asm("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; sub.f32x2 rc, ra, rb; mov.b64 {%0, %1}, rc;}"
      : "=f"(r.x), "=f"(r.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return r;
}
__device__ __forceinline__ float2 fma_f32x2(float2 a, float2 b, float2 c) {
  float2 r;
  asm("{.reg .b64 ra, rb, rc, rd; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; mov.b64 rc, {%6, %7}; "
      "fma.rn.f32x2 rd, ra, rb, rc; mov.b64 {%0, %1}, rd;}"
      : "=f"(r.x), "=f"(r.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return r;
}

struct ClassCondEpi {
  const float *sign, *Mc;  // [r] or nullptr; [C, r]
  const int32_t *valid;    // [C]
  int r, C;
  double *out64;
  float *out32;
  int64_t M;
  int64_t row;
  float2 cls[kClassMax];  // two partial sums per class (even / odd columns)
  __device__ void set_stage(uint32_t) {}
  template <class P> __device__ void bind(const P *) {}
  __device__ void begin(int, int64_t row_) {
    row = row_;
#pragma unroll
    for (int c = 0; c < kClassMax; ++c) cls[c] = make_float2(0.f, 0.f);
  }
  __device__ void consume(int64_t col0, const float (&v)[32], int, int) {
    if (col0 + 31 < r && !sign) {  // positive-definite precision (the usual case): no sign plane
      // unrolled over the classes (register accumulators): 8 LDG.128 + 16 FADD2 + 16 FFMA2 per class
#pragma unroll
      for (int c = 0; c < kClassMax; ++c) {
        if (c < C) {
          const float *mc = Mc + (size_t)c * r + col0;
          float2 a0 = cls[c], a1 = make_float2(0.f, 0.f);  // two independent FFMA2 chains
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            const float4 m4 = __ldg(reinterpret_cast<const float4 *>(mc + j));  // same address in every lane
            const float4 n4 = __ldg(reinterpret_cast<const float4 *>(mc + j + 4));
            const float2 d0 = sub_f32x2(make_float2(v[j], v[j + 1]), make_float2(m4.x, m4.y));
            const float2 d1 = sub_f32x2(make_float2(v[j + 2], v[j + 3]), make_float2(m4.z, m4.w));
            const float2 d2 = sub_f32x2(make_float2(v[j + 4], v[j + 5]), make_float2(n4.x, n4.y));
            const float2 d3 = sub_f32x2(make_float2(v[j + 6], v[j + 7]), make_float2(n4.z, n4.w));
            a0 = fma_f32x2(d0, d0, a0);
            a1 = fma_f32x2(d1, d1, a1);
            a0 = fma_f32x2(d2, d2, a0);
            a1 = fma_f32x2(d3, d3, a1);
          }
          cls[c] = make_float2(a0.x + a1.x, a0.y + a1.y);
        }
      }
      return;
    }
    // indefinite precision (sign plane) or the ragged last panel: one rolled copy of the code for all
    // classes -- the unrolled form of these paths evicted the mainloop from the instruction cache
#pragma unroll 1
    for (int c = 0; c < C; ++c) {
      const float *mc = Mc + (size_t)c * r + col0;
      float part = 0.f;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        if (col0 + j < r) {
          const float dlt = v[j] - __ldg(mc + j);
          part = fmaf((sign ? __ldg(sign + col0 + j) : 1.f) * dlt, dlt, part);
        }
      }
#pragma unroll
      for (int k = 0; k < kClassMax; ++k)
        if (k == c) cls[k].x += part;
    }
  }
  __device__ void panel_done(int) {}
  __device__ void finish() {
    if (row >= M) return;
    float best = -INFINITY;
#pragma unroll
    for (int c = 0; c < kClassMax; ++c)
      if (c < C && valid[c]) {
        const float sc = -(cls[c].x + cls[c].y);
        if (sc > best) best = sc;  // NaN never wins, like np.max after NaN -> -inf
      }
    if (out64) out64[row] = (double)best;
    if (out32) out32[row] = best;
  }
};

// (a10) ReAct / DICE / DICE+ReAct head: the 32-column panel holds the C <= 32 class logits of the row
// (clip applied by the converters); out = logsumexp_c (v_c + b_c)      postprocessors.py:1464-1472, 1340-1352
constexpr int kNarrowN = 32;      // widest narrow panel (C <= 32); C <= 16 uses a 16-column panel
constexpr int kNarrowStages = 5;  // 36 KB stages: 80 KB of rows in flight per CTA (HBM-bound stream)
// NW = panel width (16 or 32).  Accumulator columns (tc_gemm.cuh, NCAT), h = NW / 2:
//   [0, h) A_hi x B_hi for classes 0..h-1 (CTA 0's rows), [h, 2h) A_hi x B_lo for the same classes,
//   [2h, 3h) / [3h, 4h) the same for classes h..NW-1 (CTA 1's rows), [2 NW, 3 NW) A_lo x B_hi in class order.
template <int NW>
struct LinearLseEpi {
  const float *bias;  // [C]
  int C;
  float *out;
  int64_t M;
  int64_t row;
  __device__ void set_stage(uint32_t) {}
  template <class P> __device__ void bind(const P *) {}
  __device__ void begin(int, int64_t row_) { row = row_; }
  float l[NW];
  __device__ void finalize() {
    if (row >= M) return;
    constexpr float kLog2e = 1.4426950408889634f, kLn2 = 0.6931471805599453f;
    float m = -INFINITY;
#pragma unroll
    for (int c = 0; c < NW; ++c) {
      l[c] = c < C ? l[c] + __ldg(bias + c) : -INFINITY;
      m = fmaxf(m, l[c]);
    }
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < NW; ++c) s += exp2f((l[c] - m) * kLog2e);  // exp2(-inf) = 0 for the padding
    out[row] = fmaf(__log2f(s), kLn2, m);
  }
  __device__ void consume(int64_t col0, const float (&v)[32], int, int) {
    constexpr int h = NW / 2;
    if (NW == 32) {  // three 32-column groups
      if (col0 == 0) {
#pragma unroll
        for (int c = 0; c < h; ++c) l[c] = v[c] + v[h + c];
      } else if (col0 == 32) {
#pragma unroll
        for (int c = 0; c < h; ++c) l[h + c] = v[c] + v[h + c];
      } else {
#pragma unroll
        for (int c = 0; c < NW; ++c) l[c] += v[c];
        finalize();
      }
    } else {  // NW == 16: columns 0..31 hold both hi products, 32..47 the A_lo product
      if (col0 == 0) {
#pragma unroll
        for (int c = 0; c < h; ++c) {
          l[c] = v[c] + v[h + c];
          l[h + c] = v[2 * h + c] + v[3 * h + c];
        }
      } else {
#pragma unroll
        for (int c = 0; c < NW; ++c) l[c] += v[c];
        finalize();
      }
    }
  }
  __device__ void panel_done(int) {}
  __device__ void finish() {}
};

// (a10) the same head for ANY class count (ImageNet: C = 1000, COCO: C = 80): the classes are the columns of the
// 256-wide panels, the log-sum-exp runs online across panels in log2 units (like KdeEpi); columns >= C are the TMA's
// zero fill and are masked out.      postprocessors.py:1444-1474, 1325-1354, 1591-1621
struct WideLinearLseEpi {
  const float *bias;  // [C], 16-byte aligned
  int C;
  float *out;
  int64_t M;
  int64_t row;
  float m, s;
  __device__ void set_stage(uint32_t) {}
  template <class P> __device__ void bind(const P *) {}
  __device__ void begin(int, int64_t row_) {
    row = row_;
    m = -INFINITY;
    s = 0.f;
  }
  __device__ void consume(int64_t col0, const float (&v)[32], int, int) {
    constexpr float kLog2e = 1.4426950408889634f;
    if (col0 >= C) return;
    float t[32];
    float tmax = -INFINITY;
    if (col0 + 32 <= C) {
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        const float4 b4 = __ldg(reinterpret_cast<const float4 *>(bias + col0) + g);  // same address in every lane
        t[4 * g + 0] = (v[4 * g + 0] + b4.x) * kLog2e;
        t[4 * g + 1] = (v[4 * g + 1] + b4.y) * kLog2e;
        t[4 * g + 2] = (v[4 * g + 2] + b4.z) * kLog2e;
        t[4 * g + 3] = (v[4 * g + 3] + b4.w) * kLog2e;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) t[j] = col0 + j < C ? (v[j] + __ldg(bias + col0 + j)) * kLog2e : -INFINITY;
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) tmax = fmaxf(tmax, t[j]);
    if (tmax == -INFINITY) return;
    const float m_new = fmaxf(m, tmax);
    float acc = s * exp2f(m - m_new);
#pragma unroll
    for (int j = 0; j < 32; ++j) acc += exp2f(t[j] - m_new);
    s = acc;
    m = m_new;
  }
  __device__ void panel_done(int) {}
  __device__ void finish() {
    constexpr float kLn2 = 0.6931471805599453f;
    if (row < M) out[row] = (m + __log2f(s)) * kLn2;
  }
};

// (a9) DDU / GMM: columns are C blocks of dpad whitened coordinates; per class
// lp_c = -0.5 sum_j (v_j - off_j)^2 + logconst_c, out = logsumexp_c lp_c (online, per thread)
struct GmmEpi {
  const float *off, *logconst;
  int dpad, C;
  float *out;
  int64_t M;
  int64_t row;
  float acc, run_m, run_s;
  __device__ void set_stage(uint32_t) {}
  template <class P> __device__ void bind(const P *) {}
  __device__ void begin(int, int64_t row_) {
    row = row_;
    acc = 0.f;
    run_m = -INFINITY;
    run_s = 0.f;
  }
  __device__ void consume(int64_t col0, const float (&v)[32], int, int) {
    if (col0 >= (int64_t)C * dpad) return;  // zero-filled tail of the last panel
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float4 o4 = __ldg(reinterpret_cast<const float4 *>(off + col0 + j));
      const float d0 = v[j] - o4.x, d1 = v[j + 1] - o4.y, d2 = v[j + 2] - o4.z, d3 = v[j + 3] - o4.w;
      acc = fmaf(d0, d0, acc);
      acc = fmaf(d1, d1, acc);
      acc = fmaf(d2, d2, acc);
      acc = fmaf(d3, d3, acc);
    }
    if ((col0 + 32) % dpad == 0) {  // dpad is a multiple of 128: a class always ends on a 32-column chunk
      const int c = (int)(col0 / dpad);
      const float lp = fmaf(-0.5f, acc, __ldg(logconst + c));
      const float m_new = fmaxf(run_m, lp);
      if (m_new != -INFINITY) {
        run_s = run_s * expf(run_m - m_new) + expf(lp - m_new);
        run_m = m_new;
      }
      acc = 0.f;
    }
  }
  __device__ void panel_done(int) {}
  __device__ void finish() {
    if (row < M) out[row] = run_m + logf(run_s);
  }
};

struct KdeEpi {  // online log-sum-exp of -|q-b|^2/(2h^2) in log2 units
  const float *qn, *bn;
  int64_t Nq, b_hi;
  float scale;  // -0.5/h^2 * log2(e)
  float *part_m, *part_s;
  int splits, split;
  int64_t row;
  float q2, m, s;
  __device__ void set_stage(uint32_t) {}
  template <class P> __device__ void bind(const P *) {}
  __device__ void begin(int, int64_t row_) {
    row = row_;
    q2 = row < Nq ? __ldg(qn + row) : 0.f;
    m = -INFINITY;
    s = 0.f;
  }
  __device__ void consume(int64_t col0, const float (&v)[32], int, int) {
    float t[32];
    float tmax = -INFINITY;
    if (col0 + 32 <= b_hi) {
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        const float4 b4 = __ldg(reinterpret_cast<const float4 *>(bn + col0) + g);  // same address in every lane
        const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int j = 4 * g + u;
          t[j] = fmaxf(fmaf(-2.f, v[j], q2 + bb[u]), 0.f) * scale;
          tmax = fmaxf(tmax, t[j]);
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int64_t col = col0 + j;
        const float b2 = col < b_hi ? __ldg(bn + col) : 0.f;
        const float dist = fmaxf(fmaf(-2.f, v[j], q2 + b2), 0.f);
        t[j] = col < b_hi ? dist * scale : -INFINITY;
        tmax = fmaxf(tmax, t[j]);
      }
    }
    if (tmax == -INFINITY) return;
    const float m_new = fmaxf(m, tmax);
    float acc = s * exp2f(m - m_new);
#pragma unroll
    for (int j = 0; j < 32; ++j) acc += exp2f(t[j] - m_new);
    s = acc;
    m = m_new;
  }
  __device__ void panel_done(int) {}
  __device__ void finish() {
    if (row < Nq) {
      part_m[(size_t)row * splits + split] = m;
      part_s[(size_t)row * splits + split] = s;
    }
  }
};

// kNN candidate filter.  Thread t owns query row t of the tile: threshold and count live in
// registers, candidates (approximate distance, bank index) are appended to the row's buffer in
// global memory, row-major [row][split][entry] so that the re-rank kernel gathers a row's lists with
// coalesced loads.  With the seed threshold (below) a split appends a few hundred entries per row
// and never shrinks; the shrink is the overflow guard for long splits (large banks): the owning
// thread tightens its threshold by bisection on the order-preserving integer key of the distance
// until between kcap and `target` entries lie below it, and compacts its buffer in place.  Entries
// dropped at any time have distance >= the row's final threshold, which is what the certification
// in the re-rank kernel relies on.
__device__ __forceinline__ uint32_t ord_key(float v) {
  const uint32_t u = __float_as_uint(v);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord_val(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

struct KnnEpi {
  const float *qn, *bn;
  int64_t Nq, b_hi;
  float2 *buf;              // [Nq, splits, capp] entries (approximate distance, bank index as bits)
  int32_t *counts;          // [Nq, splits] entries left in each list
  float *thr_fin;           // [Nq, splits] final threshold of each list: every bank row of the split that is not listed
                            // has an approximate distance >= it (+inf: the whole range is listed)
  int kcap, fin_max, capp, splits, split;
  const uint32_t *thr_key;  // [Nq] seed: order-preserving key of an upper bound on the kcap-th distance (0 = none)
  int64_t row;
  float2 *mine;             // this row's list
  float q2, thr;
  int cnt;
  bool live;
  uint32_t stage;           // this warp's 4 KB staging buffer (shared-memory address)
  __device__ void set_stage(uint32_t smem) { stage = smem; }
  template <class P> __device__ void bind(const P *) {}
  __device__ void begin(int, int64_t row_) {
    row = row_;
    live = row < Nq;
    q2 = live ? __ldg(qn + row) : 0.f;
    thr = INFINITY;
    if (live && thr_key) {
      const uint32_t key = __ldg(thr_key + row);
      if (key != 0u && key < ord_key(INFINITY)) thr = ord_val(key);
    }
    cnt = 0;
    mine = buf + ((size_t)(live ? row : 0) * splits + split) * (size_t)capp;
  }
  // The epilogue warp runs alone on its scheduler: every instruction of this loop is on the critical path
  // of the TMEM hand-back, so a column costs one FFMA, one compare and (predicated) one address, one
  // 8-byte store and one increment; the |b|^2 terms come as eight broadcast LDG.128.
  // The epilogue warp runs alone on its scheduler and a panel's 256 columns must be through before the mainloop has
  // produced the next-but-one panel (8 k-blocks of 4 MMAs at K = 256: ~5,000 cycles), so the common case -- no
  // column of the chunk passes the filter -- is straight-line code: eight broadcast LDG.128 for the |b|^2 terms issued
  // together, 32 FFMA, a 32-bit hit mask.  Only a thread with hits parks its 32 distances in the warp's staging
  // buffer ([column][lane]: conflict-free) and walks the set bits (registers cannot be indexed by a bit position).
  __device__ void consume(int64_t col0, const float (&v)[32], int, int lane) {
    if (!live) return;
    if (col0 + 32 <= b_hi) {
      float4 b4[8];
#pragma unroll
      for (int g = 0; g < 8; ++g) b4[g] = __ldg(reinterpret_cast<const float4 *>(bn + col0) + g);  // same address in every lane
      float dist[32];
      uint32_t mask = 0u;
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        dist[4 * g + 0] = fmaf(-2.f, v[4 * g + 0], q2 + b4[g].x);
        dist[4 * g + 1] = fmaf(-2.f, v[4 * g + 1], q2 + b4[g].y);
        dist[4 * g + 2] = fmaf(-2.f, v[4 * g + 2], q2 + b4[g].z);
        dist[4 * g + 3] = fmaf(-2.f, v[4 * g + 3], q2 + b4[g].w);
      }
#pragma unroll
      for (int j = 0; j < 32; ++j) mask |= dist[j] < thr ? (1u << j) : 0u;
      if (mask) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          asm volatile("st.shared.f32 [%0], %1;" ::"r"(stage + (uint32_t)(j * 32 + lane) * 4u), "f"(dist[j]) : "memory");
        while (mask) {
          const int j = __ffs((int)mask) - 1;
          mask &= mask - 1u;
          float dj;
          asm volatile("ld.shared.f32 %0, [%1];" : "=f"(dj) : "r"(stage + (uint32_t)(j * 32 + lane) * 4u) : "memory");
          mine[cnt] = make_float2(dj, __int_as_float((int)col0 + j));
          ++cnt;
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int64_t col = col0 + j;
        if (col < b_hi) {
          const float dist = fmaf(-2.f, v[j], q2 + __ldg(bn + col));
          if (dist < thr) {
            mine[cnt] = make_float2(dist, __int_as_float((int)col));
            ++cnt;
          }
        }
      }
    }
  }
  // the entries were written by this thread: plain loads (its own stores are visible to it)
  __device__ __forceinline__ float ld_d(int e) const { return mine[e].x; }
  static constexpr int SCAN = 32;  // independent loads in flight per scan step
  __device__ __forceinline__ int count_below(int n, uint32_t bound) const {
    int c = 0;
    for (int e0 = 0; e0 < n; e0 += SCAN) {
      float v[SCAN];
#pragma unroll
      for (int j = 0; j < SCAN; ++j) v[j] = (e0 + j < n) ? ld_d(e0 + j) : INFINITY;
#pragma unroll
      for (int j = 0; j < SCAN; ++j) c += (e0 + j < n && ord_key(v[j]) < bound) ? 1 : 0;
    }
    return c;
  }
  // shrink this thread's buffer to between kcap and `target` entries
  __device__ void shrink(int target) {
    const int n = cnt;
    if (n <= target) return;
    // bisection on integer keys: invariant count(key < hi) >= kcap > count(key < lo)
    uint32_t lo, hi;
    int c_hi = n;
    {
      uint32_t kmin = 0xffffffffu, kmax = 0u;
      for (int e0 = 0; e0 < n; e0 += SCAN) {
        float v[SCAN];
#pragma unroll
        for (int j = 0; j < SCAN; ++j) v[j] = (e0 + j < n) ? ld_d(e0 + j) : 0.f;
#pragma unroll
        for (int j = 0; j < SCAN; ++j)
          if (e0 + j < n) {
            const uint32_t k = ord_key(v[j]);
            kmin = k < kmin ? k : kmin;
            kmax = k > kmax ? k : kmax;
          }
      }
      lo = kmin;  // count(key < kmin) = 0 < kcap
      hi = kmax == 0xffffffffu ? kmax : kmax + 1u;
    }
    while (c_hi > target && hi - lo > 1u) {
      const uint32_t mid = lo + ((hi - lo) >> 1);
      const int c = count_below(n, mid);
      if (c >= kcap) {
        hi = mid;
        c_hi = c;
      } else {
        lo = mid;
      }
    }
    // in-place compaction (write index never passes the read index; reads are batched first)
    const bool ties = c_hi > target;  // more than `target` entries share the key `lo` around rank kcap:
    int ties_left = 0;                // keep everything below it and just enough of the ties
    if (ties) ties_left = kcap - count_below(n, lo);
    const uint32_t bound = ties ? lo : hi;
    int w = 0;
    constexpr int CB = 16;  // compaction batch
    for (int e0 = 0; e0 < n; e0 += CB) {
      float2 v[CB];
#pragma unroll
      for (int j = 0; j < CB; ++j) v[j] = (e0 + j < n) ? mine[e0 + j] : make_float2(INFINITY, 0.f);
#pragma unroll
      for (int j = 0; j < CB; ++j) {
        if (e0 + j >= n) continue;
        const uint32_t k = ord_key(v[j].x);
        bool keep = k < bound;
        if (ties && k == lo && ties_left > 0) {
          keep = true;
          --ties_left;
        }
        if (keep) {
          mine[w] = v[j];
          ++w;
        }
      }
    }
    // later ties are rejected by the strict `< thr` filter
    thr = fminf(thr, ord_val(bound));
    cnt = w;
  }
  __device__ void panel_done(int) {
    // overflow guard: the next panel can append at most TN entries
    if (live && cnt > capp - TN) shrink((kcap + capp - TN) / 2);  // to between the floor and the trigger
  }
  __device__ void finish() {
    if (!live) return;
    shrink(fin_max);  // no-op unless the list is longer than the buffer guard allows
    counts[(size_t)row * splits + split] = cnt;
    thr_fin[(size_t)row * splits + split] = thr;
  }
};

// Seed: an upper bound on every query's kcap-th smallest distance, with no buffers at all.  The seed
// columns are cut into kcap / 4 groups of `ppg` panels (256 bank rows each); per group the thread keeps the 4
// smallest distances it has seen in registers.  B = max over groups of the 4th smallest has at least
// kcap bank rows at distance <= B, so the main pass starts with the threshold nextafter(B) and only
// ever buffers rows that can still matter.  Groups are independent: they are split over several CTA
// pairs and combined with an atomic max on the order-preserving key.
// Seed and candidate pass are the SAME single TF32 product (A_hi x B_hi, same instruction order: the same pair gets
// the same approximate distance in both), so B bounds the candidate pass's own numbers.  What the single product costs
// is distance from the EXACT value: |approx - exact| <= s = eps1 * (|q|^2 + max|b|^2) / 2 (A_hi is a 19-bit
// truncation: 2^-10 relative, B_hi rounds to nearest: 2^-11; |q.b| <= (|q|^2 + |b|^2) / 2; doubled for the -2 q.b
// term; plus FP32 accumulation).  The re-rank evaluates exactly every candidate within 2 s of the k-th smallest
// approximate distance a_k, so the lists must hold every bank row below a_k + 2 s: the bound is widened by
// `slack` = 2 eps1 per unit of (|q|^2 + max|b|^2) / 2 (B >= a_k).

struct KnnSeedEpi {
  const float *qn, *bn;
  int64_t Nq, b_hi;
  uint32_t *thr_key;
  const uint32_t *bn_max;  // [1] bit pattern of max_b |b|^2
  int ppg;  // panels per group
  float slack;  // widening of the bound per unit of (|q|^2 + max|b|^2) / 2
  int64_t row;
  float q2, m0, m1, m2, m3, bound;
  int in_group;
  bool live;
  __device__ void set_stage(uint32_t) {}
  template <class P> __device__ void bind(const P *) {}
  __device__ void begin(int, int64_t row_) {
    row = row_;
    live = row < Nq;
    q2 = live ? __ldg(qn + row) : 0.f;
    m0 = m1 = m2 = m3 = INFINITY;
    bound = -INFINITY;
    in_group = 0;
  }
  __device__ __forceinline__ void insert(float t) {
    if (t < m3) {  // rare after the first few columns of a group
      float a = fminf(m0, t);
      t = fmaxf(m0, t);
      m0 = a;
      a = fminf(m1, t);
      t = fmaxf(m1, t);
      m1 = a;
      a = fminf(m2, t);
      t = fmaxf(m2, t);
      m2 = a;
      m3 = fminf(m3, t);
    }
  }
  __device__ void consume(int64_t col0, const float (&v)[32], int, int) {
    if (col0 + 32 <= b_hi) {
      float4 b4[8];
#pragma unroll
      for (int g = 0; g < 8; ++g) b4[g] = __ldg(reinterpret_cast<const float4 *>(bn + col0) + g);  // same address in every lane
      float dist[32];
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        dist[4 * g + 0] = fmaf(-2.f, v[4 * g + 0], q2 + b4[g].x);
        dist[4 * g + 1] = fmaf(-2.f, v[4 * g + 1], q2 + b4[g].y);
        dist[4 * g + 2] = fmaf(-2.f, v[4 * g + 2], q2 + b4[g].z);
        dist[4 * g + 3] = fmaf(-2.f, v[4 * g + 3], q2 + b4[g].w);
      }
      // after the first columns of a group almost no chunk holds a distance below the 4th smallest so far: one
      // 3-input-min tree decides, the insert chains run only for the groups of four that matter
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        const float m = fminf(fminf(dist[4 * g], dist[4 * g + 1]), fminf(dist[4 * g + 2], dist[4 * g + 3]));
        if (m < m3) {
          insert(dist[4 * g + 0]);
          insert(dist[4 * g + 1]);
          insert(dist[4 * g + 2]);
          insert(dist[4 * g + 3]);
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int64_t col = col0 + j;
        insert(fmaf(-2.f, v[j], q2 + (col < b_hi ? __ldg(bn + col) : INFINITY)));
      }
    }
  }
  __device__ void panel_done(int) {
    if (++in_group == ppg) {
      bound = fmaxf(bound, m3);
      m0 = m1 = m2 = m3 = INFINITY;
      in_group = 0;
    }
  }
  __device__ void finish() {
    if (!live) return;
    if (in_group != 0) bound = INFINITY;  // a partial group proves nothing (the launcher never produces one)
    bound += slack * 0.5f * (q2 + __uint_as_float(__ldg(bn_max)));
    const uint32_t key = ord_key(bound);
    atomicMax(thr_key + row, key < 0xfffffffeu ? key + 1u : key);  // exclusive bound: the filter is strict
  }
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
tc_knn_seed_kernel(const __grid_constant__ CUtensorMap tmA, int K, const __grid_constant__ CUtensorMap tmB_hi,
                   const __grid_constant__ CUtensorMap tmB_lo, int seed_panels, int panels_per_split, KnnSeedEpi epi) {
  extern __shared__ unsigned char smem_raw[];
  Work w;
  w.tile_first = blockIdx.x >> 1;
  w.tile_end = (int64_t)(blockIdx.x >> 1) + 1;
  w.tile_step = 1;
  w.panel_lo = (int)blockIdx.y * panels_per_split;
  w.panel_hi = min(seed_panels, w.panel_lo + panels_per_split);
  const Prologue pro{nullptr, INFINITY};
  run_tiles<KnnSeedEpi, TN, 2 * STAGES, false, 1, true>(&tmA, K, pro, &tmB_hi, &tmB_lo, w, epi, smem_raw);
}

template <class E>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
tc_kernel(const __grid_constant__ CUtensorMap tmA, int64_t M, int K, Prologue pro, const __grid_constant__ CUtensorMap tmB_hi,
          const __grid_constant__ CUtensorMap tmB_lo, int panels_total, const __grid_constant__ E epi_param) {
  extern __shared__ unsigned char smem_raw[];
  E epi = epi_param;
  epi.bind(&epi_param);  // policies that own a tensor map need its address in parameter space
  Work w;  // persistent: CTA pair p takes the 256-row tiles p, p + #pairs, ...
  w.tile_first = blockIdx.x >> 1;
  w.tile_end = (M + TM2 - 1) / TM2;
  w.tile_step = gridDim.x >> 1;
  w.panel_lo = 0;
  w.panel_hi = panels_total;
  run_tiles(&tmA, K, pro, &tmB_hi, &tmB_lo, w, epi, smem_raw);
}

// Narrow-panel variant (UMMA N = 32): streaming heads whose output is a handful of columns per row.
template <class E, int NW>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
tc_narrow_kernel(const __grid_constant__ CUtensorMap tmA, int64_t M, int K, Prologue pro,
                 const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
                 const __grid_constant__ E epi_param) {
  extern __shared__ unsigned char smem_raw[];
  E epi = epi_param;
  epi.bind(&epi_param);
  Work w;
  w.tile_first = blockIdx.x >> 1;
  w.tile_end = (M + TM2 - 1) / TM2;
  w.tile_step = gridDim.x >> 1;
  w.panel_lo = 0;
  w.panel_hi = 1;
  run_tiles<E, NW, kNarrowStages, true>(&tmA, K, pro, &tmB_hi, &tmB_lo, w, epi, smem_raw);
}

// TF32 tensor peak probe: every CTA pair issues `iters` x 4 back-to-back UMMAs (256 x 256 x 8, kind::tf32, the
// instruction of the scorers) on one resident pair of shared-memory tiles filled with pseudo-random values; no
// loads, no epilogue.  MEASURED_PEAKS.json has no TF32 entry; this is the denominator the tensor-bound rooflines
// use (bench.py), measured with CUDA events on the same box in the same run.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) tc_peak_kernel(int iters, uint32_t seed) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t rank = cluster_ctarank();
  const uint32_t base = (smem_u32(smem_raw) + SMEM_ALIGN - 1) & ~(uint32_t)(SMEM_ALIGN - 1);
  unsigned char *gbase = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sA = base, sB = base + A_PLANE_BYTES, bar = base + A_PLANE_BYTES + B_PLANE_BYTES, slot = bar + 16;
  float *tiles = reinterpret_cast<float *>(gbase);
  uint32_t x = seed * 2654435761u + (blockIdx.x * 128u + threadIdx.x) * 40503u + 1u;
  for (int e = threadIdx.x; e < (A_PLANE_BYTES + B_PLANE_BYTES) / 4; e += 128) {
    x = x * 1664525u + 1013904223u;
    tiles[e] = __uint_as_float((x >> 9) | 0x3f800000u) - 1.5f;  // uniform in [-0.5, 0.5)
  }
  fence_proxy_async();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(slot, 512);
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t *>(gbase + A_PLANE_BYTES + B_PLANE_BYTES + 16);
  if (warp == 0) {
    if (rank == 0 && lane == 0) {
      const uint64_t dA = make_smem_desc(sA), dB = make_smem_desc(sB);
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k4 = 0; k4 < TK / 8; ++k4) {
          const uint64_t adv = (uint64_t)((k4 * 8 * 4) >> 4);
          umma_tf32(tmem_base + (uint32_t)((it & 1) * TN), dA + adv, dB + adv, InstrDesc<TN>::value, (it >= 2 || k4 > 0) ? 1u : 0u);
        }
      }
      umma_commit(bar);
    }
    __syncwarp();
    mbar_wait(bar, 0);  // both CTAs: the multicast commit arrives on each one's barrier
    tc_fence_after();
  }
  __syncthreads();
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

__device__ __forceinline__ Work split_work(int panels_total, int panels_per_split) {
  Work w;  // one 256-row tile x one bank split per CTA pair
  w.tile_first = blockIdx.x >> 1;
  w.tile_end = (int64_t)(blockIdx.x >> 1) + 1;
  w.tile_step = 1;
  w.panel_lo = blockIdx.y * panels_per_split;
  w.panel_hi = min(panels_total, w.panel_lo + panels_per_split);
  return w;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
tc_kde_kernel(const __grid_constant__ CUtensorMap tmA, int64_t M, int K, const __grid_constant__ CUtensorMap tmB_hi,
              const __grid_constant__ CUtensorMap tmB_lo, int panels_total, int panels_per_split, int64_t NB,
              KdeEpi epi) {
  extern __shared__ unsigned char smem_raw[];
  const Work w = split_work(panels_total, panels_per_split);
  epi.split = blockIdx.y;
  const int64_t hi = (int64_t)w.panel_hi * TN;
  epi.b_hi = hi < NB ? hi : NB;
  const Prologue pro{nullptr, INFINITY};
  run_tiles(&tmA, K, pro, &tmB_hi, &tmB_lo, w, epi, smem_raw);
}

template <int PRODUCTS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
tc_knn_kernel(const __grid_constant__ CUtensorMap tmA, int64_t M, int K, const __grid_constant__ CUtensorMap tmB_hi,
              const __grid_constant__ CUtensorMap tmB_lo, int panels_total, int panels_per_split, int64_t NB,
              int split_offset, KnnEpi epi) {
  extern __shared__ unsigned char smem_raw[];
  Work w;
  w.tile_first = blockIdx.x >> 1;
  w.tile_end = (int64_t)(blockIdx.x >> 1) + 1;
  w.tile_step = 1;
  const int split = (int)blockIdx.y + split_offset;
  w.panel_lo = split * panels_per_split;
  w.panel_hi = min(panels_total, w.panel_lo + panels_per_split);
  epi.split = split;
  const int64_t hi = (int64_t)w.panel_hi * TN;
  epi.b_hi = hi < NB ? hi : NB;
  const Prologue pro{nullptr, INFINITY};
  // PRODUCTS = 1: the pass is a FILTER (the re-rank recomputes every survivor exactly in float64 and certifies the
  // result against the rounding bound of this product), at a third of the tensor work of the 3xTF32 contraction.
  // PRODUCTS = 3: the FP32-faithful contraction, for banks whose neighbourhoods are denser than the single-product
  // bound can separate (chosen per bank by the caller: `runia_knn_search_ex_f32`).
  if constexpr (PRODUCTS == 1)
    run_tiles<KnnEpi, TN, 2 * STAGES, false, 1, true>(&tmA, K, pro, &tmB_hi, &tmB_lo, w, epi, smem_raw);
  else
    run_tiles(&tmA, K, pro, &tmB_hi, &tmB_lo, w, epi, smem_raw);
}

// ---------------------------------------------------------------------------------------------
// host side: tensor maps through the driver entry point (no link-time dependency on libcuda)
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  });
  return fn;
}

// [rows, K] fp32 row-major -> box of TNH rows x 32 floats (one CTA's half of a panel), 128B swizzle,
// zero fill out of bounds
static int make_map(CUtensorMap *map, const float *ptr, int64_t rows, int K, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return (int)cudaErrorNotSupported;
  }
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)K * 4};
  cuuint32_t box[2] = {(cuuint32_t)TK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)ptr, dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%lld K=%d)", (int)r, (long long)rows, K);
    return (int)cudaErrorInvalidValue;
  }
  return RUNIA_OK;
}

// plain (unswizzled) 2-D box over a row-major fp32 matrix: the streaming kernels outside this file (entropy.cu)
int make_plain_map(CUtensorMap *map, const float *ptr, int64_t rows, int K, int box_rows, int box_cols) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return (int)cudaErrorNotSupported;
  }
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)K * 4};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)ptr, dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (plain) failed with CUresult %d (rows=%lld K=%d)", (int)r, (long long)rows, K);
    return (int)cudaErrorInvalidValue;
  }
  return RUNIA_OK;
}

static int make_b_map(CUtensorMap *map, const float *ptr, int64_t rows, int K) { return make_map(map, ptr, rows, K, TNH); }
// streamed operand: box of TM rows x 32 floats, lands in the A_hi plane of a stage as raw fp32
static int make_a_map(CUtensorMap *map, const float *ptr, int64_t rows, int K) { return make_map(map, ptr, rows, K, TM); }

constexpr size_t kSmemMax = (size_t)STAGES * STAGE_BYTES + EPI_STAGE_BYTES + SMEM_BAR_BYTES + (size_t)kMaxK * 4 + SMEM_ALIGN;

template <class Kern>
static int set_smem(Kern kern, size_t bytes) {
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) {
    set_error("cudaFuncSetAttribute(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
    return (int)e;
  }
  return RUNIA_OK;
}

bool usable(const void *A, int K, const void *B_hi, const void *B_lo) {
  return (K % 4 == 0) && K <= kMaxK && ((reinterpret_cast<uintptr_t>(A) & 15) == 0) &&
         ((reinterpret_cast<uintptr_t>(B_hi) & 15) == 0) && ((reinterpret_cast<uintptr_t>(B_lo) & 15) == 0);
}

}  // namespace tc
}  // namespace runia

using namespace runia;
using namespace runia::tc;

extern "C" int runia_split_tf32(const float *x, int64_t total, float *hi, float *lo, void *stream) {
  RUNIA_NVTX();
  RUNIA_REQUIRE(total >= 0, RUNIA_E_BADARG, "split_tf32: bad size");
  if (total == 0) return RUNIA_OK;
  RUNIA_REQUIRE(x && hi && lo, RUNIA_E_BADARG, "split_tf32: null pointer");
  const unsigned grid = (unsigned)std::min<int64_t>(ceil_div(total, 256), (int64_t)kNumSMs * 16);
  split_tf32_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, total, hi, lo);
  count_launch();
  return finish_launch("split_tf32");
}

extern "C" int runia_rownorm_score_tc(const float *X, int64_t N, int d, const float *mu, const float *Wt_hi,
                                      const float *Wt_lo, int r, const float *sign, int mode, const float *logits,
                                      int C, float alpha, double *out_f64, float *out_f32, void *stream) {
  RUNIA_NVTX();
  RUNIA_REQUIRE(N >= 0 && d > 0 && r > 0, RUNIA_E_BADARG, "rownorm_score_tc: bad sizes");
  RUNIA_REQUIRE(mode == RUNIA_ROWNORM_MD || mode == RUNIA_ROWNORM_VIM, RUNIA_E_BADARG, "rownorm_score_tc: bad mode");
  if (N == 0) return RUNIA_OK;
  RUNIA_REQUIRE(X && Wt_hi && Wt_lo && (out_f64 || out_f32), RUNIA_E_BADARG, "rownorm_score_tc: null pointer");
  RUNIA_REQUIRE(mode != RUNIA_ROWNORM_VIM || (logits && C > 0), RUNIA_E_BADARG, "rownorm_score_tc: ViM needs logits");
  RUNIA_REQUIRE(usable(X, d, Wt_hi, Wt_lo) && (!mu || (reinterpret_cast<uintptr_t>(mu) & 15) == 0), RUNIA_E_UNSUPPORTED,
                "rownorm_score_tc: needs d %% 4 == 0 and 16-byte aligned pointers");
  CUtensorMap ma, mh, ml;
  int rc = make_a_map(&ma, X, N, d);
  if (rc) return rc;
  rc = make_b_map(&mh, Wt_hi, r, d);
  if (rc) return rc;
  rc = make_b_map(&ml, Wt_lo, r, d);
  if (rc) return rc;
  static PerDeviceFlag attr;
  if (!attr) {
    rc = set_smem(tc_kernel<RowNormEpi>, kSmemMax);
    if (rc) return rc;
    attr = true;
  }
  RowNormEpi epi{sign, r, mode, C, logits, alpha, out_f64, out_f32, N, 0, 0.f};
  const int panels = (int)ceil_div(r, TN);
  dim3 grid(2 * (unsigned)std::min<int64_t>(ceil_div(N, TM2), kNumSMs / 2), 1);  // CTA pairs, one per TPC
  tc_kernel<RowNormEpi><<<grid, THREADS, smem_bytes(d), (cudaStream_t)stream>>>(ma, N, d, Prologue{mu, INFINITY}, mh, ml,
                                                                          panels, epi);
  count_launch();
  return finish_launch("rownorm_score_tc");
}

extern "C" int runia_clip_linear_lse_tc(const float *X, int64_t N, int d, const float *W_hi, const float *W_lo,
                                        const float *b, int C, float clip, float *out, void *stream) {
  RUNIA_NVTX();
  RUNIA_REQUIRE(N >= 0 && d > 0 && C > 0, RUNIA_E_BADARG, "clip_linear_lse_tc: bad sizes");
  if (N == 0) return RUNIA_OK;
  RUNIA_REQUIRE(X && W_hi && W_lo && b && out, RUNIA_E_BADARG, "clip_linear_lse_tc: null pointer");
  RUNIA_REQUIRE(usable(X, d, W_hi, W_lo), RUNIA_E_UNSUPPORTED,
                "clip_linear_lse_tc: needs d %% 4 == 0, d <= %d and 16-byte aligned pointers", kMaxK);
  CUtensorMap ma, mh, ml;
  int rc = make_a_map(&ma, X, N, d);
  if (rc) return rc;
  dim3 grid(2 * (unsigned)std::min<int64_t>(ceil_div(N, TM2), kNumSMs / 2), 1);
  const Prologue pro{nullptr, clip};
  if (C > kNarrowN) {
    // any class count: planes are [C, d]; 256-column panels, online log-sum-exp across them
    RUNIA_REQUIRE((reinterpret_cast<uintptr_t>(b) & 15) == 0, RUNIA_E_UNSUPPORTED,
                  "clip_linear_lse_tc: the bias must be 16-byte aligned");
    rc = make_b_map(&mh, W_hi, C, d);
    if (rc) return rc;
    rc = make_b_map(&ml, W_lo, C, d);
    if (rc) return rc;
    static PerDeviceFlag attrw;
    if (!attrw) {
      rc = set_smem(tc_kernel<WideLinearLseEpi>, kSmemMax);
      if (rc) return rc;
      attrw = true;
    }
    WideLinearLseEpi epi{b, C, out, N, 0, 0.f, 0.f};
    tc_kernel<WideLinearLseEpi><<<grid, THREADS, smem_bytes(d), (cudaStream_t)stream>>>(ma, N, d, pro, mh, ml,
                                                                                  (int)ceil_div(C, TN), epi);
    count_launch();
    return finish_launch("clip_linear_lse_tc(wide)");
  }
  const int nw = C <= 16 ? 16 : 32;                 // planes are [32, d] with rows >= C zero; the first nw are read
  rc = make_map(&mh, W_hi, nw, d, nw / 2);
  if (rc) return rc;
  rc = make_map(&ml, W_lo, nw, d, nw / 2);
  if (rc) return rc;
  static PerDeviceFlag attr;
  if (!attr) {
    rc = set_smem(tc_narrow_kernel<LinearLseEpi<16>, 16>, smem_bytes_variant(kMaxK, 16, kNarrowStages));
    if (rc) return rc;
    rc = set_smem(tc_narrow_kernel<LinearLseEpi<32>, 32>, smem_bytes_variant(kMaxK, 32, kNarrowStages));
    if (rc) return rc;
    attr = true;
  }
  if (nw == 16) {
    LinearLseEpi<16> epi{b, C, out, N, 0};
    tc_narrow_kernel<LinearLseEpi<16>, 16><<<grid, THREADS, smem_bytes_variant(d, 16, kNarrowStages), (cudaStream_t)stream>>>(
        ma, N, d, pro, mh, ml, epi);
  } else {
    LinearLseEpi<32> epi{b, C, out, N, 0};
    tc_narrow_kernel<LinearLseEpi<32>, 32><<<grid, THREADS, smem_bytes_variant(d, 32, kNarrowStages), (cudaStream_t)stream>>>(
        ma, N, d, pro, mh, ml, epi);
  }
  count_launch();
  return finish_launch("clip_linear_lse_tc");
}

extern "C" int runia_tf32_peak_probe(int iters, double *flop_out, void *stream) {
  RUNIA_NVTX();
  RUNIA_REQUIRE(iters >= 1 && iters <= (1 << 24), RUNIA_E_BADARG, "tf32_peak_probe: iters out of range");
  const size_t smem = (size_t)A_PLANE_BYTES + B_PLANE_BYTES + 64 + SMEM_ALIGN;
  static PerDeviceFlag attr;
  if (!attr) {
    int rc = set_smem(tc_peak_kernel, smem);
    if (rc) return rc;
    attr = true;
  }
  const unsigned pairs = kNumSMs / 2;
  tc_peak_kernel<<<2 * pairs, 128, smem, (cudaStream_t)stream>>>(iters, 12345u);
  if (flop_out) *flop_out = (double)pairs * iters * (TK / 8) * 2.0 * TM2 * TN * 8;  // TF32 FLOP issued by the launch
  count_launch();
  return finish_launch("tf32_peak_probe");
}

extern "C" int runia_pca_transform_tc(const float *X, int64_t N, int D0, const float *mean, const float *C_hi,
                                      const float *C_lo, int d, const float *inv_scale, float *Z, void *stream) {
  RUNIA_NVTX();
  RUNIA_REQUIRE(N >= 0 && D0 > 0 && d > 0, RUNIA_E_BADARG, "pca_transform_tc: bad sizes");
  if (N == 0) return RUNIA_OK;
  RUNIA_REQUIRE(X && C_hi && C_lo && Z, RUNIA_E_BADARG, "pca_transform_tc: null pointer");
  RUNIA_REQUIRE(usable(X, D0, C_hi, C_lo) && (!mean || (reinterpret_cast<uintptr_t>(mean) & 15) == 0) &&
                    (reinterpret_cast<uintptr_t>(Z) & 15) == 0 && d % 4 == 0,
                RUNIA_E_UNSUPPORTED, "pca_transform_tc: needs D0 %% 4 == 0, d %% 4 == 0 and 16-byte aligned pointers");
  CUtensorMap ma, mh, ml;
  int rc = make_a_map(&ma, X, N, D0);
  if (rc) return rc;
  rc = make_b_map(&mh, C_hi, d, D0);
  if (rc) return rc;
  rc = make_b_map(&ml, C_lo, d, D0);
  if (rc) return rc;
  static PerDeviceFlag attr;
  if (!attr) {
    rc = set_smem(tc_kernel<PcaEpi>, kSmemMax);
    if (rc) return rc;
    attr = true;
  }
  PcaEpi epi{};
  rc = make_map(&epi.tmZ, Z, N, d, 32);
  if (rc) return rc;
  epi.inv_scale = inv_scale; epi.d = d; epi.M = N;
  const int panels = (int)ceil_div(d, TN);
  dim3 grid(2 * (unsigned)std::min<int64_t>(ceil_div(N, TM2), kNumSMs / 2), 1);
  tc_kernel<PcaEpi><<<grid, THREADS, smem_bytes(D0), (cudaStream_t)stream>>>(ma, N, D0, Prologue{mean, INFINITY}, mh, ml,
                                                                      panels, epi);
  count_launch();
  return finish_launch("pca_transform_tc");
}

extern "C" int runia_classcond_mahalanobis_tc(const float *X, int64_t N, int d, const float *g, const float *Wt_hi,
                                              const float *Wt_lo, int r, const float *sign, const float *Mc,
                                              const int32_t *class_valid, int C, double *out_f64, float *out_f32,
                                              void *stream) {
  RUNIA_NVTX();
  RUNIA_REQUIRE(N >= 0 && d > 0 && r > 0 && C > 0, RUNIA_E_BADARG, "classcond_mahalanobis_tc: bad sizes");
  if (N == 0) return RUNIA_OK;
  RUNIA_REQUIRE(X && Wt_hi && Wt_lo && Mc && class_valid && (out_f64 || out_f32), RUNIA_E_BADARG,
                "classcond_mahalanobis_tc: null pointer");
  RUNIA_REQUIRE(C <= kClassMax, RUNIA_E_UNSUPPORTED, "classcond_mahalanobis_tc: C=%d > %d classes", C, kClassMax);
  RUNIA_REQUIRE(usable(X, d, Wt_hi, Wt_lo) && (!g || (reinterpret_cast<uintptr_t>(g) & 15) == 0) && r % 4 == 0 &&
                    (reinterpret_cast<uintptr_t>(Mc) & 15) == 0 && (!sign || (reinterpret_cast<uintptr_t>(sign) & 15) == 0),
                RUNIA_E_UNSUPPORTED, "classcond_mahalanobis_tc: needs d %% 4 == 0, r %% 4 == 0 and 16-byte aligned pointers");
  CUtensorMap ma, mh, ml;
  int rc = make_a_map(&ma, X, N, d);
  if (rc) return rc;
  rc = make_b_map(&mh, Wt_hi, r, d);
  if (rc) return rc;
  rc = make_b_map(&ml, Wt_lo, r, d);
  if (rc) return rc;
  static PerDeviceFlag attr;
  if (!attr) {
    rc = set_smem(tc_kernel<ClassCondEpi>, kSmemMax);
    if (rc) return rc;
    attr = true;
  }
  ClassCondEpi epi{};
  epi.sign = sign; epi.Mc = Mc; epi.valid = class_valid; epi.r = r; epi.C = C;
  epi.out64 = out_f64; epi.out32 = out_f32; epi.M = N;
  dim3 grid(2 * (unsigned)std::min<int64_t>(ceil_div(N, TM2), kNumSMs / 2), 1);
  tc_kernel<ClassCondEpi><<<grid, THREADS, smem_bytes(d), (cudaStream_t)stream>>>(ma, N, d, Prologue{g, INFINITY}, mh,
                                                                                ml, (int)ceil_div(r, TN), epi);
  count_launch();
  return finish_launch("classcond_mahalanobis_tc");
}

extern "C" int runia_gmm_lse_tc(const float *X, int64_t N, int d, const float *At_hi, const float *At_lo,
                                const float *off, int dpad, const float *logconst, int C, float *out, void *stream) {
  RUNIA_NVTX();
  RUNIA_REQUIRE(N >= 0 && d > 0 && C > 0 && dpad >= d && dpad % 128 == 0, RUNIA_E_BADARG, "gmm_lse_tc: bad sizes");
  if (N == 0) return RUNIA_OK;
  RUNIA_REQUIRE(X && At_hi && At_lo && off && logconst && out, RUNIA_E_BADARG, "gmm_lse_tc: null pointer");
  RUNIA_REQUIRE(usable(X, d, At_hi, At_lo) && (reinterpret_cast<uintptr_t>(off) & 15) == 0, RUNIA_E_UNSUPPORTED,
                "gmm_lse_tc: needs d %% 4 == 0 and 16-byte aligned pointers");
  const int64_t cols = (int64_t)C * dpad;
  CUtensorMap ma, mh, ml;
  int rc = make_a_map(&ma, X, N, d);
  if (rc) return rc;
  rc = make_b_map(&mh, At_hi, cols, d);
  if (rc) return rc;
  rc = make_b_map(&ml, At_lo, cols, d);
  if (rc) return rc;
  static PerDeviceFlag attr;
  if (!attr) {
    rc = set_smem(tc_kernel<GmmEpi>, kSmemMax);
    if (rc) return rc;
    attr = true;
  }
  GmmEpi epi{};
  epi.off = off; epi.logconst = logconst; epi.dpad = dpad; epi.C = C; epi.out = out; epi.M = N;
  dim3 grid(2 * (unsigned)std::min<int64_t>(ceil_div(N, TM2), kNumSMs / 2), 1);
  tc_kernel<GmmEpi><<<grid, THREADS, smem_bytes(d), (cudaStream_t)stream>>>(ma, N, d, Prologue{nullptr, INFINITY}, mh, ml,
                                                                          (int)ceil_div(cols, TN), epi);
  count_launch();
  return finish_launch("gmm_lse_tc");
}

namespace runia {
namespace tc {
// shared with distance.cu
int launch_knn_candidates_tc(const float *Q, const float *qn, int64_t Nq, const float *B_hi, const float *B_lo,
                             const float *bn, int64_t Nb, int d, int kseed, float seed_slack, int kcap, int fin_max, int capp,
                             int splits, int64_t panels_per_split, float *buf_d, int32_t *buf_i, int32_t *counts,
                             uint32_t *thr_key, float *thr_fin, const uint32_t *bn_max, int phase, int products,
                             cudaStream_t st) {
  // phase 1: the seed thresholds of these rows (thr_key; 0 where the bank is too small to seed); phase 2: the
  // candidate pass against thresholds written before; 3: both
  CUtensorMap ma, mh, ml;
  int rc = make_a_map(&ma, Q, Nq, d);
  if (rc) return rc;
  rc = make_b_map(&mh, B_hi, Nb, d);
  if (rc) return rc;
  rc = make_b_map(&ml, B_lo, Nb, d);
  if (rc) return rc;
  const size_t smem = smem_bytes(d);
  static PerDeviceFlag attr;
  if (!attr) {
    rc = set_smem(tc_knn_kernel<1>, kSmemMax);
    if (rc) return rc;
    rc = set_smem(tc_knn_kernel<3>, kSmemMax);
    if (rc) return rc;
    rc = set_smem(tc_knn_seed_kernel, kSmemMax);
    if (rc) return rc;
    attr = true;
  }
  const int64_t tiles = ceil_div(Nq, TM2);
  const int panels = (int)ceil_div(Nb, TN);
  // Seed: kseed / 4 groups of `ppg` full panels each (kseed >= k + 8: the bound has at least kseed bank rows below
  // it).  The bound admits a fraction p ~ 10 / (256 ppg) of the bank, i.e. ~ p Nb / splits appends per (row, split);
  // ppg is chosen to keep that near 200 (no shrink in the main pass) while the seed stays under a quarter of the bank.
  const int groups = kseed / 4;
  int ppg = (int)std::max<int64_t>(1, ceil_div(Nb, (int64_t)5120 * splits));
  ppg = (int)std::min<int64_t>(ppg, std::max<int64_t>(1, (Nb / TN) / (4 * (int64_t)groups)));
  const int seed_panels = groups * ppg;
  const bool seeded = (Nb / TN) >= 2 * (int64_t)seed_panels;
  if (phase & 1) RUNIA_CUDA(cudaMemsetAsync(thr_key, 0, (size_t)Nq * sizeof(uint32_t), st));  // 0 = no seed
  if (seeded && (phase & 1)) {
    int ss = (int)std::min<int64_t>(groups, std::max<int64_t>(1, ceil_div(kNumSMs, tiles)));
    const int gps = (int)ceil_div(groups, ss);  // whole groups per seed split
    ss = (int)ceil_div(groups, gps);
    const int pps = gps * ppg;
    KnnSeedEpi seed{};
    seed.qn = qn; seed.bn = bn; seed.Nq = Nq; seed.b_hi = Nb; seed.thr_key = thr_key; seed.bn_max = bn_max; seed.ppg = ppg; seed.slack = seed_slack;
    dim3 grid0(2 * (unsigned)tiles, (unsigned)ss);
    tc_knn_seed_kernel<<<grid0, THREADS, smem, st>>>(ma, d, mh, ml, seed_panels, pps, seed);
    count_launch();
  }
  if (!(phase & 2)) return finish_launch("knn_seed_tc");
  KnnEpi epi{};
  epi.qn = qn; epi.bn = bn; epi.Nq = Nq; epi.b_hi = Nb;
  epi.buf = reinterpret_cast<float2 *>(buf_d);  // buf_d and buf_i are contiguous: one (distance, index) pair per entry
  (void)buf_i;
  epi.counts = counts;
  epi.thr_fin = thr_fin;
  epi.kcap = kcap; epi.fin_max = fin_max; epi.capp = capp; epi.splits = splits; epi.split = 0;
  epi.thr_key = seeded ? thr_key : nullptr;
  dim3 grid(2 * (unsigned)tiles, (unsigned)splits);
  if (products == 1)
    tc_knn_kernel<1><<<grid, THREADS, smem, st>>>(ma, Nq, d, mh, ml, panels, (int)panels_per_split, Nb, 0, epi);
  else
    tc_knn_kernel<3><<<grid, THREADS, smem, st>>>(ma, Nq, d, mh, ml, panels, (int)panels_per_split, Nb, 0, epi);
  count_launch();
  return finish_launch("knn_candidates_tc");
}

int launch_kde_partial_tc(const float *Q, const float *qn, int64_t Nq, const float *B_hi, const float *B_lo,
                          const float *bn, int64_t Nb, int d, float scale, int splits, int64_t panels_per_split,
                          float *part_m, float *part_s, cudaStream_t st) {
  CUtensorMap ma, mh, ml;
  int rc = make_a_map(&ma, Q, Nq, d);
  if (rc) return rc;
  rc = make_b_map(&mh, B_hi, Nb, d);
  if (rc) return rc;
  rc = make_b_map(&ml, B_lo, Nb, d);
  if (rc) return rc;
  static PerDeviceFlag attr;
  if (!attr) {
    rc = set_smem(tc_kde_kernel, kSmemMax);
    if (rc) return rc;
    attr = true;
  }
  KdeEpi epi{};
  epi.qn = qn; epi.bn = bn; epi.Nq = Nq; epi.b_hi = Nb; epi.scale = scale;
  epi.part_m = part_m; epi.part_s = part_s; epi.splits = splits; epi.split = 0;
  dim3 grid(2 * (unsigned)ceil_div(Nq, TM2), (unsigned)splits);
  tc_kde_kernel<<<grid, THREADS, smem_bytes(d), st>>>(ma, Nq, d, mh, ml, (int)ceil_div(Nb, TN), (int)panels_per_split, Nb,
                                                  epi);
  count_launch();
  return finish_launch("kde_partial_tc");
}
}  // namespace tc
}  // namespace runia
