// (f3) MC-DropBlock sampler fused with the spatial reducer -- the producer of the rows get_dl_h_z consumes:
// MCSamplerModule.forward (feature_extraction/abstract_classes.py:81-101, layer_type "Conv"): n_mc DropBlock2D
// layers applied to one activation map, each followed by get_mean_or_fullmean_ls_sample(..., "fullmean")
// (feature_extraction/utils.py:70-92).  DropBlock2D is the third-party `dropblock==0.3.0` (not vendored in the
// reference); its published forward is
//     seed       = rand(B, H, W) < drop_prob / block_size^2
//     block_mask = 1 - max_pool2d(seed, block_size, stride 1, padding block_size / 2)   (even sizes: last row/col cropped)
//     out        = x * block_mask[:, None] * block_mask.numel() / block_mask.sum()
// The reference runs this n_mc times per image (16 full passes over the map plus 16 mean reductions).  Here the
// Bernoulli seeds stay with the caller (drawn with torch's generator exactly like DropBlock2D does, so the RNG
// stream is the reference's), one small kernel dilates them into block masks and counts the kept cells, and one
// pass over the activation map produces all n_mc reduced samples:
//     out[b * n_mc + m, c] = sum_{hw kept by mask m} x[b, c, hw] / kept(m, b)
// (= mean_hw(x * block_mask * HW / kept); normalisation per image, i.e. the reference's per-image call).
// Rows come out item-major ([B * n_mc, C]) -- the layout evaluation/entropy.py:56-63 splits.
#include <algorithm>

#include "common.cuh"

namespace runia {

// Block masks, transposed for the reducer: maskT[b][p][m] (m padded to nmc_pad, pads pre-zeroed) as floats
// (1 keep / 0 drop); kept[m * B + b] = number of kept cells of mask (m, b).
__global__ void dropblock_mask_kernel(const uint8_t *__restrict__ seed, int B, int H, int W, int bs, int nmc_pad,
                                      float *__restrict__ maskT, int *__restrict__ kept) {
  const int img = blockIdx.x;  // (m, b) pair, m major like the seed tensor
  const int m = img / B, b = img - m * B;
  const int HW = H * W, pad = bs / 2;
  const uint8_t *s = seed + (size_t)img * HW;
  int mine = 0;
  for (int p = threadIdx.x; p < HW; p += blockDim.x) {
    const int h = p / W, w = p - h * W;
    bool hit = false;
    for (int dh = 0; dh < bs; ++dh) {
      const int hh = h - pad + dh;
      if (hh < 0 || hh >= H) continue;
      for (int dw = 0; dw < bs; ++dw) {
        const int ww = w - pad + dw;
        if (ww >= 0 && ww < W) hit |= s[hh * W + ww] != 0;
      }
    }
    maskT[((size_t)b * HW + p) * nmc_pad + m] = hit ? 0.f : 1.f;
    mine += hit ? 0 : 1;
  }
  mine = __reduce_add_sync(0xffffffffu, mine);
  if ((threadIdx.x & 31) == 0 && mine) atomicAdd(kept + img, mine);
}

__device__ __forceinline__ void cp_async_4(float *smem_dst, const float *gmem_src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src)
               : "memory");
}
__device__ __forceinline__ void cp_async_16(float *smem_dst, const float *gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src)
               : "memory");
}
__device__ __forceinline__ float2 sp_fma2(float2 a, float2 b, float2 c) {  // FFMA2
  float2 r;
  asm("{.reg .b64 ra, rb, rc, rd; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; mov.b64 rc, {%6, %7}; "
      "fma.rn.f32x2 rd, ra, rb, rc; mov.b64 {%0, %1}, rd;}"
      : "=f"(r.x), "=f"(r.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return r;
}

constexpr int SP_WARPS = 4;
constexpr int SP_CH = 32;    // channels per warp (one per lane)
constexpr int SP_HWC = 128;  // most spatial positions staged per round

// One warp owns 32 consecutive channels of one image: x[b, c0 .. c0+31, :, :] is one contiguous run of 32 * HW
// floats, staged through shared memory with asynchronous copies (16-byte when the run can be copied linearly:
// a single round and an odd H * W, else 4-byte into rows of odd stride), so that lane c walks its own channel
// without bank conflicts while the n_mc mask values of a position are broadcast LDS.128.
// hwc = positions per round, stride = row stride of the tile (odd).
template <int NMC>
__global__ void __launch_bounds__(SP_WARPS * 32) mc_dropblock_mean_kernel(const float *__restrict__ x, const float *__restrict__ maskT,
                                                                         const int *__restrict__ kept, int B, int C, int HW,
                                                                         int n_mc, int hwc, int stride, float *__restrict__ out) {
  extern __shared__ __align__(16) float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int per_warp = SP_CH * stride + hwc * NMC;
  float *tile = smem + warp * per_warp;  // [32][stride]
  float *smask = tile + SP_CH * stride;  // [hwc][NMC]
  const int cblocks = (C + SP_CH - 1) / SP_CH;
  const int64_t units = (int64_t)B * cblocks;
  for (int64_t u = (int64_t)blockIdx.x * SP_WARPS + warp; u < units; u += (int64_t)gridDim.x * SP_WARPS) {
    const int b = (int)(u / cblocks), c0 = (int)(u % cblocks) * SP_CH;
    const int nch = min(SP_CH, C - c0);
    const float *xb = x + ((size_t)b * C + c0) * HW;
    float2 acc[NMC / 2];
#pragma unroll
    for (int m = 0; m < NMC / 2; ++m) acc[m] = make_float2(0.f, 0.f);
    for (int p0 = 0; p0 < HW; p0 += hwc) {
      const int np = min(hwc, HW - p0);
      __syncwarp();
      if (np == HW && stride == HW && (reinterpret_cast<uintptr_t>(xb) & 15) == 0) {
        const int total = nch * HW, nvec = total >> 2;
        for (int i = lane; i < nvec; i += 32) cp_async_16(tile + 4 * i, xb + 4 * i);
        for (int e = 4 * nvec + lane; e < total; e += 32) cp_async_4(tile + e, xb + e);
      } else {
        for (int c = 0; c < nch; ++c)
          for (int p = lane; p < np; p += 32) cp_async_4(tile + c * stride + p, xb + (size_t)c * HW + p0 + p);
      }
      const float *mg = maskT + ((size_t)b * HW + p0) * NMC;  // np * NMC contiguous floats, 16-byte aligned
      for (int i = lane; i < np * (NMC / 4); i += 32) cp_async_16(smask + 4 * i, mg + 4 * i);
      asm volatile("cp.async.wait_all;" ::: "memory");
      __syncwarp();
      if (lane < nch) {
        const float *mine = tile + lane * stride;
#pragma unroll 2
        for (int p = 0; p < np; ++p) {
          const float v = mine[p];
          const float2 v2 = make_float2(v, v);
#pragma unroll
          for (int m4 = 0; m4 < NMC; m4 += 4) {
            const float4 k = *reinterpret_cast<const float4 *>(smask + p * NMC + m4);
            acc[m4 / 2] = sp_fma2(v2, make_float2(k.x, k.y), acc[m4 / 2]);
            acc[m4 / 2 + 1] = sp_fma2(v2, make_float2(k.z, k.w), acc[m4 / 2 + 1]);
          }
        }
      }
    }
    // 1 / kept per sample: lane m holds the reciprocal of mask (m, b); kept == 0 gives inf and 0 * inf = NaN,
    // like x * 0 * numel / 0 in the reference
    const float inv_mine = lane < n_mc ? 1.f / (float)kept[(size_t)lane * B + b] : 0.f;
#pragma unroll
    for (int m = 0; m < NMC; ++m) {
      const float inv = __shfl_sync(0xffffffffu, inv_mine, m);
      const float a = (m & 1) ? acc[m / 2].y : acc[m / 2].x;
      if (m < n_mc && lane < nch) out[((size_t)b * n_mc + m) * C + c0 + lane] = a * inv;
    }
  }
}

template <int NMC>
static int launch_sampler(const float *x, const float *maskT, const int *kept, int B, int C, int HW, int n_mc, float *out,
                          cudaStream_t st) {
  const int hwc = std::min(HW, SP_HWC), stride = hwc | 1;
  const size_t smem = (size_t)SP_WARPS * (SP_CH * stride + hwc * NMC) * sizeof(float);
  static PerDeviceFlag attr;
  if (!attr) {
    RUNIA_CUDA(cudaFuncSetAttribute(mc_dropblock_mean_kernel<NMC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)((size_t)SP_WARPS * (SP_CH * (SP_HWC | 1) + SP_HWC * NMC) * sizeof(float))));
    attr = true;
  }
  int per_sm = 1;
  RUNIA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mc_dropblock_mean_kernel<NMC>, SP_WARPS * 32, smem));
  const int64_t units = (int64_t)B * ((C + SP_CH - 1) / SP_CH);
  const unsigned grid = (unsigned)std::min<int64_t>(ceil_div(units, SP_WARPS), (int64_t)kNumSMs * std::max(per_sm, 1));
  mc_dropblock_mean_kernel<NMC><<<grid, SP_WARPS * 32, smem, st>>>(x, maskT, kept, B, C, HW, n_mc, hwc, stride, out);
  return RUNIA_OK;
}

static int nmc_padded(int n_mc) { return n_mc <= 8 ? 8 : (n_mc <= 16 ? 16 : 32); }

// "FC" / "RPN" layers (feature_extraction/abstract_classes.py:81-101 without the spatial mean): the masked maps
// themselves, out[m, ((b * C + c) * HW + p)] = x[b, c, p] * mask(m, b, p) * (B * HW) / sum_{b, p} mask(m, ., .)
// (DropBlock2D normalises over the whole batch mask it is given).
__global__ void __launch_bounds__(256)
dropblock_apply_kernel(const float *__restrict__ x, const float *__restrict__ maskT, const int *__restrict__ kept, int B, int C,
                       int HW, int n_mc, int nmc_pad, float *__restrict__ out) {
  const int64_t per = (int64_t)B * C * HW, total = per * n_mc;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int m = (int)(e / per);
    const int64_t r = e - (int64_t)m * per;
    const int p = (int)(r % HW), b = (int)(r / ((int64_t)C * HW));
    int tot = 0;
    for (int bb = 0; bb < B; ++bb) tot += kept[(size_t)m * B + bb];
    const float mk = maskT[((size_t)b * HW + p) * nmc_pad + m];
    out[e] = __ldg(x + r) * mk * (float)((int64_t)B * HW) / (float)tot;  // tot == 0: 0 * inf = NaN, like upstream
  }
}

}  // namespace runia

using namespace runia;

extern "C" size_t runia_mc_dropblock_workspace_bytes(int B, int H, int W, int n_mc) {
  if (B < 1 || H < 1 || W < 1 || n_mc < 1 || n_mc > 32) return 0;
  const size_t kept_bytes = (((size_t)n_mc * B * sizeof(int)) + 255) / 256 * 256;
  return kept_bytes + (size_t)B * H * W * nmc_padded(n_mc) * sizeof(float);
}

extern "C" int runia_mc_dropblock_mean_f32(const float *x, const uint8_t *seed, int B, int C, int H, int W, int n_mc,
                                           int block_size, float *out, void *ws, size_t ws_bytes, void *stream) {
  RUNIA_NVTX();
  RUNIA_REQUIRE(B >= 1 && C >= 1 && H >= 1 && W >= 1 && n_mc >= 1 && block_size >= 1, RUNIA_E_BADARG,
                "mc_dropblock: needs B, C, H, W, n_mc, block_size >= 1");
  RUNIA_REQUIRE(n_mc <= 32, RUNIA_E_UNSUPPORTED, "mc_dropblock: n_mc=%d not supported (max 32)", n_mc);
  RUNIA_REQUIRE((int64_t)H * W <= (1 << 24), RUNIA_E_UNSUPPORTED, "mc_dropblock: H * W = %lld not supported (max 2^24)",
                (long long)H * W);
  RUNIA_REQUIRE(x && seed && out && ws, RUNIA_E_BADARG, "mc_dropblock: null pointer");
  RUNIA_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 15) == 0, RUNIA_E_BADARG, "mc_dropblock: workspace must be 16-byte aligned");
  const size_t need = runia_mc_dropblock_workspace_bytes(B, H, W, n_mc);
  RUNIA_REQUIRE(ws_bytes >= need, RUNIA_E_BADARG, "mc_dropblock: workspace of %zu bytes, %zu needed", ws_bytes, need);
  cudaStream_t st = (cudaStream_t)stream;
  const int HW = H * W, pad = nmc_padded(n_mc);
  const size_t kept_bytes = (((size_t)n_mc * B * sizeof(int)) + 255) / 256 * 256;
  int *kept = (int *)ws;
  float *maskT = (float *)((char *)ws + kept_bytes);
  RUNIA_CUDA(cudaMemsetAsync(ws, 0, need, st));  // counters and the padded mask columns
  dropblock_mask_kernel<<<(unsigned)(n_mc * B), 128, 0, st>>>(seed, B, H, W, block_size, pad, maskT, kept);
  int rc;
  if (pad == 8)
    rc = launch_sampler<8>(x, maskT, kept, B, C, HW, n_mc, out, st);
  else if (pad == 16)
    rc = launch_sampler<16>(x, maskT, kept, B, C, HW, n_mc, out, st);
  else
    rc = launch_sampler<32>(x, maskT, kept, B, C, HW, n_mc, out, st);
  if (rc != RUNIA_OK) return rc;
  count_launch(2);
  return finish_launch("mc_dropblock");
}

extern "C" int runia_mc_dropblock_apply_f32(const float *x, const uint8_t *seed, int B, int C, int H, int W, int n_mc,
                                            int block_size, float *out, void *ws, size_t ws_bytes, void *stream) {
  RUNIA_NVTX();
  RUNIA_REQUIRE(B >= 1 && C >= 1 && H >= 1 && W >= 1 && n_mc >= 1 && block_size >= 1, RUNIA_E_BADARG,
                "mc_dropblock_apply: needs B, C, H, W, n_mc, block_size >= 1");
  RUNIA_REQUIRE(n_mc <= 32, RUNIA_E_UNSUPPORTED, "mc_dropblock_apply: n_mc=%d not supported (max 32)", n_mc);
  RUNIA_REQUIRE(x && seed && out && ws, RUNIA_E_BADARG, "mc_dropblock_apply: null pointer");
  const size_t need = runia_mc_dropblock_workspace_bytes(B, H, W, n_mc);
  RUNIA_REQUIRE(ws_bytes >= need, RUNIA_E_BADARG, "mc_dropblock_apply: workspace of %zu bytes, %zu needed", ws_bytes, need);
  cudaStream_t st = (cudaStream_t)stream;
  const int HW = H * W, pad = nmc_padded(n_mc);
  const size_t kept_bytes = (((size_t)n_mc * B * sizeof(int)) + 255) / 256 * 256;
  int *kept = (int *)ws;
  float *maskT = (float *)((char *)ws + kept_bytes);
  RUNIA_CUDA(cudaMemsetAsync(ws, 0, need, st));
  dropblock_mask_kernel<<<(unsigned)(n_mc * B), 128, 0, st>>>(seed, B, H, W, block_size, pad, maskT, kept);
  const int64_t total = (int64_t)n_mc * B * C * HW;
  const unsigned grid = (unsigned)std::min<int64_t>(ceil_div(total, 256), (int64_t)kNumSMs * 16);
  dropblock_apply_kernel<<<grid, 256, 0, st>>>(x, maskT, kept, B, C, HW, n_mc, pad, out);
  count_launch(2);
  return finish_launch("mc_dropblock_apply");
}
