// (f3) Object-level reducers -- feature_extraction/object_level.py:254-309 (`_reduce_features_to_rois`) and :312-366
// (`_dropblock_rois_get_entropy`): torchvision.ops.roi_align (third party, not vendored by the reference; its
// published RoIAlign: bilinear samples on a sampling_ratio x sampling_ratio grid per output bin, `aligned` half-pixel
// shift) followed by a mean (optionally a standard deviation) over the P x P bins of every (box, channel).
//
//   runia_roi_align_f32       the RoI maps themselves [K, C, P, P] (input of the MC-DropBlock sampler)
//   runia_roi_align_mean_f32  mean / std over the bins without materialising the maps: per box the sample positions and
//                             bilinear weights are computed ONCE into shared memory (they do not depend on the
//                             channel), then each warp walks its channels: lane = bin, 4 gathers per sample from the
//                             channel's H x W plane (L1 / L2 resident), warp reduction.
// Algorithmic bytes per box: the planes it touches (<= C*H*W*4, shared by overlapping boxes through L2) in, C*4 (*2) out.
#include <algorithm>

#include "common.cuh"

namespace runia {

struct RoiGeom {
  float start_h, start_w, bin_h, bin_w;
  int grid_h, grid_w;
};

__device__ __forceinline__ RoiGeom roi_geom(const float *box, float scale, int ph, int pw, int sampling_ratio, int aligned) {
  const float off = aligned ? 0.5f : 0.f;
  RoiGeom g;
  g.start_w = __fsub_rn(__fmul_rn(box[0], scale), off);
  g.start_h = __fsub_rn(__fmul_rn(box[1], scale), off);
  float roi_w = __fsub_rn(__fsub_rn(__fmul_rn(box[2], scale), off), g.start_w);
  float roi_h = __fsub_rn(__fsub_rn(__fmul_rn(box[3], scale), off), g.start_h);
  if (!aligned) {  // legacy behaviour: RoIs are at least one pixel wide
    roi_w = fmaxf(roi_w, 1.f);
    roi_h = fmaxf(roi_h, 1.f);
  }
  g.bin_h = roi_h / (float)ph;
  g.bin_w = roi_w / (float)pw;
  g.grid_h = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(roi_h / (float)ph);
  g.grid_w = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(roi_w / (float)pw);
  return g;
}

// the four corner offsets (into an H x W plane) and weights of one bilinear sample; outside [-1, H] x [-1, W]: weight 0
__device__ __forceinline__ void bilinear_setup(float y, float x, int H, int W, int (&o)[4], float (&w)[4]) {
  if (y < -1.f || y > (float)H || x < -1.f || x > (float)W) {
    o[0] = o[1] = o[2] = o[3] = 0;
    w[0] = w[1] = w[2] = w[3] = 0.f;
    return;
  }
  if (y <= 0.f) y = 0.f;
  if (x <= 0.f) x = 0.f;
  int y_low = (int)y, x_low = (int)x, y_high, x_high;
  if (y_low >= H - 1) {
    y_high = y_low = H - 1;
    y = (float)y_low;
  } else {
    y_high = y_low + 1;
  }
  if (x_low >= W - 1) {
    x_high = x_low = W - 1;
    x = (float)x_low;
  } else {
    x_high = x_low + 1;
  }
  const float ly = y - (float)y_low, lx = x - (float)x_low, hy = 1.f - ly, hx = 1.f - lx;
  o[0] = y_low * W + x_low;
  o[1] = y_low * W + x_high;
  o[2] = y_high * W + x_low;
  o[3] = y_high * W + x_high;
  w[0] = hy * hx;
  w[1] = hy * lx;
  w[2] = ly * hx;
  w[3] = ly * lx;
}

// Sample coordinate and bilinear sum with every product and sum rounded on its own (no FMA contraction), in the order
// of torchvision's CPU kernel, so that RoI values with cancellation agree with it to the last bits
__device__ __forceinline__ float sample_coord(float start, int p, float bin, int i, int grid) {
  return __fadd_rn(__fadd_rn(start, __fmul_rn((float)p, bin)), __fdiv_rn(__fmul_rn((float)i + 0.5f, bin), (float)grid));
}
__device__ __forceinline__ float bilinear_sum(const float *__restrict__ plane, const int *o, const float *w) {
  return __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(w[0], __ldg(plane + o[0])), __fmul_rn(w[1], __ldg(plane + o[1]))),
                             __fmul_rn(w[2], __ldg(plane + o[2]))),
                   __fmul_rn(w[3], __ldg(plane + o[3])));
}

__device__ __forceinline__ float bin_value(const float *__restrict__ plane, const RoiGeom &g, int H, int W, int ph_i, int pw_i) {
  float acc = 0.f;
  for (int iy = 0; iy < g.grid_h; ++iy) {
    const float y = sample_coord(g.start_h, ph_i, g.bin_h, iy, g.grid_h);
    for (int ix = 0; ix < g.grid_w; ++ix) {
      const float x = sample_coord(g.start_w, pw_i, g.bin_w, ix, g.grid_w);
      int o[4];
      float w[4];
      bilinear_setup(y, x, H, W, o, w);
      acc = __fadd_rn(acc, bilinear_sum(plane, o, w));
    }
  }
  const int count = g.grid_h * g.grid_w;
  return acc / (float)(count > 1 ? count : 1);
}

__global__ void __launch_bounds__(256)
roi_align_kernel(const float *__restrict__ feat, int C, int H, int W, const float *__restrict__ boxes,
                 const int32_t *__restrict__ batch_idx, int64_t total, int ph, int pw, float scale, int sampling_ratio,
                 int aligned, float *__restrict__ out) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int pw_i = (int)(e % pw), ph_i = (int)((e / pw) % ph);
    const int c = (int)((e / ((int64_t)pw * ph)) % C);
    const int64_t k = e / ((int64_t)pw * ph * C);
    const RoiGeom g = roi_geom(boxes + 4 * k, scale, ph, pw, sampling_ratio, aligned);
    const int b = batch_idx ? batch_idx[k] : 0;
    out[e] = bin_value(feat + ((int64_t)b * C + c) * H * W, g, H, W, ph_i, pw_i);
  }
}

// one block per box, 8 warps over the channels; sample table in shared memory
constexpr int kRoiTableMax = 1024;  // samples per box held in shared memory (7x7 bins x 4x4 grid, 14x14 x 2x2 ...): 32 KB

__global__ void __launch_bounds__(256)
roi_align_mean_kernel(const float *__restrict__ feat, int C, int H, int W, const float *__restrict__ boxes,
                      const int32_t *__restrict__ batch_idx, int ph, int pw, float scale, int sampling_ratio, int aligned,
                      float *__restrict__ out_mean, float *__restrict__ out_std) {
  __shared__ int s_off[kRoiTableMax][4];
  __shared__ float s_w[kRoiTableMax][4];
  const int64_t k = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const RoiGeom g = roi_geom(boxes + 4 * k, scale, ph, pw, sampling_ratio, aligned);
  const int per_bin = g.grid_h * g.grid_w, bins = ph * pw;
  const int64_t n_samples = (int64_t)bins * per_bin;
  const bool table = per_bin > 0 && n_samples <= kRoiTableMax;
  if (table) {
    for (int s = threadIdx.x; s < (int)n_samples; s += blockDim.x) {
      const int bin = s / per_bin, r = s - bin * per_bin;
      const int ph_i = bin / pw, pw_i = bin - ph_i * pw, iy = r / g.grid_w, ix = r - iy * g.grid_w;
      const float y = sample_coord(g.start_h, ph_i, g.bin_h, iy, g.grid_h);
      const float x = sample_coord(g.start_w, pw_i, g.bin_w, ix, g.grid_w);
      int o[4];
      float w[4];
      bilinear_setup(y, x, H, W, o, w);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        s_off[s][q] = o[q];
        s_w[s][q] = w[q];
      }
    }
  }
  __syncthreads();
  const int b = batch_idx ? batch_idx[k] : 0;
  const float count_f = (float)(per_bin > 1 ? per_bin : 1);
  for (int c = warp; c < C; c += 8) {
    const float *plane = feat + ((int64_t)b * C + c) * H * W;
    auto bin_of = [&](int bin) -> float {
      if (!table) return bin_value(plane, g, H, W, bin / pw, bin % pw);
      float acc = 0.f;
      const int s0 = bin * per_bin;
      for (int r = 0; r < per_bin; ++r) {
        const int s = s0 + r;
        acc = __fadd_rn(acc, bilinear_sum(plane, s_off[s], s_w[s]));
      }
      return acc / count_f;  // a division, like torchvision's output_val / count
    };
    float sum = 0.f;
    for (int bin = lane; bin < bins; bin += 32) sum += bin_of(bin);
    const float mean = warp_sum32(sum) / (float)bins;
    if (lane == 0) out_mean[k * C + c] = mean;
    if (out_std) {  // torch.std: unbiased, two passes
      float ss = 0.f;
      for (int bin = lane; bin < bins; bin += 32) {
        const float dlt = bin_of(bin) - mean;
        ss = fmaf(dlt, dlt, ss);
      }
      ss = warp_sum32(ss);
      if (lane == 0) out_std[k * C + c] = sqrtf(ss / (float)(bins - 1));  // bins == 1: NaN like torch
    }
  }
}

}  // namespace runia

using namespace runia;

static int roi_check(const float *feat, const float *boxes, int B, int C, int H, int W, int64_t K, int ph, int pw) {
  RUNIA_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && K >= 0 && ph > 0 && pw > 0, RUNIA_E_BADARG, "roi_align: bad sizes");
  RUNIA_REQUIRE(K == 0 || (feat && boxes), RUNIA_E_BADARG, "roi_align: null pointer");
  return RUNIA_OK;
}

extern "C" int runia_roi_align_f32(const float *feat, int B, int C, int H, int W, const float *boxes, const int32_t *batch_idx,
                                   int64_t K, int pooled_h, int pooled_w, float spatial_scale, int sampling_ratio,
                                   int aligned, float *out, void *stream) {
  RUNIA_NVTX();
  int rc = roi_check(feat, boxes, B, C, H, W, K, pooled_h, pooled_w);
  if (rc) return rc;
  if (K == 0) return RUNIA_OK;
  RUNIA_REQUIRE(out, RUNIA_E_BADARG, "roi_align: null output");
  const int64_t total = K * C * pooled_h * pooled_w;
  const unsigned grid = (unsigned)std::min<int64_t>(ceil_div(total, 256), (int64_t)kNumSMs * 16);
  roi_align_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(feat, C, H, W, boxes, batch_idx, total, pooled_h, pooled_w,
                                                         spatial_scale, sampling_ratio, aligned, out);
  count_launch();
  return finish_launch("roi_align");
}

extern "C" int runia_roi_align_mean_f32(const float *feat, int B, int C, int H, int W, const float *boxes,
                                        const int32_t *batch_idx, int64_t K, int pooled_h, int pooled_w, float spatial_scale,
                                        int sampling_ratio, int aligned, float *out_mean, float *out_std, void *stream) {
  RUNIA_NVTX();
  int rc = roi_check(feat, boxes, B, C, H, W, K, pooled_h, pooled_w);
  if (rc) return rc;
  if (K == 0) return RUNIA_OK;
  RUNIA_REQUIRE(out_mean, RUNIA_E_BADARG, "roi_align_mean: null output");
  RUNIA_REQUIRE(K < (int64_t)0x7fffffff, RUNIA_E_UNSUPPORTED, "roi_align_mean: too many boxes");
  roi_align_mean_kernel<<<(unsigned)K, 256, 0, (cudaStream_t)stream>>>(feat, C, H, W, boxes, batch_idx, pooled_h, pooled_w,
                                                                     spatial_scale, sampling_ratio, aligned, out_mean, out_std);
  count_launch();
  return finish_launch("roi_align_mean");
}
