// Issue rate of FMNMX (2-input) vs FMNMX3 (3-input min / max) vs FADD on sm_100a: the entropy kernels' ceiling depends
// on it.  Each thread keeps 16 independent accumulators; a block of 512 threads (4 warps per scheduler) per SM.
// build + run: nvcc -arch=sm_100a -o /tmp/fmnmx_probe scripts/probes/fmnmx_probe.cu && /tmp/fmnmx_probe
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float2 sub2(float2 a, float2 b) {
  float2 r;
  asm("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; sub.f32x2 rc, ra, rb; mov.b64 {%0, %1}, rc;}"
      : "=f"(r.x), "=f"(r.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return r;
}

template <int MODE>
__global__ void __launch_bounds__(512) probe(float *out, const float *in, int iters, long long *cycles) {
  float a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = in[threadIdx.x + 32 * i];
  const float x = in[threadIdx.x + 1000], y = in[threadIdx.x + 2000];
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const float xs = (it & 1) ? x : y;  // keeps the compiler from collapsing the idempotent max
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (MODE == 0) a[i] = fmaxf(a[i], xs);                   // FMNMX
      if (MODE == 1) a[i] = fmaxf(fmaxf(a[i], x), y);          // FMNMX3
      if (MODE == 2) a[i] = a[i] + x;                          // FADD
      if (MODE == 3) a[i] = fmaxf(a[i], fabsf(x - a[(i + 1) & 15]));  // FADD + FMNMX(|.|)
      if (MODE == 4 && (i & 1) == 0) {                                // FADD2 (sub.f32x2): 8 per inner loop
        const float2 r = sub2(make_float2(a[i], a[i + 1]), make_float2(xs, y));
        a[i] = r.x;
        a[i + 1] = r.y;
      }
      if (MODE == 5 && (i & 1) == 0) {                                // FADD2 + FMNMX3: the joint estimator's update
        const float2 r = sub2(make_float2(x, y), make_float2(a[(i + 2) & 15], a[(i + 3) & 15]));
        a[i] = fmaxf(fmaxf(a[i], fabsf(r.x)), fabsf(r.y));
      }
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

int main() {
  float *in, *out;
  long long *cyc, h;
  cudaMalloc(&in, 1 << 20);
  cudaMalloc(&out, 148 * 512 * 4);
  cudaMalloc(&cyc, 8);
  cudaMemset(in, 0, 1 << 20);
  const int iters = 4096;
  const char *names[] = {"FMNMX", "FMNMX3", "FADD", "FADD+FMNMX", "FADD2", "FADD2+FMNMX3"};
  for (int m = 0; m < 6; ++m) {
    for (int rep = 0; rep < 2; ++rep) {
      if (m == 0) probe<0><<<148, 512>>>(out, in, iters, cyc);
      if (m == 1) probe<1><<<148, 512>>>(out, in, iters, cyc);
      if (m == 2) probe<2><<<148, 512>>>(out, in, iters, cyc);
      if (m == 3) probe<3><<<148, 512>>>(out, in, iters, cyc);
      if (m == 4) probe<4><<<148, 512>>>(out, in, iters, cyc);
      if (m == 5) probe<5><<<148, 512>>>(out, in, iters, cyc);
      cudaDeviceSynchronize();
    }
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    // per scheduler: 4 warps x 16 x iters warp instructions (x2 in mode 3)
    const double winstr = 4.0 * iters * (m == 3 ? 32 : m == 4 ? 8 : m == 5 ? 16 : 16);
    printf("{\"op\": \"%s\", \"cycles\": %lld, \"cycles_per_warp_instruction_per_scheduler\": %.3f}\n", names[m], h,
           (double)h / winstr);
  }
  return 0;
}
