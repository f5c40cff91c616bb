// Issue / pipe model of the min-max and add forms the entropy kernels are built from (sm_100a).  The first probe
// (fmnmx_probe.cu) let ptxas fuse its "2-input" chain into FMNMX3, so its FMNMX row measured FMNMX3; here every mode's
// SASS mix is checked with cuobjdump (scripts/probes/pipe_probe_sass.sh) before the numbers are read.
// Each thread keeps independent accumulators; 512 threads (4 warps per scheduler) per SM, 148 CTAs.
// build + run: nvcc -arch=sm_100a -o /tmp/pipe_probe scripts/probes/pipe_probe.cu && /tmp/pipe_probe
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float2 sub2(float2 a, float2 b) {
  float2 r;
  asm("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; sub.f32x2 rc, ra, rb; mov.b64 {%0, %1}, rc;}"
      : "=f"(r.x), "=f"(r.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return r;
}

constexpr int NA = 12;

template <int MODE>
__global__ void __launch_bounds__(512) probe(float *out, const float *in, int iters, long long *cycles) {
  float a[NA], c[NA];
#pragma unroll
  for (int i = 0; i < NA; ++i) {
    a[i] = in[threadIdx.x + 32 * i];
    c[i] = in[threadIdx.x + 32 * i + 700];
  }
  float x = in[threadIdx.x + 1000], y = in[threadIdx.x + 2000];
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NA; ++i) {
      if (MODE == 0) a[i] = fminf(fmaxf(a[i], x), y);                    // 2 FMNMX (max then min: not fusable)
      if (MODE == 1) a[i] = fmaxf(fmaxf(a[i], x), c[i]);                 // 1 FMNMX3
      if (MODE == 2) a[i] = (a[i] + x) + y;                              // 2 FADD
      if (MODE == 3 && (i & 1) == 0) {                                   // 1 FADD2 per two accumulators
        const float2 r = sub2(make_float2(a[i], a[i + 1]), make_float2(x, y));
        a[i] = r.x, a[i + 1] = r.y;
      }
      if (MODE == 4) { a[i] = fminf(fmaxf(a[i], x), y); c[i] = (c[i] + x) + y; }          // 2 FMNMX + 2 FADD
      if (MODE == 5) { a[i] = fmaxf(fmaxf(a[i], x), y); c[i] = (c[i] + x) + y; }          // 1 FMNMX3 + 2 FADD
      if (MODE == 6) {                                                                    // 1 FMNMX3 + 1 FADD2, independent
        a[i] = fmaxf(fmaxf(a[i], x), y);
        if ((i & 1) == 0) { const float2 r = sub2(make_float2(c[i], c[i + 1]), make_float2(x, y)); c[i] = r.x, c[i + 1] = r.y; }
        else { const float2 r = sub2(make_float2(c[i], c[i - 1]), make_float2(y, x)); c[i] = r.x, c[i - 1] = r.y; }
      }
      if (MODE == 7) {                                                                    // joint update, packed: FADD2 + FMNMX3
        const float2 r = sub2(make_float2(x, y), make_float2(c[i], c[(i + 1) % NA]));
        a[i] = fmaxf(fmaxf(a[i], fabsf(r.x)), fabsf(r.y));
      }
      if (MODE == 8) {                                                                    // joint update, scalar: 2 FADD + FMNMX3
        const float d0 = x - c[i], d1 = y - c[(i + 1) % NA];
        a[i] = fmaxf(fmaxf(a[i], fabsf(d0)), fabsf(d1));
      }
      if (MODE == 9) {                                                                    // 2 FADD + 2 FMNMX (max then min)
        const float d0 = x - c[i], d1 = y - c[(i + 1) % NA];
        a[i] = fminf(fmaxf(a[i], fabsf(d0)), fabsf(d1));
      }
      if (MODE == 10) { a[i] = fminf(fmaxf(a[i], x), y); c[i] = fmaf(c[i], x, y); }       // 2 FMNMX + 1 FFMA
      if (MODE == 11) {                                                                   // integer min / max pair
        int v = __float_as_int(a[i]);
        v = min(max(v, __float_as_int(x)), __float_as_int(y));
        a[i] = __int_as_float(v);
      }
      if (MODE == 12) {                                                                   // 2 FMNMX + integer min/max pair
        a[i] = fminf(fmaxf(a[i], x), y);
        int v = __float_as_int(c[i]);
        v = min(max(v, __float_as_int(x)), __float_as_int(y));
        c[i] = __int_as_float(v);
      }
    }
    if (MODE == 7 || MODE == 8 || MODE == 9) {  // keep the differences loop-variant
      x = __int_as_float(__float_as_int(x) ^ it);
      y = __int_as_float(__float_as_int(y) ^ (it << 1));
    }
  }
  const long long t1 = clock64();
  float s = x + y;
#pragma unroll
  for (int i = 0; i < NA; ++i) s += a[i] + c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int MODE>
void run(const char *name, float *out, float *in, long long *cyc) {
  const int iters = 4096;
  long long h;
  for (int rep = 0; rep < 2; ++rep) {
    probe<MODE><<<148, 512>>>(out, in, iters, cyc);
    cudaDeviceSynchronize();
  }
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  // cycles per scheduler per (one inner-loop body of one accumulator): 4 warps x NA bodies x iters
  printf("{\"mode\": %d, \"body\": \"%s\", \"cycles\": %lld, \"cycles_per_body_per_scheduler\": %.3f}\n", MODE, name, h,
         (double)h / (4.0 * NA * iters));
}

int main() {
  float *in, *out;
  long long *cyc;
  cudaMalloc(&in, 1 << 20);
  cudaMalloc(&out, 148 * 512 * 4);
  cudaMalloc(&cyc, 8);
  cudaMemset(in, 0, 1 << 20);
  run<0>("2 FMNMX", out, in, cyc);
  run<1>("1 FMNMX3", out, in, cyc);
  run<2>("2 FADD", out, in, cyc);
  run<3>("0.5 FADD2", out, in, cyc);
  run<4>("2 FMNMX + 2 FADD", out, in, cyc);
  run<5>("1 FMNMX3 + 2 FADD", out, in, cyc);
  run<6>("1 FMNMX3 + 1 FADD2 (independent)", out, in, cyc);
  run<7>("joint update packed: FADD2 + FMNMX3 ", out, in, cyc);
  run<8>("joint update scalar: 2 FADD + FMNMX3 ", out, in, cyc);
  run<9>("2 FADD + 2 FMNMX ", out, in, cyc);
  run<10>("2 FMNMX + 1 FFMA", out, in, cyc);
  run<11>("2 integer min/max", out, in, cyc);
  run<12>("2 FMNMX + 2 integer min/max", out, in, cyc);
  return cudaGetLastError() != cudaSuccess;
}
