// ex2_rate.cu — development micro-benchmark: exponentials per clock per SM on B200 for the softmax inner loop.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tests/micro/ex2_rate tests/micro/ex2_rate.cu && ./tests/micro/ex2_rate
// Variants: 0 = MUFU.EX2 (ex2.approx.ftz.f32), 1 = ex2.approx.ftz.f16x2 (two per instruction), 2 = degree-3 polynomial on
// the FMA pipe (Cody-Waite split, exponent through integer add), 3 = every 4th element on the FMA pipe, the rest MUFU,
// 4 = half / half.  Each thread owns 64 independent values (one softmax row block), as the attention kernel does.
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2_mufu(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ unsigned ex2_h2(unsigned x) {
  unsigned y;
  asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x));
  return y;
}

// 2^x for x <= 0 (softmax arguments), relative error ~1e-4: x = n + f, f in [0, 1), 2^f by a cubic, 2^n by exponent add
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -126.f);
  const float n = floorf(x);
  const float f = x - n;
  float p = fmaf(f, 0.0790210f, 0.2240570f);
  p = fmaf(p, f, 0.6965820f);
  p = fmaf(p, f, 0.9999252f);
  return __int_as_float(__float_as_int(p) + ((int)n << 23));
}

template <int V>
__global__ void __launch_bounds__(128) k(float *out, float seed, int iters, long long *cycles) {
  float v[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) v[i] = -seed * (float)(i + 1 + threadIdx.x % 7);
  float acc = 0.f;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (V == 1) {
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const __half2 h = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
        const unsigned r = ex2_h2(*reinterpret_cast<const unsigned *>(&h));
        const float2 f = __half22float2(*reinterpret_cast<const __half2 *>(&r));
        acc += f.x + f.y;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 64; ++i) {
        const bool poly = V == 2 || (V == 3 && (i & 3) == 3) || (V == 4 && (i & 1));
        acc += poly ? ex2_poly(v[i]) : ex2_mufu(v[i]);
      }
    }
#pragma unroll
    for (int i = 0; i < 64; ++i) v[i] = fmaf(v[i], 0.999f, -1e-3f);  // new arguments every round (1 FFMA per element, as the scale-subtract)
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int V>
void run(const char *name, int ctas_per_sm) {
  const int sms = 148, iters = 2000;
  float *out;
  long long *cyc;
  cudaMalloc(&out, sizeof(float) * sms * ctas_per_sm * 128);
  cudaMalloc(&cyc, sizeof(long long) * sms * ctas_per_sm);
  k<V><<<sms * ctas_per_sm, 128>>>(out, 0.01f, 10, cyc);
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  cudaEventRecord(a);
  k<V><<<sms * ctas_per_sm, 128>>>(out, 0.01f, iters, cyc);
  cudaEventRecord(b);
  cudaDeviceSynchronize();
  float ms = 0.f;
  cudaEventElapsedTime(&ms, a, b);
  long long h[148 * 8];
  cudaMemcpy(h, cyc, sizeof(long long) * sms * ctas_per_sm, cudaMemcpyDeviceToHost);
  double mean = 0;
  for (int i = 0; i < sms * ctas_per_sm; ++i) mean += (double)h[i];
  mean /= sms * ctas_per_sm;
  const double elems_per_sm = (double)ctas_per_sm * 128 * 64 * iters;
  printf("%-28s ctas/SM %d  %8.3f ms  %6.2f exp/clk/SM (clock64)  %7.2f Gexp/s/SM (events)\n", name, ctas_per_sm, ms, elems_per_sm / mean,
         elems_per_sm / (ms * 1e6));
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  for (int c = 1; c <= 4; c *= 2) {
    run<0>("mufu f32", c);
    run<1>("mufu f16x2", c);
    run<2>("fma cubic", c);
    run<3>("1/4 fma + 3/4 mufu", c);
    run<4>("1/2 fma + 1/2 mufu", c);
  }
  return 0;
}
