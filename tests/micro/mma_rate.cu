// mma_rate.cu — development micro-benchmark: clocks per tcgen05.mma (kind::f16, cta_group::1, M = 128, K = 16) as a function of
// N, with both operands in shared memory (SS) or A in tensor memory (TS), issued back to back by one thread of one CTA per SM.
// Operands are whatever the buffers hold (zeros): only the pipe's pace is measured.  A second mode adds concurrent TMA-like
// shared-memory write pressure from the other warps (st.shared streams) to see whether the operand port is shared.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I ape_b200/csrc -o tests/micro/mma_rate tests/micro/mma_rate.cu -lcuda
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>

#include "tc.cuh"

using namespace ape;

struct alignas(1024) Smem {
  uint8_t a[128 * 64 * 2];
  uint8_t b[256 * 64 * 2];
  uint8_t junk[64 * 1024];
  uint64_t done;
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(256, 1) k(int n, int ts, int count, int pressure, long long *cycles) {
  extern __shared__ uint8_t raw[];
  Smem &s = *reinterpret_cast<Smem *>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (int)(sizeof(s.a) + sizeof(s.b)) / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(s.a)[i] = 0;
  if (threadIdx.x == 0) {
    tc::mbar_init(&s.done, 1);
    tc::fence_mbar_init();
  }
  if (warp == 1) {
    tc::tmem_alloc(&s.tmem_base, 512);
    tc::tmem_relinquish();
  }
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = s.tmem_base;
  if (warp == 1 && lane == 0) {
    const uint32_t idesc = tc::make_idesc_f16(128, n, 0);
    const uint64_t da = tc::make_smem_desc_sw128(tc::smem_u32(s.a));
    const uint64_t db = tc::make_smem_desc_sw128(tc::smem_u32(s.b));
    const long long t0 = clock64();
    for (int i = 0; i < count; ++i) {
      const int kk = i & 3;
      if (ts) tc::mma_f16_ts(tmem, tmem + 256 + 8 * kk, db + 2 * kk, idesc, i != 0);
      else tc::mma_f16(tmem, da + 2 * kk, db + 2 * kk, idesc, i != 0);
    }
    tc::mma_commit(&s.done);
    tc::mbar_wait(&s.done, 0);
    const long long t1 = clock64();
    cycles[blockIdx.x] = t1 - t0;
    s.tmem_base = 0xffffffffu;  // tells the pressure warps to stop
  } else if (warp >= 2 && pressure) {
    // warps 2..7 stream 16-byte stores over a 64 KB scratch area until the MMAs are done
    volatile uint32_t *flag = &s.tmem_base;
    uint4 *dst = reinterpret_cast<uint4 *>(s.junk);
    uint4 v = make_uint4(lane, warp, 0, 0);
    int off = (warp - 2) * 32 + lane;
    while (*flag != 0xffffffffu) {
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        dst[off] = v;
        off = (off + 192) & 4095;
      }
    }
  }
  __syncthreads();
  if (warp == 1) {
    tc::fence_after_sync();
    tc::tmem_dealloc(tmem, 512);
  }
}

int main() {
  long long *cyc;
  cudaMalloc(&cyc, 148 * sizeof(long long));
  const int smem = sizeof(Smem) + 1024;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int count = 4096;
  for (int pressure = 0; pressure < 2; ++pressure)
    for (int ts = 0; ts < 2; ++ts)
      for (int n : {64, 128, 256}) {
        for (int ctas : {1, 148}) {
          k<<<ctas, 256, smem>>>(n, ts, 64, pressure, cyc);
          k<<<ctas, 256, smem>>>(n, ts, count, pressure, cyc);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
          long long h[148];
          cudaMemcpy(h, cyc, ctas * sizeof(long long), cudaMemcpyDeviceToHost);
          double mean = 0;
          for (int i = 0; i < ctas; ++i) mean += (double)h[i];
          mean /= ctas;
          printf("M=128 N=%3d K=16 %s  smem-store pressure %d  CTAs %3d: %7.1f clk per MMA  (%.0f flop/clk/SM)\n", n, ts ? "TS" : "SS", pressure,
                 ctas, mean / count, 2.0 * 128 * n * 16 / (mean / count));
        }
      }
  return 0;
}
